// K2, fused form: pooled cross-spectral density + coherence straight from K-major spectra rows
// ([bin][channel][segment] complex64, written by cmc_fft_segments_kmajor) - no pack pass.
//
// The rows ARE the K-major "X" / "Y" operands of the contraction in csd_tc.cu (K = (segment, re/im)
// contiguous), so TMA stages them as they lie in HBM: 8 KB of X and 8 KB of Y per k-block instead of the
// 48 KB of pre-packed planes.  What the pack kernel used to materialise is derived on chip by four
// converter warps between the TMA and the MMA issue:
//     A rows 0..63   = tf32(X)      (rounded in place)        A_lo = tf32(X - tf32(X))
//     A rows 64..127 = i * tf32(X)  (-im, re)                 and its lo plane
//     B rows         = tf32(Y)                                B_lo
//     Pxx, Pyy       = sum_l |X|^2, |Y|^2 per row, accumulated in a fixed order
// after which the 3xTF32 MMAs, the TMEM accumulators and the normalising epilogue are those of
// csd_gemm_kernel<0>.  HBM traffic drops from 69 MB (operand planes) + the pack pass to the 21.5 MB of
// spectra for the 64 x 64 x 100-bin, L = 210 configuration.
//
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue,
// warps 8-11 converters.  Stage ring: full (TMA bytes) -> conv (128 converter arrivals) -> MMAs -> empty.
#include "common.cuh"
#include "tc_common.cuh"
#include "csd_layout.cuh"

namespace cmc {

using namespace tc;

constexpr int kFuThreads = 384;
constexpr int kFuStages = 3;
constexpr int kFuABytes = kTileM * kKBlock * 4;          // 16 KB: X rows then i*X rows
constexpr int kFuBBytes = kTileN * kKBlock * 4;          // 8 KB
constexpr int kFuStageBytes = 2 * (kFuABytes + kFuBBytes);   // A_hi, A_lo, B_hi, B_lo = 48 KB
constexpr int kFuHalfA = kFuABytes / 2;                  // 8 KB: offset of the i*X rows inside an A plane
constexpr int kFuPitch = kTileN + 1;                     // padded staging row (floats)

struct FusedParams {
    int F, MT, NT, Ne, Nm, KB;
    float* coh;           // [F][Ne][Nm]
    float2* sxy;          // optional
    float* pxx;           // [F][Ne] workspace copy (consumed by the surrogate kernels)
    float* pyy;           // [F][Nm]
    float* sxx_out;       // optional user outputs
    float* syy_out;
    long long total_tiles;
};

struct __align__(8) FusedBarriers {
    uint64_t full[kFuStages];
    uint64_t conv[kFuStages];
    uint64_t empty[kFuStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint64_t pw_full[2];
    uint64_t pw_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ float4 lds128f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128f(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(kFuThreads, 1)
csd_fused_kernel(const __grid_constant__ CUtensorMap mX, const __grid_constant__ CUtensorMap mY, const FusedParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* sS = base;                                                         // [kFuStages][48 KB]
    float* stage_tile = reinterpret_cast<float*>(sS + kFuStages * kFuStageBytes);     // [128][65]
    float* pw = stage_tile + kTileM * kFuPitch;                                        // [2][128]: Pxx rows, Pyy rows
    FusedBarriers* bars = reinterpret_cast<FusedBarriers*>(pw + 2 * 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = p.total_tiles * blockIdx.x / gridDim.x;
    const long long t1 = p.total_tiles * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kFuStages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->conv[s], 128);
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->tmem_full[a], 1);
            mbar_init(&bars->tmem_empty[a], 4);
            mbar_init(&bars->pw_full[a], 128);
            mbar_init(&bars->pw_empty[a], 128);
        }
        fence_barrier_init();
        tma_prefetch_desc(&mX);
        tma_prefetch_desc(&mY);
    }
    if (warp == 2) {
        tmem_alloc(&bars->tmem_base, 2 * kTileN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = t0; t < t1; ++t) {
                const int nt = (int)(t % p.NT);
                const long long r = t / p.NT;
                const int mt = (int)(r % p.MT), f = (int)(r / p.MT);
                const int xrow = f * p.Ne + mt * 64, yrow = f * p.Nm + nt * 64;
                for (int kb = 0; kb < p.KB; ++kb) {
                    unsigned char* st = sS + stage * kFuStageBytes;
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars->full[stage], 2 * kFuBBytes);
                    tma_load_2d(st, &mX, &bars->full[stage], kb * kKBlock, xrow);                   // X rows of A_hi
                    tma_load_2d(st + 2 * kFuABytes, &mY, &bars->full[stage], kb * kKBlock, yrow);   // B_hi
                    if (++stage == kFuStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(kTileM, kTileN);
            int stage = 0;
            uint32_t phase = 0, it = 0;
            for (long long t = t0; t < t1; ++t) {
                const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
                mbar_wait(&bars->tmem_empty[acc], accphase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * kTileN;
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(&bars->conv[stage], phase);
                    tc_fence_after();
                    const uint32_t ahi = smem_u32(sS + stage * kFuStageBytes), alo = ahi + kFuABytes;
                    const uint32_t bhi = ahi + 2 * kFuABytes, blo = bhi + kFuBBytes;
#pragma unroll
                    for (int k = 0; k < kKBlock / 8; ++k) {
                        const uint64_t dah = make_smem_desc_k_sw128(ahi + k * 32), dal = make_smem_desc_k_sw128(alo + k * 32);
                        const uint64_t dbh = make_smem_desc_k_sw128(bhi + k * 32), dbl = make_smem_desc_k_sw128(blo + k * 32);
                        umma_tf32(d, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);      // small terms first
                        umma_tf32(d, dah, dbl, idesc, 1u);
                        umma_tf32(d, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&bars->empty[stage]);
                    if (++stage == kFuStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars->tmem_full[acc]);
                ++it;
            }
        }
    } else if (warp >= 8) {
        // ===================== converters: hi/lo split, i*X rows, auto-spectra =====================
        const int tc_ = threadIdx.x - 256;           // 0..127
        const int r = tc_ >> 1, h = tc_ & 1;         // row of the 64-row tile, half of its 128-byte k-block
        const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        int stage = 0;
        uint32_t phase = 0, it = 0;
        for (long long t = t0; t < t1; ++t) {
            float px = 0.f, py = 0.f;
            for (int kb = 0; kb < p.KB; ++kb) {
                mbar_wait(&bars->full[stage], phase);
                const uint32_t ahi = smem_u32(sS + stage * kFuStageBytes), alo = ahi + kFuABytes;
                const uint32_t bhi = ahi + 2 * kFuABytes, blo = bhi + kFuBBytes;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t off = row_off + (uint32_t)((((4 * h + c) ^ (r & 7))) << 4);   // SWIZZLE_128B chunk
                    const float4 x = lds128f(ahi + off);
                    const float4 y = lds128f(bhi + off);
                    px += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
                    py += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
                    const float4 xh = make_float4(to_tf32(x.x), to_tf32(x.y), to_tf32(x.z), to_tf32(x.w));
                    const float4 xl = make_float4(to_tf32(x.x - xh.x), to_tf32(x.y - xh.y), to_tf32(x.z - xh.z),
                                                  to_tf32(x.w - xh.w));
                    const float4 yh = make_float4(to_tf32(y.x), to_tf32(y.y), to_tf32(y.z), to_tf32(y.w));
                    const float4 yl = make_float4(to_tf32(y.x - yh.x), to_tf32(y.y - yh.y), to_tf32(y.z - yh.z),
                                                  to_tf32(y.w - yh.w));
                    sts128f(ahi + off, xh);
                    sts128f(alo + off, xl);
                    sts128f(ahi + kFuHalfA + off, make_float4(-xh.y, xh.x, -xh.w, xh.z));   // i * X = (-im, re)
                    sts128f(alo + kFuHalfA + off, make_float4(-xl.y, xl.x, -xl.w, xl.z));
                    sts128f(bhi + off, yh);
                    sts128f(blo + off, yl);
                }
                fence_proxy_async();                 // generic-proxy writes -> visible to the MMA's async-proxy reads
                mbar_arrive(&bars->conv[stage]);
                if (++stage == kFuStages) { stage = 0; phase ^= 1; }
            }
            // auto-spectra of this tile's rows: the two halves of a row sit in adjacent lanes
            px += __shfl_xor_sync(0xffffffffu, px, 1);
            py += __shfl_xor_sync(0xffffffffu, py, 1);
            const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
            mbar_wait(&bars->pw_empty[acc], accphase ^ 1);
            if (h == 0) {
                pw[acc * 128 + r] = px;
                pw[acc * 128 + 64 + r] = py;
            }
            mbar_arrive(&bars->pw_full[acc]);
            ++it;
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int q = warp - 4;                  // TMEM lane quadrant of this warp
        const int te = threadIdx.x - 128;        // 0..127
        uint32_t it = 0;
        for (long long t = t0; t < t1; ++t) {
            const int nt = (int)(t % p.NT);
            const long long rr = t / p.NT;
            const int mt = (int)(rr % p.MT), f = (int)(rr / p.MT);
            const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
            mbar_wait(&bars->tmem_full[acc], accphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * kTileN + (static_cast<uint32_t>(q * 32) << 16);
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(taddr, r0);
            tmem_ld_32x32(taddr + 32, r1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);   // accumulator may be overwritten
            float* row = stage_tile + (q * 32 + lane) * kFuPitch;
#pragma unroll
            for (int cidx = 0; cidx < 32; ++cidx) {
                row[cidx] = __uint_as_float(r0[cidx]);
                row[32 + cidx] = __uint_as_float(r1[cidx]);
            }
            mbar_wait(&bars->pw_full[acc], accphase);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const float* pwa = pw + acc * 128;
            {   // auto-spectra outputs (every pair tile of a row block carries the same values)
                const bool is_x = te < 64;
                const int ch = (is_x ? mt : nt) * 64 + (te & 63);
                if (is_x ? (nt == 0 && ch < p.Ne) : (mt == 0 && ch < p.Nm)) {
                    const long long o = (long long)f * (is_x ? p.Ne : p.Nm) + ch;
                    (is_x ? p.pxx : p.pyy)[o] = pwa[te];
                    float* user = is_x ? p.sxx_out : p.syy_out;
                    if (user) user[o] = pwa[te];
                }
            }
#pragma unroll 4
            for (int n = 0; n < 32; ++n) {
                const int idx = te + 128 * n;
                const int il = idx >> 6, jl = idx & 63;
                const int i = mt * 64 + il, j = nt * 64 + jl;
                if (i < p.Ne && j < p.Nm) {
                    const float re = stage_tile[il * kFuPitch + jl];
                    const float im = stage_tile[(64 + il) * kFuPitch + jl];
                    const float pxv = pwa[il], pyv = pwa[64 + jl];
                    // |S|^2 / (Pxx Pyy) as |S / sqrt(Pxx) / sqrt(Pyy)|^2: no overflow, zero-power channels give 0
                    const float sc = (pxv > 0.f && pyv > 0.f) ? rsqrtf(pxv) * rsqrtf(pyv) : 0.f;
                    const float a = re * sc, b = im * sc;
                    const long long o = ((long long)f * p.Ne + i) * p.Nm + j;
                    p.coh[o] = fminf(a * a + b * b, 1.0f);
                    if (p.sxy) p.sxy[o] = make_float2(re, im);
                }
            }
            mbar_arrive(&bars->pw_empty[acc]);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            ++it;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 2 * kTileN);
}

// Operand planes for the surrogate kernels from K-major rows (the layout csd_tc.cu's pack kernel produces):
// elementwise and coalesced along K; columns k >= 2L and the rows of padding channels are written as zeros.
// grid (padded rows of X then padded rows of Y), block 128.
__global__ void __launch_bounds__(128)
planes_from_kmajor_kernel(const float2* __restrict__ Xk, int64_t pitch_x, const float2* __restrict__ Yk, int64_t pitch_y,
                          int L, int F, int Ne, int Nm, int MT, int NT, int KP, float* __restrict__ Ahi,
                          float* __restrict__ Alo, float* __restrict__ Bhi, float* __restrict__ Blo) {
    const int64_t rix = blockIdx.x;
    const int64_t a_rows = (int64_t)F * MT * 64;
    const bool is_a = rix < a_rows;
    const int64_t rloc = is_a ? rix : rix - a_rows;
    const int CP = (is_a ? MT : NT) * 64, C = is_a ? Ne : Nm;
    const int f = (int)(rloc / CP), ch = (int)(rloc % CP);
    const float2* src = is_a ? Xk + ((int64_t)f * Ne + ch) * pitch_x : Yk + ((int64_t)f * Nm + ch) * pitch_y;
    for (int l = threadIdx.x; l < KP / 2; l += blockDim.x) {
        const float2 v = (l < L && ch < C) ? __ldg(src + l) : make_float2(0.f, 0.f);
        const float2 hv = make_float2(to_tf32(v.x), to_tf32(v.y));
        const float2 lv = make_float2(to_tf32(v.x - hv.x), to_tf32(v.y - hv.y));
        if (is_a) {
            const int64_t o = (((int64_t)f * MT + (ch >> 6)) * kTileM + (ch & 63)) * KP + 2 * l;
            *reinterpret_cast<float2*>(Ahi + o) = hv;
            *reinterpret_cast<float2*>(Alo + o) = lv;
            const int64_t o2 = o + (int64_t)64 * KP;
            *reinterpret_cast<float2*>(Ahi + o2) = make_float2(-hv.y, hv.x);
            *reinterpret_cast<float2*>(Alo + o2) = make_float2(-lv.y, lv.x);
        } else {
            const int64_t o = (((int64_t)f * NT + (ch >> 6)) * kTileN + (ch & 63)) * KP + 2 * l;
            *reinterpret_cast<float2*>(Bhi + o) = hv;
            *reinterpret_cast<float2*>(Blo + o) = lv;
        }
    }
}

// K-major spectra rows as a TMA tensor: dim0 = 2L floats (columns past 2L read as zeros), dim1 = rows
static int make_rows_map(CUtensorMap* m, const float* base, int L, int64_t pitch_complex, int64_t rows) {
    EncodeTiledFn enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    cuuint64_t dims[2] = {(cuuint64_t)(2 * (int64_t)L), (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch_complex * 8};
    cuuint32_t box[2] = {(cuuint32_t)kKBlock, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (L=%d pitch=%lld rows=%lld)", (int)r, L,
                  (long long)pitch_complex, (long long)rows);
        return CMC_ECUDA;
    }
    return CMC_OK;
}

}  // namespace cmc

namespace cmc {
static int kmajor_check(const float* Xk, const float* Yk, int L, int F, int Ne, int Nm, int64_t pitch_x, int64_t pitch_y,
                        const void* ws) {
    CMC_REQUIRE(Xk && Yk && ws, "cmc_csd_*_kmajor: null pointer");
    CMC_REQUIRE(L >= 1 && F >= 1 && Ne >= 1 && Nm >= 1, "cmc_csd_*_kmajor: bad shape");
    CMC_REQUIRE(pitch_x >= L && pitch_y >= L && (pitch_x & 1) == 0 && (pitch_y & 1) == 0,
                "cmc_csd_*_kmajor: row pitch must be even and >= L");
    CMC_REQUIRE((reinterpret_cast<uintptr_t>(Xk) & 15) == 0 && (reinterpret_cast<uintptr_t>(Yk) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(ws) & 255) == 0,
                "cmc_csd_*_kmajor: spectra must be 16-byte and the workspace 256-byte aligned");
    CMC_REQUIRE((int64_t)F * Ne < (1ll << 31) && (int64_t)F * Nm < (1ll << 31), "cmc_csd_*_kmajor: too many rows");
    return CMC_OK;
}

static int launch_planes(const float* Xk, const float* Yk, const CsdLayout& y, int64_t pitch_x, int64_t pitch_y,
                         unsigned char* w, cudaStream_t st) {
    planes_from_kmajor_kernel<<<(unsigned)((int64_t)y.F * 64 * (y.MT + y.NT)), 128, 0, st>>>(
        reinterpret_cast<const float2*>(Xk), pitch_x, reinterpret_cast<const float2*>(Yk), pitch_y, y.L, y.F, y.Ne, y.Nm,
        y.MT, y.NT, y.KP, reinterpret_cast<float*>(w + y.off_ahi), reinterpret_cast<float*>(w + y.off_alo),
        reinterpret_cast<float*>(w + y.off_bhi), reinterpret_cast<float*>(w + y.off_blo));
    CMC_CHECK_LAUNCH("planes_from_kmajor_kernel");
    return CMC_OK;
}
}  // namespace cmc

extern "C" int cmc_csd_operands_kmajor(const float* Xk, const float* Yk, int L, int F, int Ne, int Nm, int64_t pitch_x,
                                       int64_t pitch_y, void* ws, int64_t ws_bytes, void* stream) {
    using namespace cmc;
    int rc = kmajor_check(Xk, Yk, L, F, Ne, Nm, pitch_x, pitch_y, ws);
    if (rc) return rc;
    const CsdLayout y = csd_layout(L, F, Ne, Nm);
    if (ws_bytes < y.total) {
        set_error("cmc_csd_operands_kmajor: workspace %lld < %lld bytes", (long long)ws_bytes, (long long)y.total);
        return CMC_EWORKSPACE;
    }
    return launch_planes(Xk, Yk, y, pitch_x, pitch_y, static_cast<unsigned char*>(ws), static_cast<cudaStream_t>(stream));
}

extern "C" int cmc_csd_msc_kmajor(const float* Xk, const float* Yk, int L, int F, int Ne, int Nm,
                                  int64_t pitch_x, int64_t pitch_y, float* coh, float* sxx, float* syy, float* sxy,
                                  void* ws, int64_t ws_bytes, int keep_operands, void* stream) {
    using namespace cmc;
    CMC_REQUIRE(coh, "cmc_csd_msc_kmajor: null pointer");
    int rc = kmajor_check(Xk, Yk, L, F, Ne, Nm, pitch_x, pitch_y, ws);
    if (rc) return rc;
    const CsdLayout y = csd_layout(L, F, Ne, Nm);
    const int64_t need = keep_operands ? y.total : y.off_ahi;
    if (ws_bytes < need) {
        set_error("cmc_csd_msc_kmajor: workspace %lld < %lld bytes", (long long)ws_bytes, (long long)need);
        return CMC_EWORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char* w = static_cast<unsigned char*>(ws);
    CUtensorMap mX, mY;
    if ((rc = make_rows_map(&mX, Xk, L, pitch_x, (int64_t)F * Ne))) return rc;
    if ((rc = make_rows_map(&mY, Yk, L, pitch_y, (int64_t)F * Nm))) return rc;
    FusedParams p{};
    p.F = F; p.MT = y.MT; p.NT = y.NT; p.Ne = Ne; p.Nm = Nm; p.KB = y.KP / kKBlock;
    p.coh = coh;
    p.sxy = reinterpret_cast<float2*>(sxy);
    p.pxx = reinterpret_cast<float*>(w + y.off_pxx);
    p.pyy = reinterpret_cast<float*>(w + y.off_pyy);
    p.sxx_out = sxx;
    p.syy_out = syy;
    p.total_tiles = (long long)F * y.MT * y.NT;
    const size_t smem = 1024 + (size_t)kFuStages * kFuStageBytes + (size_t)kTileM * kFuPitch * 4 + 2 * 128 * 4 +
                        sizeof(FusedBarriers) + 16;
    rc = ensure_smem_attr(reinterpret_cast<const void*>(csd_fused_kernel), smem);
    if (rc) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long grid = p.total_tiles < sms ? p.total_tiles : sms;
    csd_fused_kernel<<<(unsigned)grid, kFuThreads, smem, st>>>(mX, mY, p);
    CMC_CHECK_LAUNCH("csd_fused_kernel");
    if (keep_operands) return launch_planes(Xk, Yk, y, pitch_x, pitch_y, w, st);
    return CMC_OK;
}

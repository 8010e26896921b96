// K2 / K3: pooled cross-spectral density on the 5th-gen tensor cores (tcgen05.mma kind::tf32,
// TMA-staged K-major operands, FP32 accumulators in TMEM), fused with the auto-spectra
// normalisation into magnitude-squared coherence, and the shift-surrogate null that re-runs the
// same contraction on the cached operands with a rotated EMG segment index.
//
// Replaces signal_features.py:750-770 for averages over L segments / windows x tapers.
//
// Formulation.  Spectra are pre-whitened per (frequency, channel): Xh = X / sqrt(sum_l |X|^2), so
// coherence is |sum_l conj(Xh) Yh|^2 and the TF32 inputs are scale free.  The complex contraction is
// one real GEMM per frequency with K = (l, re/im) contiguous ("K-major" = complex64 memory order):
//     A rows 0..63   = Xh            (re, im interleaved along K)
//     A rows 64..127 = i * Xh        (-im, re)
//     B rows         = Yh
//     D[i][j] = Re S_ij,  D[64 + i][j] = Im S_ij            (M = 128, N = 64, K = 2L)
// The observed pass is error-compensated 3xTF32 (hi*hi + hi*lo + lo*hi with hi = tf32(x),
// lo = tf32(x - hi)), which keeps |dC| ~ 1e-6; it is HBM bound, so the extra MMAs are free.
// Shift surrogates read the B operand at K offset 2 * shift * group from a doubled row
// [Yh | Yh | 0...] - a TMA coordinate, no data movement - and use a single TF32 term.
//
// Pipeline per CTA (persistent over a contiguous tile range): warp 0 = TMA producer, warp 1 = MMA
// issuer (one thread), warp 2 = TMEM allocator, warps 4-7 = epilogue.  4-stage smem ring
// (full/empty mbarriers), 2 TMEM accumulators (tmem_full/tmem_empty) so the epilogue of tile t
// overlaps the MMAs of tile t + 1.
#include "common.cuh"
#include "tc_common.cuh"
#include "csd_layout.cuh"

namespace cmc {

using namespace tc;

constexpr int kStages = 4;
constexpr int kABytes = kTileM * kKBlock * 4;   // 16 KB
constexpr int kBBytes = kTileN * kKBlock * 4;   // 8 KB
constexpr int kStagePitch = kTileN + 1;         // padded staging row (floats)
constexpr int kGemmThreads = 256;

struct CsdParams {
    int F, MT, NT, Ne, Nm, KB, nterms, n_shift;
    const int32_t* shift_off;     // [n_shift] K offset (floats) into the doubled B rows, or null
    const uint32_t* shift_mult;   // [n_shift] surrogates that use this shift (0 = skip), or null
    float* coh;                   // EPI 0 out [F][Ne][Nm]
    float2* sxy;                  // EPI 0 optional out
    const float* pxx;             // [F][Ne] auto-spectra (for sxy)
    const float* pyy;             // [F][Nm]
    const float* coh_obs;         // EPI 1 in
    uint32_t* exceed;             // EPI 1 in/out [F][Ne][Nm]
    uint32_t* max_u;              // EPI 1 out [n_shift] float bits
    long long total_tiles;
};

struct TileCoord {
    int f, mt, nt, sh;
};
__device__ __forceinline__ TileCoord decode_tile(long long t, const CsdParams& p) {
    TileCoord c;
    c.sh = (int)(t % p.n_shift);
    long long r = t / p.n_shift;
    c.nt = (int)(r % p.NT);
    r /= p.NT;
    c.mt = (int)(r % p.MT);
    c.f = (int)(r / p.MT);
    return c;
}

struct __align__(8) GemmBarriers {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
csd_gemm_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                const __grid_constant__ CUtensorMap mBhi, const __grid_constant__ CUtensorMap mBlo,
                const __grid_constant__ CUtensorMap mBodd, const CsdParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = base;                                   // [kStages][16 KB]
    unsigned char* sB = base + kStages * kABytes;               // [kStages][8 KB]
    float* stage_tile = reinterpret_cast<float*>(sB + kStages * kBBytes);          // [128][65]
    uint32_t* cnt = reinterpret_cast<uint32_t*>(stage_tile + kTileM * kStagePitch);  // [64*64] (EPI 1)
    GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(cnt + (EPI == 1 ? 64 * 64 : 0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = p.total_tiles * blockIdx.x / gridDim.x;
    const long long t1 = p.total_tiles * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->tmem_full[a], 1);
            mbar_init(&bars->tmem_empty[a], 4);
        }
        fence_barrier_init();
        tma_prefetch_desc(&mAhi);
        tma_prefetch_desc(&mBhi);
    }
    if (warp == 2) {
        tmem_alloc(&bars->tmem_base, 2 * kTileN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    const int n_v = p.nterms * p.KB;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = t0; t < t1; ++t) {
                const TileCoord c = decode_tile(t, p);
                if (p.shift_mult && p.shift_mult[c.sh] == 0) continue;
                // TMA needs a 16-byte aligned box start: offsets = 2 (mod 4) floats read the copy of
                // the B rows that is pre-shifted by one complex element
                int off = p.shift_off ? p.shift_off[c.sh] : 0;
                const bool odd = (off & 2) != 0;
                off -= odd ? 2 : 0;
                const int arow = (c.f * p.MT + c.mt) * kTileM;
                const int brow = (c.f * p.NT + c.nt) * kTileN;
                for (int v = 0; v < n_v; ++v) {
                    const int term = v / p.KB, kb = v - term * p.KB;
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars->full[stage], kABytes + kBBytes);
                    tma_load_2d(sA + stage * kABytes, term == 2 ? &mAlo : &mAhi, &bars->full[stage], kb * kKBlock, arow);
                    tma_load_2d(sB + stage * kBBytes, odd ? &mBodd : (term == 1 ? &mBlo : &mBhi), &bars->full[stage],
                                off + kb * kKBlock, brow);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(kTileM, kTileN);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (long long t = t0; t < t1; ++t) {
                const TileCoord c = decode_tile(t, p);
                if (p.shift_mult && p.shift_mult[c.sh] == 0) continue;
                const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
                mbar_wait(&bars->tmem_empty[acc], accphase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * kTileN;
                for (int v = 0; v < n_v; ++v) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + stage * kABytes);
                    const uint32_t b0 = smem_u32(sB + stage * kBBytes);
#pragma unroll
                    for (int k = 0; k < kKBlock / 8; ++k)
                        umma_tf32(d, make_smem_desc_k_sw128(a0 + k * 32), make_smem_desc_k_sw128(b0 + k * 32), idesc,
                                  (v | k) != 0 ? 1u : 0u);
                    umma_commit(&bars->empty[stage]);      // frees the smem slot when these MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars->tmem_full[acc]);        // accumulator ready for the epilogue
                ++it;
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int q = warp - 4;                  // TMEM lane quadrant of this warp
        const int te = threadIdx.x - 128;        // 0..127
        if (EPI == 1)
            for (int n = 0; n < 32; ++n) cnt[te + 128 * n] = 0;
        long long key = -1;
        int kf = 0, kmt = 0, knt = 0;
        uint32_t it = 0;
        for (long long t = t0; t < t1; ++t) {
            const TileCoord c = decode_tile(t, p);
            const uint32_t mult = p.shift_mult ? p.shift_mult[c.sh] : 1u;
            if (mult == 0) continue;
            if (EPI == 1) {
                const long long k2 = ((long long)c.f * p.MT + c.mt) * p.NT + c.nt;
                if (k2 != key) {
                    if (key >= 0) {
                        for (int n = 0; n < 32; ++n) {
                            const int idx = te + 128 * n, i = kmt * 64 + (idx >> 6), j = knt * 64 + (idx & 63);
                            if (cnt[idx] && i < p.Ne && j < p.Nm)
                                atomicAdd(&p.exceed[((long long)kf * p.Ne + i) * p.Nm + j], cnt[idx]);
                            cnt[idx] = 0;
                        }
                    }
                    key = k2; kf = c.f; kmt = c.mt; knt = c.nt;
                }
            }
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
            float* row = stage_tile + (q * 32 + lane) * kStagePitch;
#pragma unroll
            for (int cidx = 0; cidx < 32; ++cidx) {
                row[cidx] = __uint_as_float(r0[cidx]);
                row[32 + cidx] = __uint_as_float(r1[cidx]);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            float vmax = 0.f;
#pragma unroll 4
            for (int n = 0; n < 32; ++n) {
                const int idx = te + 128 * n;
                const int il = idx >> 6, jl = idx & 63;
                const int i = c.mt * 64 + il, j = c.nt * 64 + jl;
                if (i < p.Ne && j < p.Nm) {
                    const float re = stage_tile[il * kStagePitch + jl];
                    const float im = stage_tile[(64 + il) * kStagePitch + jl];
                    const float cval = fminf(re * re + im * im, 1.0f);
                    const long long o = ((long long)c.f * p.Ne + i) * p.Nm + j;
                    if (EPI == 0) {
                        p.coh[o] = cval;
                        if (p.sxy) {
                            const float s = sqrtf(p.pxx[c.f * p.Ne + i]) * sqrtf(p.pyy[c.f * p.Nm + j]);
                            p.sxy[o] = make_float2(re * s, im * s);
                        }
                    } else {
                        if (cval >= __ldg(p.coh_obs + o)) cnt[idx] += mult;
                        vmax = fmaxf(vmax, cval);
                    }
                }
            }
            if (EPI == 1) {
                const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax));
                if (lane == 0) atomicMax(&p.max_u[c.sh], m);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            ++it;
        }
        if (EPI == 1 && key >= 0) {
            for (int n = 0; n < 32; ++n) {
                const int idx = te + 128 * n, i = kmt * 64 + (idx >> 6), j = knt * 64 + (idx & 63);
                if (cnt[idx] && i < p.Ne && j < p.Nm)
                    atomicAdd(&p.exceed[((long long)kf * p.Ne + i) * p.Nm + j], cnt[idx]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 2 * kTileN);
}

// ------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------
// P[f][c] = sum_l |S[l][f][c]|^2, fixed summation order (deterministic)
__global__ void __launch_bounds__(256)
power_kernel(const float2* __restrict__ S, int L, int F, int C, int64_t ld, float* __restrict__ P) {
    __shared__ float part[8][33];
    const int f = blockIdx.x, c = blockIdx.y * 32 + threadIdx.x, ly = threadIdx.y;
    float acc = 0.f;
    if (c < C)
        for (int l = ly; l < L; l += 8) {
            const float2 v = __ldg(S + ((int64_t)l * F + f) * ld + c);
            acc += v.x * v.x + v.y * v.y;
        }
    part[ly][threadIdx.x] = acc;
    __syncthreads();
    if (ly == 0 && c < C) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
        P[(int64_t)f * C + c] = t;
    }
}

// Transposes S[l][f][c] into whitened K-major TF32 operand rows (hi and lo planes).
//   MODE 0 (A operand): rows_per_f = MT * 128; row (mt*128 + r): r < 64 -> channel mt*64 + r (Xh),
//                       r >= 64 -> i * Xh of channel mt*64 + r - 64; columns k >= 2L are zero.
//   MODE 1 (B operand): rows_per_f = NT * 64; row = channel; columns [0,2L) and [2L,4L) both hold Yh,
//                       columns >= 4L are zero; l_shift = 1 writes the same rows advanced by one
//                       complex element (16-byte aligned access to odd segment shifts).
// grid (F, rows_per_f / 32, ceil(row_len / 64)); block (32, 8)
template <int MODE>
__global__ void __launch_bounds__(256)
pack_kernel(const float2* __restrict__ S, int L, int F, int C, int64_t ld, const float* __restrict__ P,
            int rows_per_f, int row_len, int l_shift, float* __restrict__ hi, float* __restrict__ lo) {
    __shared__ float2 tile[32][33];       // [l][row]
    const int f = blockIdx.x, r0 = blockIdx.y * 32, l0 = blockIdx.z * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;
    // load: thread (tx = row, ty + 8n = l)
    {
        const int r = r0 + tx;
        int ch;
        bool rot = false;
        if (MODE == 0) {
            const int mt = r >> 7, rr = r & 127;
            rot = rr >= 64;
            ch = mt * 64 + (rr & 63);
        } else {
            ch = r;
        }
        float scale = 0.f;
        if (ch < C) {
            const float pw = P[(int64_t)f * C + ch];
            scale = pw > 0.f ? rsqrtf(pw) : 0.f;
        }
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int lv = l0 + ty + 8 * n + l_shift;   // virtual l (column pair index)
            int l = -1;
            if (lv < L) l = lv;
            else if (MODE == 1 && lv < 2 * L) l = lv - L;
            float2 v = make_float2(0.f, 0.f);
            if (l >= 0 && ch < C) {
                const float2 s = __ldg(S + ((int64_t)l * F + f) * ld + ch);
                v = rot ? make_float2(-s.y * scale, s.x * scale) : make_float2(s.x * scale, s.y * scale);
            }
            tile[ty + 8 * n][tx] = v;
        }
    }
    __syncthreads();
    // store: thread (tx = l, ty + 8n = row): 32 lanes write 256 contiguous bytes of one operand row
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const int r = r0 + ty + 8 * n;
        const int k = 2 * (l0 + tx);
        if (k < row_len) {
            const float2 v = tile[tx][ty + 8 * n];
            const float2 h = make_float2(to_tf32(v.x), to_tf32(v.y));
            const int64_t o = ((int64_t)f * rows_per_f + r) * row_len + k;
            *reinterpret_cast<float2*>(hi + o) = h;
            if (lo) *reinterpret_cast<float2*>(lo + o) = make_float2(to_tf32(v.x - h.x), to_tf32(v.y - h.y));
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static size_t gemm_smem_bytes(int epi) {
    return 1024 + kStages * (kABytes + kBBytes) + sizeof(float) * kTileM * kStagePitch +
           (epi == 1 ? 64 * 64 * 4 : 0) + sizeof(GemmBarriers) + 16;
}

template <int EPI>
static int launch_gemm(const CsdLayout& y, unsigned char* ws, CsdParams p, cudaStream_t st) {
    CUtensorMap mAhi, mAlo, mBhi, mBlo, mBodd;
    int rc;
    if ((rc = make_operand_map(&mAhi, reinterpret_cast<float*>(ws + y.off_ahi), y.KP, (int64_t)y.F * y.MT * kTileM, kTileM))) return rc;
    if ((rc = make_operand_map(&mAlo, reinterpret_cast<float*>(ws + y.off_alo), y.KP, (int64_t)y.F * y.MT * kTileM, kTileM))) return rc;
    if ((rc = make_operand_map(&mBhi, reinterpret_cast<float*>(ws + y.off_bhi), y.LB, (int64_t)y.F * y.NT * kTileN, kTileN))) return rc;
    if ((rc = make_operand_map(&mBlo, reinterpret_cast<float*>(ws + y.off_blo), y.LB, (int64_t)y.F * y.NT * kTileN, kTileN))) return rc;
    if ((rc = make_operand_map(&mBodd, reinterpret_cast<float*>(ws + y.off_bodd), y.LB, (int64_t)y.F * y.NT * kTileN, kTileN))) return rc;
    const size_t smem = gemm_smem_bytes(EPI);
    rc = ensure_smem_attr(reinterpret_cast<const void*>(csd_gemm_kernel<EPI>), smem);
    if (rc) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long grid = p.total_tiles < sms ? p.total_tiles : sms;
    csd_gemm_kernel<EPI><<<(unsigned)grid, kGemmThreads, smem, st>>>(mAhi, mAlo, mBhi, mBlo, mBodd, p);
    CMC_CHECK_LAUNCH("csd_gemm_kernel");
    return CMC_OK;
}

// shift-surrogate bookkeeping
__global__ void shift_hist_kernel(const int32_t* __restrict__ shifts, int64_t n, int n_pos, int group,
                                  uint32_t* __restrict__ mult, int32_t* __restrict__ off, uint32_t* __restrict__ max_u) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_pos) off[i] = 2 * (int)i * group;
    if (i < n) {
        int s = shifts[i] % n_pos;
        if (s < 0) s += n_pos;
        atomicAdd(&mult[s], 1u);
    }
    (void)max_u;
}
__global__ void shift_gather_kernel(const int32_t* __restrict__ shifts, int64_t n, int n_pos,
                                    const uint32_t* __restrict__ max_u, float* __restrict__ max_stat) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        int s = shifts[i] % n_pos;
        if (s < 0) s += n_pos;
        max_stat[i] = __uint_as_float(max_u[s]);
    }
}

int64_t phase_workspace_bytes(int L, int F, int Ne, int Nm, int64_t n_surr);
int phase_surrogate_null(const void* ws, int L, int F, int Ne, int Nm, uint64_t seed, int64_t s_begin, int64_t s_end,
                         const float* coh_obs, uint32_t* exceed, float* max_stat, void* ws2, int64_t ws2_bytes,
                         cudaStream_t st);

}  // namespace cmc

extern "C" int64_t cmc_csd_workspace_bytes(int L, int F, int Ne, int Nm) {
    if (L < 1 || F < 1 || Ne < 1 || Nm < 1) return CMC_EINVAL;
    return cmc::csd_layout(L, F, Ne, Nm).total;
}

extern "C" int cmc_csd_msc(const float* X, const float* Y, int L, int F, int Ne, int Nm, int64_t ldx, int64_t ldy,
                           float* coh, float* sxx, float* syy, float* sxy, void* ws, int64_t ws_bytes,
                           void* stream) {
    using namespace cmc;
    CMC_REQUIRE(X && Y && coh && ws, "cmc_csd_msc: null pointer");
    CMC_REQUIRE(L >= 1 && F >= 1 && Ne >= 1 && Nm >= 1 && ldx >= Ne && ldy >= Nm, "cmc_csd_msc: bad shape");
    const CsdLayout y = csd_layout(L, F, Ne, Nm);
    if (ws_bytes < y.total) {
        set_error("cmc_csd_msc: workspace %lld < %lld bytes", (long long)ws_bytes, (long long)y.total);
        return CMC_EWORKSPACE;
    }
    CMC_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "cmc_csd_msc: workspace must be 256-byte aligned");
    CMC_REQUIRE((int64_t)F * y.MT * kTileM < (1ll << 31) && (int64_t)F * y.NT * kTileN < (1ll << 31),
                "cmc_csd_msc: operand too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char* w = static_cast<unsigned char*>(ws);
    float* pxx = reinterpret_cast<float*>(w + y.off_pxx);
    float* pyy = reinterpret_cast<float*>(w + y.off_pyy);
    const float2* Xc = reinterpret_cast<const float2*>(X);
    const float2* Yc = reinterpret_cast<const float2*>(Y);
    power_kernel<<<dim3(F, (Ne + 31) / 32), dim3(32, 8), 0, st>>>(Xc, L, F, Ne, ldx, pxx);
    CMC_CHECK_LAUNCH("power_kernel(X)");
    power_kernel<<<dim3(F, (Nm + 31) / 32), dim3(32, 8), 0, st>>>(Yc, L, F, Nm, ldy, pyy);
    CMC_CHECK_LAUNCH("power_kernel(Y)");
    pack_kernel<0><<<dim3(F, y.MT * kTileM / 32, (y.KP + 63) / 64), dim3(32, 8), 0, st>>>(
        Xc, L, F, Ne, ldx, pxx, y.MT * kTileM, y.KP, 0, reinterpret_cast<float*>(w + y.off_ahi),
        reinterpret_cast<float*>(w + y.off_alo));
    CMC_CHECK_LAUNCH("pack_kernel<A>");
    pack_kernel<1><<<dim3(F, y.NT * kTileN / 32, (y.LB + 63) / 64), dim3(32, 8), 0, st>>>(
        Yc, L, F, Nm, ldy, pyy, y.NT * kTileN, y.LB, 0, reinterpret_cast<float*>(w + y.off_bhi),
        reinterpret_cast<float*>(w + y.off_blo));
    CMC_CHECK_LAUNCH("pack_kernel<B>");
    pack_kernel<1><<<dim3(F, y.NT * kTileN / 32, (y.LB + 63) / 64), dim3(32, 8), 0, st>>>(
        Yc, L, F, Nm, ldy, pyy, y.NT * kTileN, y.LB, 1, reinterpret_cast<float*>(w + y.off_bodd), nullptr);
    CMC_CHECK_LAUNCH("pack_kernel<B odd>");
    if (sxx) {
        int rc = check_cuda(cudaMemcpyAsync(sxx, pxx, sizeof(float) * F * Ne, cudaMemcpyDeviceToDevice, st), "copy sxx");
        if (rc) return rc;
    }
    if (syy) {
        int rc = check_cuda(cudaMemcpyAsync(syy, pyy, sizeof(float) * F * Nm, cudaMemcpyDeviceToDevice, st), "copy syy");
        if (rc) return rc;
    }
    CsdParams p{};
    p.F = F; p.MT = y.MT; p.NT = y.NT; p.Ne = Ne; p.Nm = Nm; p.KB = y.KP / kKBlock; p.nterms = 3; p.n_shift = 1;
    p.coh = coh; p.sxy = reinterpret_cast<float2*>(sxy); p.pxx = pxx; p.pyy = pyy;
    p.total_tiles = (long long)F * y.MT * y.NT;
    return launch_gemm<0>(y, w, p, st);
}

extern "C" int64_t cmc_surrogate_workspace_bytes(int L, int F, int Ne, int Nm, int mode, int64_t n_surr) {
    (void)F; (void)Ne; (void)Nm; (void)n_surr;
    if (L < 1 || F < 1 || Ne < 1 || Nm < 1 || n_surr < 0) return CMC_EINVAL;
    if (mode == CMC_SURR_SHIFT) return (int64_t)L * 12 + 256;
    if (mode == CMC_SURR_PHASE) return cmc::phase_workspace_bytes(L, F, Ne, Nm, n_surr);
    return CMC_EINVAL;
}

extern "C" int cmc_surrogate_null(const void* ws, int L, int F, int Ne, int Nm, int mode, int group,
                                  const int32_t* shifts, uint64_t seed, int64_t s_begin, int64_t s_end,
                                  const float* coh_obs, uint32_t* exceed, float* max_stat, void* ws2,
                                  int64_t ws2_bytes, void* stream) {
    using namespace cmc;
    CMC_REQUIRE(ws && coh_obs && exceed && max_stat && ws2, "cmc_surrogate_null: null pointer");
    CMC_REQUIRE(s_end >= s_begin, "cmc_surrogate_null: bad surrogate range");
    const int64_t n = s_end - s_begin;
    if (n == 0) return CMC_OK;
    if (mode == CMC_SURR_PHASE)
        return phase_surrogate_null(ws, L, F, Ne, Nm, seed, s_begin, s_end, coh_obs, exceed, max_stat, ws2, ws2_bytes,
                                    static_cast<cudaStream_t>(stream));
    CMC_REQUIRE(mode == CMC_SURR_SHIFT, "cmc_surrogate_null: unknown mode %d", mode);
    CMC_REQUIRE(shifts, "cmc_surrogate_null: shift mode needs a shift table");
    CMC_REQUIRE(group >= 1 && L % group == 0, "cmc_surrogate_null: group must divide L");
    const int n_pos = L / group;
    if (ws2_bytes < cmc_surrogate_workspace_bytes(L, F, Ne, Nm, mode, n)) {
        set_error("cmc_surrogate_null: workspace too small");
        return CMC_EWORKSPACE;
    }
    const CsdLayout y = csd_layout(L, F, Ne, Nm);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint32_t* mult = static_cast<uint32_t*>(ws2);
    uint32_t* max_u = mult + L;
    int32_t* off = reinterpret_cast<int32_t*>(max_u + L);
    int rc = check_cuda(cudaMemsetAsync(ws2, 0, (size_t)L * 8, st), "memset(shift tables)");
    if (rc) return rc;
    const int64_t nthreads = n > n_pos ? n : n_pos;
    shift_hist_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(shifts, n, n_pos, group, mult, off, max_u);
    CMC_CHECK_LAUNCH("shift_hist_kernel");
    CsdParams p{};
    p.F = F; p.MT = y.MT; p.NT = y.NT; p.Ne = Ne; p.Nm = Nm; p.KB = y.KP / kKBlock; p.nterms = 1; p.n_shift = n_pos;
    p.shift_off = off; p.shift_mult = mult; p.coh_obs = coh_obs; p.exceed = exceed; p.max_u = max_u;
    p.total_tiles = (long long)F * y.MT * y.NT * n_pos;
    rc = launch_gemm<1>(y, const_cast<unsigned char*>(static_cast<const unsigned char*>(ws)), p, st);
    if (rc) return rc;
    shift_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(shifts, n, n_pos, max_u, max_stat);
    CMC_CHECK_LAUNCH("shift_gather_kernel");
    return CMC_OK;
}

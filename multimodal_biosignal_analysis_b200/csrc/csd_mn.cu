// K2, direct form: pooled cross-spectral density + coherence straight from the spectra as K1 writes them
// ([segment][bin][channel] complex64) - no pack pass and no re-layout.
//
// For one frequency the spectra of 64 channels are, per segment l, 128 contiguous floats
// (re, im interleaved over channels): an MN-major UMMA operand with M = (channel, re/im) and K = l.
// One GEMM per (frequency, 64 x 64 channel tile)
//     D[(i,ci)][(j,cj)] = sum_l X[l][i].ci * Y[l][j].cj              (M = N = 128, K = L)
// holds the four real products of every pair; the epilogue folds the 2 x 2 blocks,
//     Re S_ij = D[(i,0)][(j,0)] + D[(i,1)][(j,1)],   Im S_ij = D[(i,0)][(j,1)] - D[(i,1)][(j,0)],
// with two shuffles between the two TMEM lanes of a channel, then normalises with the auto-spectra.
// TMA stages 32 segments x 32 floats boxes (128-byte swizzle with 32-byte base, the only layout UMMA accepts for
// MN-major TF32 operands; segments past L and channels past the last read as zeros); eight converter warps round
// the tile to TF32 in place, write the lo plane (3xTF32: lo*hi + hi*lo + hi*hi) and accumulate Pxx / Pyy in a
// fixed order.  HBM traffic = the spectra, once.
//
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue,
// warps 8-15 converters.  Stage ring: full (TMA bytes) -> conv (256 converter arrivals) -> MMAs -> empty.
#include "common.cuh"
#include "tc_common.cuh"
#include "csd_layout.cuh"

namespace cmc {

using namespace tc;

constexpr int kMnThreads = 512;
constexpr int kMnConvThreads = 256;                        // warps 8-15
constexpr int kMnStages = 3;
constexpr int kMnKBlock = 32;                             // segments per stage
constexpr int kMnAtomBytes = kMnKBlock * 128;             // 4 KB: 32 k-rows of one 32-float M/N atom
constexpr int kMnOpBytes = 4 * kMnAtomBytes;              // 16 KB: 128 floats (64 channels) x 32 segments
constexpr int kMnStageBytes = 4 * kMnOpBytes;             // A_hi, A_lo, B_hi, B_lo = 64 KB
constexpr int kMnPitch = 65;                              // padded staging row (floats)

struct MnParams {
    int F, MT, NT, Ne, Nm, KB;
    float* coh;           // [F][Ne][Nm]
    float2* sxy;          // optional
    float* pxx;           // [F][Ne] workspace copy
    float* pyy;           // [F][Nm]
    float* sxx_out;       // optional user outputs
    float* syy_out;
    long long total_tiles;
};

struct __align__(8) MnBarriers {
    uint64_t full[kMnStages];
    uint64_t conv[kMnStages];
    uint64_t empty[kMnStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint64_t pw_full[2];
    uint64_t pw_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

#ifdef CMC_K2_TRACE
// instrumented build only: globaltimer stamps of CTA 0 (ns), slot = event id
__device__ unsigned long long g_k2_trace[64];
__device__ __forceinline__ void k2_stamp(int slot) {
    if (blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_k2_trace[slot] = t;
    }
}
#define K2_STAMP(slot) k2_stamp(slot)
#else
#define K2_STAMP(slot) do { } while (0)
#endif

__device__ __forceinline__ float4 mn_lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mn_sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Epilogue of one chunk of 32 accumulator columns (16 EMG channels) for the accumulator row te = (EEG channel, re/im)
// of this thread: fold the 2 x 2 blocks, normalise, park the coherence in the staging tile, optionally store Sxy.
__device__ __forceinline__ void mn_fold_chunk(const MnParams& p, uint32_t taddr, int ch, const float* pwa, float* stage_tile,
                                              int te, int mt, int nt, int f) {
    const int il = te >> 1, odd = te & 1;
    const float sx = pwa[128 + il];
    const int i = mt * 64 + il;
    const uint32_t sy_addr = smem_u32(pwa + 192);
    uint32_t v[32];
    tmem_ld_32x32(taddr + ch * 32, v);
    float sy[16];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float4 w = mn_lds128(sy_addr + (uint32_t)(ch * 64 + g * 16));
        sy[4 * g] = w.x; sy[4 * g + 1] = w.y; sy[4 * g + 2] = w.z; sy[4 * g + 3] = w.w;
    }
    tmem_ld_wait();
    // Both lanes of a channel pair fetch the partner's row (two independent shuffles per column pair);
    // the even lane then finishes the even EMG channels, the odd lane the odd ones.
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
        const float e = __uint_as_float(v[2 * jj]), o = __uint_as_float(v[2 * jj + 1]);
        const float ep = __shfl_xor_sync(0xffffffffu, e, 1);
        const float op = __shfl_xor_sync(0xffffffffu, o, 1);
        if ((jj & 1) == odd) {
            // rows (i, re) / (i, im): Re S = xr yr + xi yi, Im S = xr yi - xi yr
            const float re = odd ? ep + o : e + op;
            const float im = odd ? op - e : o - ep;
            const int jl = ch * 16 + jj;
            // |S|^2 / (Pxx Pyy) as |S / sqrt(Pxx) / sqrt(Pyy)|^2: no overflow, silent channels give 0
            const float sc = sx * sy[jj];
            const float ar = re * sc, ai = im * sc;
            stage_tile[il * kMnPitch + jl] = fminf(ar * ar + ai * ai, 1.0f);
            if (p.sxy) {
                const int j = nt * 64 + jl;
                if (i < p.Ne && j < p.Nm) p.sxy[((long long)f * p.Ne + i) * p.Nm + j] = make_float2(re, im);
            }
        }
    }
}
// coalesced copy-out of the 64 x 64 coherence tile by n_thr threads
__device__ __forceinline__ void mn_copy_out(const MnParams& p, const float* stage_tile, int tid, int n_thr, int mt, int nt,
                                            int f) {
#pragma unroll 4
    for (int idx = tid; idx < 64 * 64; idx += n_thr) {
        const int r = idx >> 6, c = idx & 63;
        const int gi = mt * 64 + r, gj = nt * 64 + c;
        if (gi < p.Ne && gj < p.Nm) p.coh[((long long)f * p.Ne + gi) * p.Nm + gj] = stage_tile[r * kMnPitch + c];
    }
}

__global__ void __launch_bounds__(kMnThreads, 1)
csd_mn_kernel(const __grid_constant__ CUtensorMap mX, const __grid_constant__ CUtensorMap mY, const MnParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* sS = base;                                                         // [kMnStages][64 KB]
    float* stage_tile = reinterpret_cast<float*>(sS + kMnStages * kMnStageBytes);     // [64][65] coherence tile
    float* pw = stage_tile + 64 * kMnPitch;            // [2][256]: Pxx, Pyy, 1/sqrt(Pxx), 1/sqrt(Pyy) of the tile's channels
    float* pw_part = pw + 2 * 256;                     // [128] partial auto-spectra of the second converter warp per atom
    MnBarriers* bars = reinterpret_cast<MnBarriers*>(pw_part + 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) K2_STAMP(0);
    const long long t0 = p.total_tiles * blockIdx.x / gridDim.x;
    const long long t1 = p.total_tiles * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMnStages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->conv[s], kMnConvThreads);
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->tmem_full[a], 1);
            mbar_init(&bars->tmem_empty[a], 4);
            mbar_init(&bars->pw_full[a], kMnConvThreads);
            mbar_init(&bars->pw_empty[a], 128);
        }
        fence_barrier_init();
        tma_prefetch_desc(&mX);
        tma_prefetch_desc(&mY);
    }
    if (warp == 2) {
        tmem_alloc(&bars->tmem_base, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    if (threadIdx.x == 0) K2_STAMP(1);

    if (warp == 0) {
        // ===================== TMA producer: lane 0 owns the barriers, lanes 0-7 each issue one box =====================
        int stage = 0;
        uint32_t phase = 0;
        const int a = lane & 3;                      // 32-float atom of the 128-float operand row
        const bool is_y = (lane & 4) != 0;
        for (long long t = t0; t < t1; ++t) {
            const int nt = (int)(t % p.NT);
            const long long r = t / p.NT;
            const int mt = (int)(r % p.MT), f = (int)(r / p.MT);
            for (int kb = 0; kb < p.KB; ++kb) {
                unsigned char* st = sS + stage * kMnStageBytes;
                if (lane == 0) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    if (t == t0) K2_STAMP(2 + kb);
                    mbar_arrive_expect_tx(&bars->full[stage], 2 * kMnOpBytes);
                }
                __syncwarp();
                if (lane < 8)
                    tma_load_3d(st + (is_y ? 2 * kMnOpBytes : 0) + a * kMnAtomBytes, is_y ? &mY : &mX, &bars->full[stage],
                                (is_y ? nt : mt) * 128 + a * 32, f, kb * kMnKBlock);
                if (++stage == kMnStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32_mn(128, 128);
            int stage = 0;
            uint32_t phase = 0, it = 0;
            for (long long t = t0; t < t1; ++t) {
                const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
                mbar_wait(&bars->tmem_empty[acc], accphase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * 128;
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(&bars->conv[stage], phase);
                    if (t == t0) K2_STAMP(30 + kb);
                    tc_fence_after();
                    const uint32_t ahi = smem_u32(sS + stage * kMnStageBytes), alo = ahi + kMnOpBytes;
                    const uint32_t bhi = ahi + 2 * kMnOpBytes, blo = bhi + kMnOpBytes;
#pragma unroll
                    for (int k = 0; k < kMnKBlock / 8; ++k) {            // one 8-segment k-group (1 KB) per MMA
                        const uint64_t dah = make_smem_desc_mn_sw128_b32(ahi + k * 1024, kMnAtomBytes);
                        const uint64_t dal = make_smem_desc_mn_sw128_b32(alo + k * 1024, kMnAtomBytes);
                        const uint64_t dbh = make_smem_desc_mn_sw128_b32(bhi + k * 1024, kMnAtomBytes);
                        const uint64_t dbl = make_smem_desc_mn_sw128_b32(blo + k * 1024, kMnAtomBytes);
                        umma_tf32(d, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);      // small terms first
                        umma_tf32(d, dah, dbl, idesc, 1u);
                        umma_tf32(d, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&bars->empty[stage]);
                    if (++stage == kMnStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars->tmem_full[acc]);
                ++it;
            }
        }
    } else if (warp >= 8) {
        // ===================== converters: TF32 hi/lo split and auto-spectra =====================
        // thread -> atom a (32 floats = 16 channels), 16-byte chunk cc (2 channels), rows rg, rg + 8, rg + 16, rg + 24
        const int tc_ = threadIdx.x - 256;           // 0..255
        const int a = tc_ >> 6, cc = tc_ & 7, rg = (tc_ >> 3) & 7;
        int stage = 0;
        uint32_t phase = 0, it = 0;
        for (long long t = t0; t < t1; ++t) {
            float px0 = 0.f, px1 = 0.f, py0 = 0.f, py1 = 0.f;
            for (int kb = 0; kb < p.KB; ++kb) {
                mbar_wait(&bars->full[stage], phase);
                if (t == t0 && tc_ == 0) K2_STAMP(10 + kb);
                const uint32_t ahi = smem_u32(sS + stage * kMnStageBytes) + a * kMnAtomBytes;
                const uint32_t alo = ahi + kMnOpBytes, bhi = ahi + 2 * kMnOpBytes, blo = bhi + kMnOpBytes;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = rg + 8 * j;                                      // segment row of the k-block
                    // 128-byte swizzle with 32-byte base: 32-byte chunk index XOR (row & 3)
                    const uint32_t off = (uint32_t)(k * 128 + ((((cc >> 1) ^ (k & 3)) << 5) | ((cc & 1) << 4)));
                    const float4 x = mn_lds128(ahi + off);
                    const float4 y = mn_lds128(bhi + off);
                    px0 += x.x * x.x + x.y * x.y;
                    px1 += x.z * x.z + x.w * x.w;
                    py0 += y.x * y.x + y.y * y.y;
                    py1 += y.z * y.z + y.w * y.w;
                    const float4 xh = make_float4(to_tf32(x.x), to_tf32(x.y), to_tf32(x.z), to_tf32(x.w));
                    const float4 yh = make_float4(to_tf32(y.x), to_tf32(y.y), to_tf32(y.z), to_tf32(y.w));
                    mn_sts128(ahi + off, xh);
                    mn_sts128(alo + off, make_float4(to_tf32(x.x - xh.x), to_tf32(x.y - xh.y), to_tf32(x.z - xh.z),
                                                     to_tf32(x.w - xh.w)));
                    mn_sts128(bhi + off, yh);
                    mn_sts128(blo + off, make_float4(to_tf32(y.x - yh.x), to_tf32(y.y - yh.y), to_tf32(y.z - yh.z),
                                                     to_tf32(y.w - yh.w)));
                }
                fence_proxy_async();                 // generic-proxy writes -> visible to the MMA's async-proxy reads
                mbar_arrive(&bars->conv[stage]);
                if (t == t0 && tc_ == 0) K2_STAMP(20 + kb);
                if (++stage == kMnStages) { stage = 0; phase ^= 1; }
            }
            // fold the eight row groups: four inside the warp (lanes differing in bits 3 and 4), then the two warps
            // of an atom through shared memory - a fixed order
            px0 += __shfl_xor_sync(0xffffffffu, px0, 8);  px0 += __shfl_xor_sync(0xffffffffu, px0, 16);
            px1 += __shfl_xor_sync(0xffffffffu, px1, 8);  px1 += __shfl_xor_sync(0xffffffffu, px1, 16);
            py0 += __shfl_xor_sync(0xffffffffu, py0, 8);  py0 += __shfl_xor_sync(0xffffffffu, py0, 16);
            py1 += __shfl_xor_sync(0xffffffffu, py1, 8);  py1 += __shfl_xor_sync(0xffffffffu, py1, 16);
            const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
            mbar_wait(&bars->pw_empty[acc], accphase ^ 1);
            const int ch = a * 16 + cc * 2;
            const bool upper = (rg & 4) != 0;        // second warp of the atom
            if ((rg & 3) == 0 && upper) {
                pw_part[ch] = px0;  pw_part[ch + 1] = px1;
                pw_part[64 + ch] = py0;  pw_part[64 + ch + 1] = py1;
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if ((rg & 3) == 0 && !upper) {
                float* o = pw + acc * 256;
                px0 += pw_part[ch];  px1 += pw_part[ch + 1];
                py0 += pw_part[64 + ch];  py1 += pw_part[64 + ch + 1];
                o[ch] = px0;  o[ch + 1] = px1;  o[64 + ch] = py0;  o[64 + ch + 1] = py1;
                // normalisation factors 1 / sqrt(P), 0 for a silent channel
                o[128 + ch] = px0 > 0.f ? rsqrtf(px0) : 0.f;
                o[128 + ch + 1] = px1 > 0.f ? rsqrtf(px1) : 0.f;
                o[192 + ch] = py0 > 0.f ? rsqrtf(py0) : 0.f;
                o[192 + ch + 1] = py1 > 0.f ? rsqrtf(py1) : 0.f;
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");        // pw_part is free again
            mbar_arrive(&bars->pw_full[acc]);
            ++it;
        }
        if (t1 > t0) {
            // help with the epilogue of the last tile: warps 8-11 fold column chunk 2, warps 12-15 chunk 3 (a warp reads
            // the TMEM lane quadrant warp % 4), then everybody copies the tile out
            const long long t = t1 - 1;
            const int nt = (int)(t % p.NT);
            const long long rr = t / p.NT;
            const int mt = (int)(rr % p.MT), f = (int)(rr / p.MT);
            const uint32_t acc = (it - 1) & 1, accphase = ((it - 1) >> 1) & 1;
            mbar_wait(&bars->pw_full[acc], accphase);
            mbar_wait(&bars->tmem_full[acc], accphase);
            asm volatile("bar.sync 4, 384;" ::: "memory");      // the epilogue warps are done with the previous tile
            tc_fence_after();
            const int q = warp & 3, te = q * 32 + lane;
            const uint32_t taddr = tmem_base + acc * 128 + (static_cast<uint32_t>(q * 32) << 16);
            mn_fold_chunk(p, taddr, warp < 12 ? 2 : 3, pw + acc * 256, stage_tile, te, mt, nt, f);
            tc_fence_before();
            asm volatile("bar.sync 3, 384;" ::: "memory");
            mn_copy_out(p, stage_tile, 128 + tc_, 384, mt, nt, f);
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int q = warp - 4;                  // TMEM lane quadrant of this warp
        const int te = threadIdx.x - 128;        // 0..127 = accumulator row (i, ci)
        uint32_t it = 0;
        for (long long t = t0; t < t1; ++t) {
            const int nt = (int)(t % p.NT);
            const long long rr = t / p.NT;
            const int mt = (int)(rr % p.MT), f = (int)(rr / p.MT);
            const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
            mbar_wait(&bars->pw_full[acc], accphase);
            mbar_wait(&bars->tmem_full[acc], accphase);
            if (t == t0 && te == 0) K2_STAMP(40);
            tc_fence_after();
            const float* pwa = pw + acc * 256;
            const uint32_t taddr = tmem_base + acc * 128 + (static_cast<uint32_t>(q * 32) << 16);
            // On the last tile of this CTA the converter warps are idle: they take the upper two column chunks and a
            // share of the copy-out (with one tile per CTA, config 2, that shortens the launch by the fold time).
            const bool last = t == t1 - 1;
            if (last) asm volatile("bar.arrive 4, 384;" ::: "memory");    // the staging tile is free for the helpers
#pragma unroll 1
            for (int ch = 0; ch < (last ? 2 : 4); ++ch) mn_fold_chunk(p, taddr, ch, pwa, stage_tile, te, mt, nt, f);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);   // accumulator may be overwritten
            if (t == t0 && te == 0) K2_STAMP(41);
            if (last) asm volatile("bar.sync 3, 384;" ::: "memory");      // with the helping converter warps
            else asm volatile("bar.sync 1, 128;" ::: "memory");
            {   // auto-spectra outputs (every pair tile of a row block carries the same values)
                const bool is_x = te < 64;
                const int c = (is_x ? mt : nt) * 64 + (te & 63);
                if (is_x ? (nt == 0 && c < p.Ne) : (mt == 0 && c < p.Nm)) {
                    const long long o = (long long)f * (is_x ? p.Ne : p.Nm) + c;
                    (is_x ? p.pxx : p.pyy)[o] = pwa[te];
                    float* user = is_x ? p.sxx_out : p.syy_out;
                    if (user) user[o] = pwa[te];
                }
            }
            mn_copy_out(p, stage_tile, te, last ? 384 : 128, mt, nt, f);
            mbar_arrive(&bars->pw_empty[acc]);
            if (!last) asm volatile("bar.sync 1, 128;" ::: "memory");
            if (t == t0 && te == 0) K2_STAMP(42);
            ++it;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 256);
    if (threadIdx.x == 0) K2_STAMP(43);
}

#ifdef CMC_K2_TRACE
}  // namespace cmc
extern "C" CMC_API int cmc_dbg_k2_trace(unsigned long long* out) {
    return (int)cudaMemcpyFromSymbol(out, cmc::g_k2_trace, sizeof(unsigned long long) * 64);
}
namespace cmc {
#endif

// Spectra [L][F][ld] complex64 as a 3-D fp32 tensor: dim0 = 2 * n_ch floats of one (l, f), dim1 = F, dim2 = L;
// box = 32 floats x 1 bin x 32 segments.  Out-of-range channels / segments read as zeros.
static int make_spectra_map(CUtensorMap* m, const float* base, int L, int F, int n_ch, int64_t ld) {
    EncodeTiledFn enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    cuuint64_t dims[3] = {(cuuint64_t)(2 * (int64_t)n_ch), (cuuint64_t)F, (cuuint64_t)L};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8, (cuuint64_t)F * (cuuint64_t)ld * 8};
    cuuint32_t box[3] = {32, 1, (cuuint32_t)kMnKBlock};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (spectra) failed with CUresult %d (L=%d F=%d n_ch=%d ld=%lld)", (int)r, L, F,
                  n_ch, (long long)ld);
        return CMC_ECUDA;
    }
    return CMC_OK;
}

// Whether the direct kernel can read these spectra: TMA needs 16-byte aligned rows.
bool csd_direct_ok(const float* X, const float* Y, int64_t ldx, int64_t ldy) {
    return (ldx & 1) == 0 && (ldy & 1) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(Y) & 15) == 0;
}

// coherence (+ auto-spectra into ws / user arrays, optional sxy) without touching the operand planes of ws
int csd_msc_direct(const float* X, const float* Y, int L, int F, int Ne, int Nm, int64_t ldx, int64_t ldy, float* coh,
                   float* sxx, float* syy, float* sxy, unsigned char* w, const CsdLayout& y, cudaStream_t st) {
    CUtensorMap mX, mY;
    int rc;
    if ((rc = make_spectra_map(&mX, X, L, F, Ne, ldx))) return rc;
    if ((rc = make_spectra_map(&mY, Y, L, F, Nm, ldy))) return rc;
    MnParams p{};
    p.F = F; p.MT = y.MT; p.NT = y.NT; p.Ne = Ne; p.Nm = Nm; p.KB = (L + kMnKBlock - 1) / kMnKBlock;
    p.coh = coh;
    p.sxy = reinterpret_cast<float2*>(sxy);
    p.pxx = reinterpret_cast<float*>(w + y.off_pxx);
    p.pyy = reinterpret_cast<float*>(w + y.off_pyy);
    p.sxx_out = sxx;
    p.syy_out = syy;
    p.total_tiles = (long long)F * y.MT * y.NT;
    const size_t smem = 1024 + (size_t)kMnStages * kMnStageBytes + (size_t)64 * kMnPitch * 4 + (2 * 256 + 128) * 4 +
                        sizeof(MnBarriers) + 16;
    rc = ensure_smem_attr(reinterpret_cast<const void*>(csd_mn_kernel), smem);
    if (rc) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long grid = p.total_tiles < sms ? p.total_tiles : sms;
    csd_mn_kernel<<<(unsigned)grid, kMnThreads, smem, st>>>(mX, mY, p);
    CMC_CHECK_LAUNCH("csd_mn_kernel");
    return CMC_OK;
}

}  // namespace cmc

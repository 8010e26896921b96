// K1t: band-limited hann-windowed Welch spectra on the tensor cores (tcgen05.mma kind::f16, BF16 x 3).
//
// The Welch path of the reference is scipy.signal.coherence with its defaults (preprocessing.py:1228-1230): periodic
// hann window, 50 % overlap, per-segment constant detrend - and the callers only keep a narrow band (1 - 100 Hz =
// 100 of the 1,025 bins at nperseg 2048).  A full FFT per segment computes ten times the bins that are kept and
// transforms every sample twice.  This kernel computes exactly what is kept, and every sample once:
//
//   * half blocks.  With hop N / 2 a segment is two half blocks of N / 2 samples, and each half block belongs to two
//     segments.  For the rectangular-window DFT, R_s[b] = P_h[b] + (-1)^b P_{h+1}[b] with the HALF-BLOCK sums
//     P_h[b] = sum_{n < N/2} x[h N/2 + n] exp(-2 pi i b n / N)        (one GEMM row per half block).
//   * hann in the frequency domain.  w[n] = 1/2 - 1/2 cos(2 pi n / N) is a three-tap filter over bins:
//     X_s[b] = 1/2 R_s[b] - 1/4 (R_s[b-1] + R_s[b+1]).  Removing the segment mean only changes R_s[0] (-> 0).
//   * so ONE GEMM per half block, D[channel][(bin, re/im)] = sum_n x[n][channel] W[(bin, re/im)][n] with the
//     constant table W = (cos, -sin)(2 pi b n / N), M = 128 channels (64 EEG + 64 EMG), N = 208 columns = 104 bins,
//     K = N / 2 samples, and an epilogue that emits  1/2 P[b] -+ 1/4 (P[b-1] + P[b+1])  into the two segments the
//     half block belongs to (second-half emission carries the sign (-1)^b).
//
// Precision.  Operands are split into BF16 hi + lo (16 significand bits) and contracted with three MMAs
// (lo*hi + hi*lo + hi*hi, FP32 accumulation in TMEM): relative error ~1e-5 of the spectrum's rms, i.e. <= ~1e-6 on a
// coherence - inside the 1e-4 gate.  Before the split every channel is shifted by the first sample of its chain of
// overlapping segments (y = x - c; exact for b != 0 because both halves of a segment use the same c, undone for
// b = 0), so a DC offset does not eat the 16 bits.
//
// Determinism without a zero-fill.  Every output element receives exactly two contributions, from two consecutive
// half blocks of a chain.  Half blocks of even chain position STORE, odd ones ADD (red.global.add.v2.f32); add-phase
// epilogues wait on a device counter until every store-phase unit of the launch has finished (store units come first
// in every CTA's static schedule, all CTAs are co-resident: no deadlock).  store + one add is order independent.
//
// Roles (512 threads, one persistent CTA per SM): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue (TMEM lane = channel), warps 8-15 converters.  Two 84 KB stages: the raw FP32 tile
// [64 samples][128 channels] lands where the A operand will live; the converters read it into registers, meet at a
// barrier and write the K-major BF16 hi / lo planes (manual 128-byte swizzle) in place; W hi / lo k-blocks arrive
// by TMA next to it.  Two 208-column accumulators (TMEM columns 0 and 256) overlap the epilogue with the next unit.
#include "common.cuh"
#include "csd_layout.cuh"
#include "tc_common.cuh"
#include "tile_counter.cuh"

#include <cuda_bf16.h>
#include <algorithm>
#include <numeric>
#include <vector>

namespace cmc {

using namespace tc;

constexpr int kDtThreads = 512;
constexpr int kDtConvThreads = 256;                          // warps 8-15
constexpr int kDtStages = 2;
constexpr int kDtKB = 64;                                    // samples per k-block (128 bytes of BF16)
constexpr int kDtM = 128;                                    // channels per unit (two 64-channel groups)
constexpr int kDtBins = 104;                                 // accumulator bins
constexpr int kDtCols = 2 * kDtBins;                         // 208 accumulator columns (re, im interleaved)
constexpr int kDtPlaneA = kDtM * kDtKB * 2;                  // 16 KB: one BF16 plane of the A k-block
constexpr int kDtABytes = 2 * kDtPlaneA;                     // 32 KB = the raw FP32 tile [64][128]
constexpr int kDtPlaneB = kDtCols * kDtKB * 2;               // 26 KB: one BF16 plane of the W k-block
constexpr int kDtStageBytes = kDtABytes + 2 * kDtPlaneB;     // 84 KB
constexpr int kDtChainCap = 32;                              // segments per chain (bounds the drift y = x - c sees)

struct DtItem {           // one half block
    int x_row;            // first sample
    int c_row;            // sample whose value is subtracted before the BF16 split (first sample of the chain)
    int seg_a;            // segment whose FIRST half this is, or -1
    int seg_b;            // segment whose SECOND half this is, or -1
    int phase;            // 0: emissions are stores, 1: emissions are adds
    int pad[3];
};

struct DtParams {
    const DtItem* items;
    int n_items, n_gp;             // half blocks, channel-group pairs per half block (units = n_items * n_gp)
    int n_store_units;             // units with phase 0 (they come first)
    int KB, N, bin_lo, F, b0, detrend;
    const float* x[2];
    long long ld[2];
    int n_ch[2];
    int n_grp0;                    // 64-channel groups of recording 0 (groups of recording 1 follow)
    float2* spec[2];
    long long spec_ld;
    TileCounter* ctr;
};

struct __align__(8) DtBarriers {
    uint64_t full[kDtStages];
    uint64_t conv[kDtStages];
    uint64_t empty[kDtStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ float dt_lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void dt_sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned dt_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void dt_emit(float2* o, float re, float im, int add) {
    if (add)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(o), "f"(re), "f"(im) : "memory");
    else
        asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(o), "f"(re), "f"(im) : "memory");
}
// BF16 hi / lo split of two values: hi = rn(v), lo = rn(v - hi); element 0 in the low half-word
__device__ __forceinline__ void dt_split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - __low2float(h), v1 - __high2float(h));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// group g of the launch -> (recording, first channel); groups past the last read as zeros (rec 0, channels out of range)
__device__ __forceinline__ void dt_group(const DtParams& p, int g, int& rec, int& c0) {
    const int n_grp1 = (p.n_ch[1] + 63) >> 6;
    if (g < p.n_grp0) { rec = 0; c0 = g * 64; }
    else if (g < p.n_grp0 + n_grp1) { rec = 1; c0 = (g - p.n_grp0) * 64; }
    else { rec = 0; c0 = p.n_grp0 * 64; }
}

__global__ void __launch_bounds__(kDtThreads, 1)
dft_hann_tc_kernel(const __grid_constant__ CUtensorMap mX0, const __grid_constant__ CUtensorMap mX1,
                   const __grid_constant__ CUtensorMap mW, const DtParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    DtBarriers* bars = reinterpret_cast<DtBarriers*>(base + kDtStages * kDtStageBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_units = p.n_items * p.n_gp;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kDtStages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->conv[s], kDtConvThreads);
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->tmem_full[a], 1);
            mbar_init(&bars->tmem_empty[a], 4);
        }
        fence_barrier_init();
        tma_prefetch_desc(&mX0);
        tma_prefetch_desc(&mX1);
        tma_prefetch_desc(&mW);
    }
    if (warp == 2) {
        tmem_alloc(&bars->tmem_base, 512);
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
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const DtItem it = p.items[u / p.n_gp];
                const int gp = u % p.n_gp;
                int ra, ca, rb, cb;
                dt_group(p, 2 * gp, ra, ca);
                dt_group(p, 2 * gp + 1, rb, cb);
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars->full[stage], kDtStageBytes);
                    unsigned char* st = base + stage * kDtStageBytes;
                    const int row = it.x_row + kb * kDtKB;
                    tma_load_2d(st, ra ? &mX1 : &mX0, &bars->full[stage], ca, row);
                    tma_load_2d(st + kDtPlaneA, rb ? &mX1 : &mX0, &bars->full[stage], cb, row);
                    tma_load_2d(st + kDtABytes, &mW, &bars->full[stage], kb * kDtKB, 0);
                    tma_load_2d(st + kDtABytes + kDtPlaneB, &mW, &bars->full[stage], kb * kDtKB, kDtCols);
                    if (++stage == kDtStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(kDtM, kDtCols);
            int stage = 0;
            uint32_t phase = 0, n = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const uint32_t acc = n & 1, accphase = (n >> 1) & 1;
                mbar_wait(&bars->tmem_empty[acc], accphase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * 256;
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(&bars->full[stage], phase);          // W planes (async proxy) have landed
                    mbar_wait(&bars->conv[stage], phase);          // A planes written by the converters
                    tc_fence_after();
                    const uint32_t ahi = smem_u32(base + stage * kDtStageBytes), alo = ahi + kDtPlaneA;
                    const uint32_t bhi = ahi + kDtABytes, blo = bhi + kDtPlaneB;
#pragma unroll
                    for (int k = 0; k < kDtKB / 16; ++k) {
                        const uint64_t dah = make_smem_desc_k_sw128(ahi + k * 32), dal = make_smem_desc_k_sw128(alo + k * 32);
                        const uint64_t dbh = make_smem_desc_k_sw128(bhi + k * 32), dbl = make_smem_desc_k_sw128(blo + k * 32);
                        umma_f16(d, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);      // small terms first
                        umma_f16(d, dah, dbl, idesc, 1u);
                        umma_f16(d, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&bars->empty[stage]);
                    if (++stage == kDtStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars->tmem_full[acc]);
                ++n;
            }
        }
    } else if (warp >= 8) {
        // ===================== converters: y = x - c, BF16 hi / lo planes in place =====================
        const int t = threadIdx.x - 256;
        const int m = t & 127;                   // A row = channel of the unit
        const int g = t >> 7;                    // samples 32 g .. 32 g + 31 of the k-block
        int stage = 0;
        uint32_t phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const DtItem it = p.items[u / p.n_gp];
            const int gp = u % p.n_gp;
            int rec, c0;
            dt_group(p, 2 * gp + (m >> 6), rec, c0);
            const int ch = c0 + (m & 63);
            const float* xr = rec ? p.x[1] : p.x[0];
            const long long ldr = rec ? p.ld[1] : p.ld[0];
            const float c = ch < (rec ? p.n_ch[1] : p.n_ch[0]) ? __ldg(xr + (long long)it.c_row * ldr + ch) : 0.f;
            for (int kb = 0; kb < p.KB; ++kb) {
                mbar_wait(&bars->full[stage], phase);
                const uint32_t sb = smem_u32(base + stage * kDtStageBytes);
                const uint32_t src = sb + (uint32_t)((m >> 6) * kDtPlaneA + (m & 63) * 4 + g * 32 * 256);
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = dt_lds32(src + i * 256) - c;
                asm volatile("bar.sync 2, 256;" ::: "memory");       // every raw value is in a register
                const uint32_t row = sb + (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 hi, lo;
                    dt_split2(v[8 * q], v[8 * q + 1], hi.x, lo.x);
                    dt_split2(v[8 * q + 2], v[8 * q + 3], hi.y, lo.y);
                    dt_split2(v[8 * q + 4], v[8 * q + 5], hi.z, lo.z);
                    dt_split2(v[8 * q + 6], v[8 * q + 7], hi.w, lo.w);
                    const uint32_t off = (uint32_t)(((4 * g + q) ^ (m & 7)) << 4);     // 128-byte swizzle
                    dt_sts128(row + off, hi);
                    dt_sts128(row + kDtPlaneA + off, lo);
                }
                fence_proxy_async();             // generic-proxy writes -> visible to the MMA's async-proxy reads
                mbar_arrive(&bars->conv[stage]);
                if (++stage == kDtStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: three-tap hann, two emissions per half block =====================
        const int q = warp - 4;                  // TMEM lane quadrant
        const int m = threadIdx.x - 128;         // accumulator lane = channel of the unit
        uint32_t n = 0;
        const unsigned store_target = 4u * (unsigned)p.n_store_units;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const DtItem it = p.items[u / p.n_gp];
            const int gp = u % p.n_gp;
            int rec, c0;
            dt_group(p, 2 * gp + (m >> 6), rec, c0);
            const int ch = c0 + (m & 63);
            const bool valid = ch < (rec ? p.n_ch[1] : p.n_ch[0]);
            float2* sp = rec ? p.spec[1] : p.spec[0];
            float2* outA = (valid && it.seg_a >= 0) ? sp + (long long)it.seg_a * p.F * p.spec_ld + ch : nullptr;
            float2* outB = (valid && it.seg_b >= 0) ? sp + (long long)it.seg_b * p.F * p.spec_ld + ch : nullptr;
            float dc_fix = 0.f;                  // (N / 2) c: what y = x - c removed from P[0]
            if (p.b0 == 0 && p.detrend != CMC_DETREND_CONSTANT && valid)
                dc_fix = 0.5f * (float)p.N * __ldg((rec ? p.x[1] : p.x[0]) + (long long)it.c_row * (rec ? p.ld[1] : p.ld[0]) + ch);
            const uint32_t acc = n & 1, accphase = (n >> 1) & 1;
            mbar_wait(&bars->tmem_full[acc], accphase);
            tc_fence_after();
            if (it.phase) {
                // adds may only start once every store of the launch is visible
                if (lane == 0) {
                    const long long t0 = clock64();
                    while (dt_ld_acquire(&p.ctr->next) < store_target) {
                        __nanosleep(64);
                        if (clock64() - t0 > 4000000000LL) {
                            printf("cmc: dft_hann_tc store-phase wait timed out (block %d)\n", blockIdx.x);
                            __trap();
                        }
                    }
                }
                __syncwarp();
            }
            const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
            uint32_t cur[32], nxt[32];
            tmem_ld_32x32(taddr, cur);
            tmem_ld_wait();
            if (p.b0 == 0) {                     // accumulator bin 0 is the DC bin
                if (p.detrend == CMC_DETREND_CONSTANT) cur[0] = 0u;      // segment mean removed: R[0] = 0
                else cur[0] = __float_as_uint(__uint_as_float(cur[0]) + dc_fix);
                cur[1] = 0u;
            }
            float pm_re = 0.f, pm_im = 0.f;      // P[j - 1] of the first bin of the chunk
#pragma unroll
            for (int cq = 0; cq < 7; ++cq) {
                if (cq < 6) {
                    tmem_ld_32x32(taddr + 32 * (cq + 1), nxt);
                    tmem_ld_wait();
                }
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const int j = 16 * cq + jj;
                    if (j >= kDtBins - 1) break;                           // bin 103 has no right neighbour
                    const int b = p.b0 + j;
                    const float pc_re = __uint_as_float(cur[2 * jj]), pc_im = __uint_as_float(cur[2 * jj + 1]);
                    const float pp_re = __uint_as_float(jj < 15 ? cur[2 * jj + 2] : nxt[0]);
                    const float pp_im = __uint_as_float(jj < 15 ? cur[2 * jj + 3] : nxt[1]);
                    float qm_re = jj > 0 ? __uint_as_float(cur[2 * jj - 2]) : pm_re;
                    float qm_im = jj > 0 ? __uint_as_float(cur[2 * jj - 1]) : pm_im;
                    if (j == 0 && p.b0 == 0) { qm_re = pp_re; qm_im = -pp_im; }   // P[-1] = conj(P[1])
                    if (b >= p.bin_lo && b < p.bin_lo + p.F) {
                        const float s_re = 0.25f * (qm_re + pp_re), s_im = 0.25f * (qm_im + pp_im);
                        float a_re = 0.5f * pc_re - s_re, a_im = 0.5f * pc_im - s_im;
                        float b_re = 0.5f * pc_re + s_re, b_im = 0.5f * pc_im + s_im;
                        if (b & 1) { b_re = -b_re; b_im = -b_im; }
                        if (b == 0) {
                            a_im = b_im = 0.f;
                            if (p.detrend == CMC_DETREND_POST_TAPER) a_re = b_re = 0.f;
                        }
                        const long long o = (long long)(b - p.bin_lo) * p.spec_ld;
                        if (outA) dt_emit(outA + o, a_re, a_im, it.phase);
                        if (outB) dt_emit(outB + o, b_re, b_im, it.phase);
                    }
                }
                pm_re = __uint_as_float(cur[30]);
                pm_im = __uint_as_float(cur[31]);
                if (cq < 6) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) cur[i] = nxt[i];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);   // accumulator may be overwritten
            if (!it.phase) {
                __threadfence();                 // this thread's stores before the warp's arrival
                __syncwarp();
                if (lane == 0) atomicAdd(&p.ctr->next, 1u);
            }
            ++n;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
    if (threadIdx.x == 0) {
        // the last CTA to leave hands the counter back at zero (stream-ordered launches and graph replays reuse it)
        __threadfence();
        if (atomicAdd(&p.ctr->done, 1u) == gridDim.x - 1u) {
            p.ctr->next = 0u;
            p.ctr->done = 0u;
            __threadfence();
        }
    }
}

// W[r][n], r = 2 j + part: part 0 = cos, part 1 = -sin of 2 pi (b0 + j) n / N; rows [0, 208) hi, [208, 416) lo
__global__ void dft_w_table_kernel(__nv_bfloat16* W, int Kw, int N, int b0) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (n >= Kw) return;
    const int b = b0 + (r >> 1);
    const long long q = ((long long)b * n) % N;
    double s, c;
    sincospi(2.0 * (double)q / (double)N, &s, &c);
    const double v = (r & 1) ? -s : c;
    const __nv_bfloat16 hi = __double2bfloat16(v);
    const __nv_bfloat16 lo = __double2bfloat16(v - (double)__bfloat162float(hi));
    W[(long long)r * Kw + n] = hi;
    W[(long long)(kDtCols + r) * Kw + n] = lo;
}

struct WelchHannPlan {
    int dev, N, bin_lo, F, b0, KB, n_seg, n_items, n_store;
    long long max_row_end;
    DtItem* d_items;
    __nv_bfloat16* d_W;
    CUtensorMap mW;
};

static int make_raw_map(CUtensorMap* m, const float* x, int64_t n_samples, int n_ch, int64_t ld) {
    EncodeTiledFn enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    cuuint64_t dims[2] = {(cuuint64_t)n_ch, (cuuint64_t)n_samples};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {64u, (cuuint32_t)kDtKB};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(recording, tensor-core DFT) failed with CUresult %d", (int)r);
        return CMC_ECUDA;
    }
    return CMC_OK;
}

}  // namespace cmc

using namespace cmc;

extern "C" int cmc_welch_hann_plan_create(const int64_t* seg_starts_host, int n_seg, int N, int bin_lo, int bin_hi,
                                          void** plan_out) {
    CMC_REQUIRE(seg_starts_host && plan_out, "cmc_welch_hann_plan_create: null pointer");
    *plan_out = nullptr;
    CMC_REQUIRE(n_seg >= 1, "cmc_welch_hann_plan_create: no segments");
    CMC_REQUIRE(bin_lo >= 0 && bin_hi >= bin_lo && bin_hi <= N / 2,
                "cmc_welch_hann_plan_create: bins [%d, %d] outside [0, %d]", bin_lo, bin_hi, N / 2);
    const int F = bin_hi - bin_lo + 1;
    const int b0 = bin_lo > 0 ? bin_lo - 1 : 0;
    // accumulator bins b0 .. b0 + 103 must cover bin_lo - 1 .. bin_hi + 1 and stay below the Nyquist bin
    if (N < 256 || N > 16384 || (N % 128) != 0 || bin_hi + 1 - b0 > kDtBins - 1 || b0 + kDtBins > N / 2) {
        set_error("cmc_welch_hann_plan_create: N=%d bins [%d, %d] outside the tensor-core kernel (N %% 128 == 0, "
                  "256 <= N <= 16384, at most %d bins, band below Nyquist)", N, bin_lo, bin_hi, kDtBins - 2);
        return CMC_EUNSUPPORTED;
    }
    for (int s = 0; s < n_seg; ++s)
        CMC_REQUIRE(seg_starts_host[s] >= 0 && seg_starts_host[s] + N < (1ll << 31),
                    "cmc_welch_hann_plan_create: segment start %lld out of range", (long long)seg_starts_host[s]);
    // chains of segments that overlap by exactly N / 2 share their half blocks
    std::vector<int> order(n_seg);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return seg_starts_host[a] < seg_starts_host[b]; });
    std::vector<DtItem> items;
    long long max_end = 0;
    for (int i = 0; i < n_seg;) {
        int j = i;
        while (j + 1 < n_seg && j + 1 - i < kDtChainCap &&
               seg_starts_host[order[j + 1]] == seg_starts_host[order[j]] + N / 2)
            ++j;
        const int len = j - i + 1;
        const long long first = seg_starts_host[order[i]];
        for (int h = 0; h <= len; ++h) {
            DtItem it{};
            it.x_row = (int)(first + (long long)h * (N / 2));
            it.c_row = (int)first;
            it.seg_a = h < len ? order[i + h] : -1;
            it.seg_b = h >= 1 ? order[i + h - 1] : -1;
            it.phase = h & 1;
            items.push_back(it);
        }
        max_end = std::max(max_end, first + (long long)(len + 1) * (N / 2));
        i = j + 1;
    }
    std::stable_sort(items.begin(), items.end(), [](const DtItem& a, const DtItem& b) { return a.phase < b.phase; });
    auto* pl = new WelchHannPlan();
    pl->N = N; pl->bin_lo = bin_lo; pl->F = F; pl->b0 = b0; pl->KB = N / 2 / kDtKB; pl->n_seg = n_seg;
    pl->n_items = (int)items.size();
    pl->n_store = (int)std::count_if(items.begin(), items.end(), [](const DtItem& a) { return a.phase == 0; });
    pl->max_row_end = max_end;
    pl->d_items = nullptr; pl->d_W = nullptr;
    int rc = check_cuda(cudaGetDevice(&pl->dev), "cudaGetDevice");
    const int Kw = N / 2;
    if (!rc) rc = check_cuda(cudaMalloc(&pl->d_items, items.size() * sizeof(DtItem)), "cudaMalloc(plan items)");
    if (!rc) rc = check_cuda(cudaMemcpy(pl->d_items, items.data(), items.size() * sizeof(DtItem), cudaMemcpyHostToDevice),
                             "cudaMemcpy(plan items)");
    if (!rc) rc = check_cuda(cudaMalloc(&pl->d_W, (size_t)2 * kDtCols * Kw * sizeof(__nv_bfloat16)), "cudaMalloc(plan W)");
    if (!rc) {
        dft_w_table_kernel<<<dim3((Kw + 127) / 128, kDtCols), 128>>>(pl->d_W, Kw, N, b0);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        rc = check_cuda(cudaGetLastError(), "dft_w_table_kernel");
    }
    if (!rc) rc = check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize(plan)");
    if (!rc) rc = make_kmajor_map(&pl->mW, pl->d_W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Kw, 2 * kDtCols, kDtCols);
    if (rc) {
        cudaFree(pl->d_items);
        cudaFree(pl->d_W);
        delete pl;
        return rc;
    }
    *plan_out = pl;
    return CMC_OK;
}

extern "C" int cmc_welch_hann_plan_destroy(void* plan) {
    if (!plan) return CMC_OK;
    auto* pl = static_cast<WelchHannPlan*>(plan);
    cudaFree(pl->d_items);
    cudaFree(pl->d_W);
    delete pl;
    return CMC_OK;
}

extern "C" int cmc_welch_hann_plan_info(const void* plan, int* n_half_blocks, int* n_segments, int* n_bins) {
    CMC_REQUIRE(plan, "cmc_welch_hann_plan_info: null plan");
    const auto* pl = static_cast<const WelchHannPlan*>(plan);
    if (n_half_blocks) *n_half_blocks = pl->n_items;
    if (n_segments) *n_segments = pl->n_seg;
    if (n_bins) *n_bins = pl->F;
    return CMC_OK;
}

extern "C" int cmc_welch_hann_spectra(const void* plan, const float* x1, int n_ch1, int64_t ld1, float* spec1,
                                      const float* x2, int n_ch2, int64_t ld2, float* spec2, int64_t n_samples,
                                      int detrend, int64_t spec_ld, void* stream) {
    CMC_REQUIRE(plan && x1 && spec1, "cmc_welch_hann_spectra: null pointer");
    const auto* pl = static_cast<const WelchHannPlan*>(plan);
    if (!x2) n_ch2 = 0;
    CMC_REQUIRE(n_ch1 >= 1 && ld1 >= n_ch1 && n_ch2 >= 0 && (n_ch2 == 0 || (spec2 && ld2 >= n_ch2)) &&
                spec_ld >= n_ch1 && spec_ld >= n_ch2, "cmc_welch_hann_spectra: bad channel count / pitch");
    CMC_REQUIRE(detrend >= 0 && detrend <= 2, "cmc_welch_hann_spectra: detrend must be 0, 1 or 2");
    CMC_REQUIRE(pl->max_row_end <= n_samples, "cmc_welch_hann_spectra: segment outside the recording (%lld > %lld)",
                pl->max_row_end, (long long)n_samples);
    auto tma_ok = [](const float* x, int64_t ld) { return (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0; };
    if (!tma_ok(x1, ld1) || (n_ch2 && !tma_ok(x2, ld2)) || (reinterpret_cast<uintptr_t>(spec1) & 7) ||
        (n_ch2 && (reinterpret_cast<uintptr_t>(spec2) & 7))) {
        set_error("cmc_welch_hann_spectra: recordings need 16-byte aligned rows (channel pitch %% 4 == 0), spectra 8-byte alignment");
        return CMC_EUNSUPPORTED;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    CMC_REQUIRE(dev == pl->dev, "cmc_welch_hann_spectra: plan built on device %d, current device %d", pl->dev, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TileCounter* ctr = tile_counter_for(dev, st);
    if (!ctr) {
        set_error("cmc_welch_hann_spectra: no launch counter available (too many captured launches)");
        return CMC_EUNSUPPORTED;
    }
    CUtensorMap m0, m1;
    int rc = make_raw_map(&m0, x1, n_samples, n_ch1, ld1);
    if (rc) return rc;
    if (n_ch2) { if ((rc = make_raw_map(&m1, x2, n_samples, n_ch2, ld2))) return rc; }
    else m1 = m0;
    DtParams p{};
    p.items = pl->d_items;
    p.n_items = pl->n_items;
    p.n_grp0 = (n_ch1 + 63) / 64;
    const int n_grp = p.n_grp0 + (n_ch2 + 63) / 64;
    p.n_gp = (n_grp + 1) / 2;
    p.n_store_units = pl->n_store * p.n_gp;
    p.KB = pl->KB; p.N = pl->N; p.bin_lo = pl->bin_lo; p.F = pl->F; p.b0 = pl->b0; p.detrend = detrend;
    p.x[0] = x1; p.x[1] = n_ch2 ? x2 : x1;
    p.ld[0] = ld1; p.ld[1] = n_ch2 ? ld2 : ld1;
    p.n_ch[0] = n_ch1; p.n_ch[1] = n_ch2;
    p.spec[0] = reinterpret_cast<float2*>(spec1);
    p.spec[1] = reinterpret_cast<float2*>(n_ch2 ? spec2 : spec1);
    p.spec_ld = spec_ld;
    p.ctr = ctr;
    const size_t smem = 1024 + (size_t)kDtStages * kDtStageBytes + sizeof(DtBarriers) + 16;
    rc = ensure_smem_attr(reinterpret_cast<const void*>(dft_hann_tc_kernel), smem);
    if (rc) return rc;
    const long long n_units = (long long)p.n_items * p.n_gp;
    const unsigned grid = (unsigned)(n_units < sms ? n_units : sms);
    dft_hann_tc_kernel<<<grid, kDtThreads, smem, st>>>(m0, m1, pl->mW, p);
    CMC_CHECK_LAUNCH("dft_hann_tc_kernel");
    return CMC_OK;
}

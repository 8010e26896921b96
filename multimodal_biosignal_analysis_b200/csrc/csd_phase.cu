// K3 (phase mode): phase-randomised surrogate null as ONE dense GEMM per frequency on the tensor
// cores (tcgen05.mma kind::f16, FP16 inputs, FP32 accumulation in TMEM).
//
// Operand precision.  The definition (oracle/surrogate.py) uses exact unit-circle phases and unquantised
// cross-products; this kernel rounds both to FP16 (11-bit significand - kind::f16 runs FP16 and BF16 at the same
// rate, FP16 rounds 8 x finer).  Whitened cross-products obey |Z| <= 1, so a fixed 2^14 prescale keeps them in
// the normal FP16 range down to 2^-28.  The rounding moves a surrogate coherence by <~ 3.5e-3 / L (measured
// maximum over 2e5 samples), i.e. <= 4e-5 for L > 85.  Shorter averages (L <= kPhSplitMaxL) run the error-compensated
// three-term split instead - operands hi + lo = 22 bits, A' = [P_hi | P_hi | P_lo], B' = [Z_hi | Z_lo | Z_hi]
// concatenated along K (K' = 6 L <= 512 still fits the resident panel) - so every L stays inside the 1e-4 gate.
//
// Surrogate s rotates every EMG spectrum by one phase per (segment l, frequency f), shared by all EMG
// channels:  S_s[i][j] = sum_l conj(Xh[l][i]) Yh[l][j] P[s][l]  (Xh, Yh whitened, P on the unit circle).
// With Z[l][(i,j)] = conj(Xh[l][i]) Yh[l][j] - computed ONCE - every surrogate is a row of
//     D[s][n] = sum_k A[s][k] B[n][k],   K = (l, re/im),
//     A[s]        = (P_re, P_im) interleaved                      M = surrogates
//     B[re-form]  = (Z_re, -Z_im),  B[im-form] = (Z_im, Z_re)      N = 2 x pairs
// so Re S and Im S of one pair land in the SAME accumulator lane (columns c and 128 + c of a
// 128-pair tile) and the epilogue is lane-local: C = Re^2 + Im^2, compare with the observed
// coherence (warp ballot -> per-pair exceedance count), running max per surrogate.
//
// CTA work unit = one (frequency, 128-surrogate panel): the A panel (128 x K) stays resident in
// shared memory while the 256-row B tiles of all pair groups stream through a 3-stage TMA ring;
// two 256-column TMEM accumulators overlap the epilogue with the next tile's MMAs.
#include "csd_layout.cuh"

#include <cuda_fp16.h>
#include <math.h>
#include <map>
#include <mutex>
#include <vector>

namespace cmc {

using namespace tc;

constexpr int kPhM = 128;            // surrogates per panel
constexpr int kPhPairs = 128;        // pairs per B tile
constexpr int kPhN = 2 * kPhPairs;   // B tile rows (re-forms then im-forms)
constexpr int kPhKB = 64;            // fp16 per k-block = 128 bytes
constexpr int kPhStages = 3;             // B ring of the resident-panel variant
constexpr int kPhStagesStream = 4;       // (A block + B block) ring of the streamed-panel variant
constexpr int kPhABytes = kPhM * 128;    // 16 KB: one k-block of the A panel
constexpr int kPhBBytes = kPhN * 128;    // 32 KB: one k-block of a B tile
constexpr int kPhThreads = 256;
constexpr int kPhaseN = 1 << CMC_PHASE_TABLE_BITS;
constexpr float kZScale = 16384.0f;                      // 2^14 prescale of the whitened cross-products
constexpr float kZUnscaleSq = 1.0f / (16384.0f * 16384.0f);
constexpr int kPhSplitMaxL = 85;                        // 6 L <= 512: the three-term operands still fit the resident panel

// ------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// float -> IEEE binary16 bits, round to nearest even (|x| <= 1: no overflow handling needed beyond inf)
static uint16_t host_f16_rne(float x) {
    return __half_as_ushort(__float2half_rn(x));
}
static float host_f16_to_float(uint16_t h) {
    return __half2float(__ushort_as_half(h));
}

// per-device tables [2][4096]: P_hi[a] = (fp16(cos), fp16(sin)) of 2 pi a / 4096 packed as two fp16 in a uint32,
// followed by P_lo[a] = fp16(P[a] - P_hi[a]) (second operand term of the short-L split)
static void host_phase_entry(int a, uint32_t* hi, uint32_t* lo) {
    const double ang = 6.283185307179586476925286766559 * a / kPhaseN;
    const double c = cos(ang), sn = sin(ang);
    const uint16_t ch = host_f16_rne((float)c), sh = host_f16_rne((float)sn);
    *hi = (uint32_t)ch | ((uint32_t)sh << 16);
    const uint16_t cl = host_f16_rne((float)(c - (double)host_f16_to_float(ch)));
    const uint16_t sl = host_f16_rne((float)(sn - (double)host_f16_to_float(sh)));
    *lo = (uint32_t)cl | ((uint32_t)sl << 16);
}

static int get_phase_table(const uint32_t** table) {
    static std::mutex mu;
    static std::map<int, uint32_t*> cache;
    int dev = 0;
    int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(dev);
    if (it == cache.end()) {
        std::vector<uint32_t> h(2 * kPhaseN);
        for (int a = 0; a < kPhaseN; ++a) host_phase_entry(a, &h[a], &h[kPhaseN + a]);
        uint32_t* d = nullptr;
        rc = check_cuda(cudaMalloc(&d, h.size() * 4), "cudaMalloc(phase table)");
        if (rc) return rc;
        rc = check_cuda(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice), "cudaMemcpy(phase table)");
        if (rc) return rc;
        it = cache.emplace(dev, d).first;
    }
    *table = it->second;
    return CMC_OK;
}

// A operand: Phi[f][s_local][k] (fp16, K-major, row length KPb); one thread = (s, complex column, group of 4
// frequencies).  Complex column lp = term * L + l: terms 0 / 1 carry P_hi, term 2 carries P_lo (nterms = 1 or 3).
__global__ void __launch_bounds__(256)
phase_gen_kernel(const uint32_t* __restrict__ table, uint64_t seed, int64_t s_begin, int n_local, int S_pad, int L,
                 int nterms, int F, int f_lo, int KPb, __half* __restrict__ A) {
    // frequencies [f_lo, f_lo + F) of the GLOBAL axis land in rows [0, F) of A; f_lo need not be a multiple of 4
    const int lp = blockIdx.x * blockDim.x + threadIdx.x;      // complex column index, covers [0, KPb / 2)
    const int s = blockIdx.y;                                   // local surrogate row, covers [0, S_pad)
    const int fg0 = f_lo >> 2;
    const int fg = blockIdx.z;                                  // frequency group of 4, relative to fg0
    if (lp >= KPb / 2) return;
    uint32_t vals[4] = {0u, 0u, 0u, 0u};
    if (s < n_local && lp < nterms * L) {
        const int term = nterms == 1 ? 0 : lp / L, l = lp - term * L;
        const uint32_t* tab = table + (term == 2 ? kPhaseN : 0);
        const uint64_t sg = (uint64_t)(s_begin + s);
        // counter = (surrogate, segment, GLOBAL frequency group, surrogate >> 32): word q is frequency 4 fg + q
        const uint4 r = philox4x32_10(make_uint4((uint32_t)sg, (uint32_t)l, (uint32_t)(fg0 + fg), (uint32_t)(sg >> 32)),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        vals[0] = __ldg(tab + (r.x >> (32 - CMC_PHASE_TABLE_BITS)));
        vals[1] = __ldg(tab + (r.y >> (32 - CMC_PHASE_TABLE_BITS)));
        vals[2] = __ldg(tab + (r.z >> (32 - CMC_PHASE_TABLE_BITS)));
        vals[3] = __ldg(tab + (r.w >> (32 - CMC_PHASE_TABLE_BITS)));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int f = (fg0 + fg) * 4 + q - f_lo;                // row of A
        if (f >= 0 && f < F)
            reinterpret_cast<uint32_t*>(A)[(((int64_t)f * S_pad + s) * KPb) / 2 + lp] = vals[q];
    }
}

// B operand: 2^14 Z re-/im-form rows (fp16, K-major) from the whitened 3xTF32 operands left by cmc_csd_msc.
// grid (F, Ne); block 256: the thread block holds Xh row i in shared memory and sweeps j.  Complex column
// lp = term * L + l: terms 0 / 2 carry Z_hi = fp16(2^14 Z), term 1 carries Z_lo = fp16(2^14 Z - Z_hi).
template <int NTERMS>
__global__ void __launch_bounds__(256)
z_gen_kernel(const float* __restrict__ Ahi, const float* __restrict__ Alo, const float* __restrict__ Bhi,
             const float* __restrict__ Blo, const float* __restrict__ pxx, const float* __restrict__ pyy, int L, int Ne,
             int Nm, int MT, int NT, int KP, int LB, int KPb, int R_pad, __half* __restrict__ Z) {
    extern __shared__ float2 xs[];                      // [L] complex Xh[l][i]
    const int f = blockIdx.x, i = blockIdx.y;
    const int half = KPb / 2;
    const float* ah = Ahi + ((int64_t)(f * MT + (i >> 6)) * kTileM + (i & 63)) * KP;
    const float* al = Alo + ((int64_t)(f * MT + (i >> 6)) * kTileM + (i & 63)) * KP;
    const float px = pxx[(int64_t)f * Ne + i];
    const float sx = px > 0.f ? rsqrtf(px) * kZScale : 0.f;      // whitening: Xh = X / sqrt(Pxx), times the prescale
    for (int l = threadIdx.x; l < L; l += blockDim.x)
        xs[l] = make_float2((ah[2 * l] + al[2 * l]) * sx, (ah[2 * l + 1] + al[2 * l + 1]) * sx);
    __syncthreads();
    for (int j = 0; j < Nm; ++j) {
        const float* bh = Bhi + ((int64_t)(f * NT + (j >> 6)) * kTileN + (j & 63)) * LB;
        const float* bl = Blo + ((int64_t)(f * NT + (j >> 6)) * kTileN + (j & 63)) * LB;
        const float py = pyy[(int64_t)f * Nm + j];
        const float sy = py > 0.f ? rsqrtf(py) : 0.f;
        const int p = i * Nm + j;
        const int64_t row_re = (int64_t)f * R_pad + (p / kPhPairs) * kPhN + (p % kPhPairs);
        uint32_t* zre = reinterpret_cast<uint32_t*>(Z + row_re * KPb);
        uint32_t* zim = reinterpret_cast<uint32_t*>(Z + (row_re + kPhPairs) * KPb);
        for (int lp = threadIdx.x; lp < half; lp += blockDim.x) {
            __half2 re = __floats2half2_rn(0.f, 0.f), im = re;
            if (lp < NTERMS * L) {
                const int term = NTERMS == 1 ? 0 : lp / L, l = lp - term * L;
                const float2 x = xs[l];
                const float yr = (bh[2 * l] + bl[2 * l]) * sy, yi = (bh[2 * l + 1] + bl[2 * l + 1]) * sy;
                float zr = x.x * yr + x.y * yi;          // conj(x) * y (prescaled through sx)
                float zi = x.x * yi - x.y * yr;
                const __half hr = __float2half_rn(zr), hi = __float2half_rn(zi);
                if (term == 1) {                           // second operand term: what the FP16 rounding lost
                    zr -= __half2float(hr);
                    zi -= __half2float(hi);
                    re = __floats2half2_rn(zr, -zi);
                    im = __floats2half2_rn(zi, zr);
                } else {
                    re = __halves2half2(hr, __hneg(hi));
                    im = __halves2half2(hi, hr);
                }
            }
            zre[lp] = *reinterpret_cast<const uint32_t*>(&re);
            zim[lp] = *reinterpret_cast<const uint32_t*>(&im);
        }
    }
}

struct PhaseParams {
    int F, MT, NT, KB, n_local, n_pairs, S_pad, R_pad;
    const float* coh_obs;      // [F][n_pairs]
    uint32_t* exceed;          // [F][n_pairs]
    uint32_t* max_u;           // [S_pad] float bits
};

struct __align__(8) PhaseBarriers {
    uint64_t full[kPhStagesStream];
    uint64_t empty[kPhStagesStream];
    uint64_t a_full;
    uint64_t a_empty;
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

// STREAM_A = false: the A panel of the CTA's (frequency, surrogate panel) is resident (2L <= 512).
// STREAM_A = true : long segment axes - every ring stage carries the A k-block next to the B k-block, the panel
//                   is re-read from L2 for each pair tile (1.5 x the operand traffic, no limit on L).
template <bool STREAM_A>
__global__ void __launch_bounds__(kPhThreads, 1)
phase_gemm_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB, const PhaseParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    constexpr int kStages = STREAM_A ? kPhStagesStream : kPhStages;
    constexpr int kStageBytes = STREAM_A ? kPhABytes + kPhBBytes : kPhBBytes;
    constexpr int kBOff = STREAM_A ? kPhABytes : 0;            // B block inside a stage
    const int a_bytes = STREAM_A ? 0 : p.KB * kPhABytes;       // resident A panel: KB k-blocks of 16 KB
    unsigned char* sA = base;
    unsigned char* sB = base + a_bytes;                        // [kStages][kStageBytes]
    float* cobs_s = reinterpret_cast<float*>(sB + kStages * kStageBytes);   // [128]
    uint32_t* cnt_s = reinterpret_cast<uint32_t*>(cobs_s + kPhPairs);       // [128]
    PhaseBarriers* bars = reinterpret_cast<PhaseBarriers*>(cnt_s + kPhPairs);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_panels = p.F * p.MT;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        mbar_init(&bars->a_full, 1);
        mbar_init(&bars->a_empty, 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->tmem_full[a], 1);
            mbar_init(&bars->tmem_empty[a], 4);
        }
        fence_barrier_init();
        tma_prefetch_desc(&mA);
        tma_prefetch_desc(&mB);
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
            uint32_t phase = 0, a_phase = 0;
            for (int pn = blockIdx.x; pn < n_panels; pn += gridDim.x) {
                const int f = pn / p.MT, mt = pn - f * p.MT;
                if (!STREAM_A) {
                    mbar_wait(&bars->a_empty, a_phase ^ 1);       // previous panel fully consumed
                    mbar_arrive_expect_tx(&bars->a_full, (uint32_t)a_bytes);
                    for (int kb = 0; kb < p.KB; ++kb)
                        tma_load_2d(sA + kb * kPhABytes, &mA, &bars->a_full, kb * kPhKB, f * p.S_pad + mt * kPhM);
                    a_phase ^= 1;
                }
                for (int nt = 0; nt < p.NT; ++nt)
                    for (int kb = 0; kb < p.KB; ++kb) {
                        mbar_wait(&bars->empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&bars->full[stage], kStageBytes);
                        unsigned char* st = sB + stage * kStageBytes;
                        if (STREAM_A)
                            tma_load_2d(st, &mA, &bars->full[stage], kb * kPhKB, f * p.S_pad + mt * kPhM);
                        tma_load_2d(st + kBOff, &mB, &bars->full[stage], kb * kPhKB, f * p.R_pad + nt * kPhN);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(kPhM, kPhN);
            int stage = 0;
            uint32_t phase = 0, a_phase = 0, it = 0;
            for (int pn = blockIdx.x; pn < n_panels; pn += gridDim.x) {
                if (!STREAM_A) {
                    mbar_wait(&bars->a_full, a_phase);
                    a_phase ^= 1;
                    tc_fence_after();
                }
                for (int nt = 0; nt < p.NT; ++nt) {
                    const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
                    mbar_wait(&bars->tmem_empty[acc], accphase ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + acc * kPhN;
                    for (int kb = 0; kb < p.KB; ++kb) {
                        mbar_wait(&bars->full[stage], phase);
                        tc_fence_after();
                        const uint32_t a0 = smem_u32(STREAM_A ? sB + stage * kStageBytes : sA + kb * kPhABytes);
                        const uint32_t b0 = smem_u32(sB + stage * kStageBytes + kBOff);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16(d, make_smem_desc_k_sw128(a0 + k * 32), make_smem_desc_k_sw128(b0 + k * 32),
                                      idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit(&bars->empty[stage]);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&bars->tmem_full[acc]);
                    ++it;
                }
                if (!STREAM_A) umma_commit(&bars->a_empty);       // A panel may be overwritten
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp - 4;
        const int te = threadIdx.x - 128;                         // accumulator lane = surrogate row in the panel
        uint32_t it = 0;
        for (int pn = blockIdx.x; pn < n_panels; pn += gridDim.x) {
            const int f = pn / p.MT, mt = pn - f * p.MT;
            const int s_local = mt * kPhM + te;
            const bool s_ok = s_local < p.n_local;
            float vmax = 0.f;
            for (int nt = 0; nt < p.NT; ++nt) {
                const int pair_t = nt * kPhPairs + te;
                cobs_s[te] = pair_t < p.n_pairs ? __ldg(p.coh_obs + (int64_t)f * p.n_pairs + pair_t) : 2.0f;
                cnt_s[te] = 0;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
                mbar_wait(&bars->tmem_full[acc], accphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * kPhN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t re[32], im[32];
                    tmem_ld_32x32(taddr + ch * 32, re);
                    tmem_ld_32x32(taddr + kPhPairs + ch * 32, im);
                    tmem_ld_wait();
                    uint32_t mine = 0;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float a = __uint_as_float(re[c]), b = __uint_as_float(im[c]);
                        const float cv = fminf((a * a + b * b) * kZUnscaleSq, 1.0f);
                        const bool hit = s_ok && (cv >= cobs_s[ch * 32 + c]);
                        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                        mine = (lane == c) ? __popc(bal) : mine;
                        vmax = fmaxf(vmax, cv);
                    }
                    if (mine) atomicAdd(&cnt_s[ch * 32 + lane], mine);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (cnt_s[te]) atomicAdd(&p.exceed[(int64_t)f * p.n_pairs + pair_t], cnt_s[te]);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                ++it;
            }
            if (s_ok) atomicMax(&p.max_u[s_local], __float_as_uint(vmax));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------ per-pair null histograms
// Same GEMM with the loops swapped: CTA work unit = one (frequency, 128-pair tile), inner loop over ALL surrogate
// panels, so the histograms of the unit's 128 pairs live in shared memory for the whole null and leave the SM once
// ([128][n_bins] uint32, n_bins <= 128).  The pair tile no longer fits next to the panel, so both operands stream:
// every ring stage carries the A k-block (16 KB) next to the B k-block (32 KB); B is re-read per panel from L2.
// Bin of a surrogate coherence C: floor((C - lo[pair]) * scale[pair]); values under the window only raise the pair's
// `below` counter, values over it are dropped (lo = 0, scale = n_bins: uniform bins over [0, 1]).  Per-pair windows
// are what makes the pass cheap: a quantile search only has to resolve the upper tail of each pair's null, so the
// ~80-90 % of the surrogates that fall under the window cost a register increment instead of a shared-memory atomic
// (measured: the atomics alone were half of a full-range pass).  data_surrogation.py places the first window from
// the null's analytic mean and zooms with further passes.  One CTA owns a (frequency, pair tile) for the whole
// launch, so the flush is a plain read-add-write and successive launches (surrogate chunks) add up.
constexpr int kPhHistStages = 3;
constexpr int kPhHistMaxBins = 128;
constexpr int kPhHistThreads = 384;      // TMA / MMA / TMEM-allocator warps + TWO epilogue warpgroups

struct PhaseHistParams {
    int F, MT, NT, KB, n_local, n_pairs, S_pad, R_pad, n_bins;
    const float* bin_lo;       // [F][n_pairs] or null (0)
    const float* bin_scale;    // [F][n_pairs] or null (n_bins)
    uint32_t* hist;            // [F][n_pairs][n_bins]
    uint32_t* below;           // [F][n_pairs] surrogates under the window, or null
};

__global__ void __launch_bounds__(kPhHistThreads, 1)
phase_hist_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB,
                  const PhaseHistParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    constexpr int kStageBytes = kPhABytes + kPhBBytes;
    unsigned char* sB = base;                                                      // [kPhHistStages][48 KB]
    uint32_t* hist_s = reinterpret_cast<uint32_t*>(sB + kPhHistStages * kStageBytes);   // [128][n_bins + 1]
    const int hp = p.n_bins + 1;                  // padded histogram row: lanes that hit the same bin of 32 different
                                                  // pairs fall into 32 different banks
    float* lo_s = reinterpret_cast<float*>(hist_s + kPhPairs * hp);               // [128] window origin per pair
    float* sc_s = lo_s + kPhPairs;                                                 // [128] bins per unit of coherence
    PhaseBarriers* bars = reinterpret_cast<PhaseBarriers*>(sc_s + kPhPairs);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_units = p.F * p.NT;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kPhHistStages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->tmem_full[a], 1);
            mbar_init(&bars->tmem_empty[a], 4);
        }
        fence_barrier_init();
        tma_prefetch_desc(&mA);
        tma_prefetch_desc(&mB);
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
        if (lane == 0) {                                          // ===== TMA producer =====
            int stage = 0;
            uint32_t phase = 0;
            for (int un = blockIdx.x; un < n_units; un += gridDim.x) {
                const int f = un / p.NT, nt = un - f * p.NT;
                for (int mt = 0; mt < p.MT; ++mt)
                    for (int kb = 0; kb < p.KB; ++kb) {
                        mbar_wait(&bars->empty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&bars->full[stage], kStageBytes);
                        unsigned char* st = sB + stage * kStageBytes;
                        tma_load_2d(st, &mA, &bars->full[stage], kb * kPhKB, f * p.S_pad + mt * kPhM);
                        tma_load_2d(st + kPhABytes, &mB, &bars->full[stage], kb * kPhKB, f * p.R_pad + nt * kPhN);
                        if (++stage == kPhHistStages) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                          // ===== MMA issuer =====
            constexpr uint32_t idesc = make_idesc_f16(kPhM, kPhN);
            int stage = 0;
            uint32_t phase = 0, it = 0;
            for (int un = blockIdx.x; un < n_units; un += gridDim.x)
                for (int mt = 0; mt < p.MT; ++mt) {
                    const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
                    mbar_wait(&bars->tmem_empty[acc], accphase ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + acc * kPhN;
                    for (int kb = 0; kb < p.KB; ++kb) {
                        mbar_wait(&bars->full[stage], phase);
                        tc_fence_after();
                        const uint32_t a0 = smem_u32(sB + stage * kStageBytes);
                        const uint32_t b0 = a0 + kPhABytes;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16(d, make_smem_desc_k_sw128(a0 + k * 32), make_smem_desc_k_sw128(b0 + k * 32), idesc,
                                     (kb | k) != 0 ? 1u : 0u);
                        umma_commit(&bars->empty[stage]);
                        if (++stage == kPhHistStages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&bars->tmem_full[acc]);
                    ++it;
                }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM lane = surrogate row of the panel, columns = the unit's 128 pairs =====
        // All 32 lanes of a warp hold the SAME pair in a given column.  With a window that keeps most of the null
        // under it, a column costs one ballot for the `below` counter (accumulated by lane c for column c, as in
        // the exceedance kernel) and a shared-memory atomic only for the few lanes whose value lies inside the window,
        // so the same-address serialisation that made full-range binning slow (half of the pass) hardly occurs.
        // The ~16 instructions per column run with one warp per scheduler, i.e. latency bound: TWO epilogue
        // warpgroups take the two TMEM accumulators in turn (warps 4-7: even tile-panels, warps 8-11: odd ones), each
        // with two MMA periods per panel; both add into the same shared histograms.
        // Measured per 1,000 surrogates of config 3 (one pass, operands reused): full-range binning from this layout
        // 4.4 ms; every 32-column chunk transposed through shared memory so that a thread walks 32 surrogates of one
        // pair 2.9 ms; windows + ballot counters 2.5 ms; + the second warpgroup 1.7 ms (GEMM floor of this loop
        // order, both operands streamed from L2: 0.6 ms).
        const int wg = (warp - 4) >> 2;                         // epilogue warpgroup 0 / 1
        const int q = warp & 3;                                 // TMEM lane quadrant this warp may read
        const int te = q * 32 + lane;                           // accumulator lane = surrogate row of the panel
        const int t256 = threadIdx.x - 128;                     // 0..255 over both warpgroups
        const float nb = (float)p.n_bins;
        uint32_t it = 0;
        for (int un = blockIdx.x; un < n_units; un += gridDim.x) {
            const int f = un / p.NT, nt = un - f * p.NT;
            if (t256 < kPhPairs) {
                const int pair_t = nt * kPhPairs + t256;
                const bool pair_ok = pair_t < p.n_pairs;
                lo_s[t256] = (pair_ok && p.bin_lo) ? __ldg(p.bin_lo + (int64_t)f * p.n_pairs + pair_t) : 0.f;
                sc_s[t256] = (pair_ok && p.bin_scale) ? __ldg(p.bin_scale + (int64_t)f * p.n_pairs + pair_t) : nb;
            }
            for (int i = t256; i < kPhPairs * hp; i += 256) hist_s[i] = 0u;
            uint32_t below_r[4] = {0u, 0u, 0u, 0u};              // lane c: surrogates of this warp under the window
            asm volatile("bar.sync 1, 256;" ::: "memory");       // of column ch * 32 + c
            for (int mt = 0; mt < p.MT; ++mt, ++it) {
                if ((int)(it & 1) != wg) continue;
                const bool s_ok = mt * kPhM + te < p.n_local;
                const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
                mbar_wait(&bars->tmem_full[acc], accphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + acc * kPhN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    uint32_t re[32], im[32];
                    tmem_ld_32x32(taddr + ch * 32, re);
                    tmem_ld_32x32(taddr + kPhPairs + ch * 32, im);
                    tmem_ld_wait();
                    uint32_t mine = 0;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float a = __uint_as_float(re[c]), b = __uint_as_float(im[c]);
                        const float cv = fminf((a * a + b * b) * kZUnscaleSq, 1.0f);
                        const float x = (cv - lo_s[ch * 32 + c]) * sc_s[ch * 32 + c];
                        const uint32_t bal = __ballot_sync(0xffffffffu, s_ok && x < 0.f);
                        mine = (lane == c) ? __popc(bal) : mine;
                        if (s_ok && x >= 0.f && x < nb) atomicAdd(hist_s + (ch * 32 + c) * hp + (int)x, 1u);
                    }
                    below_r[ch] += mine;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // flush: rows of padding pairs (beyond n_pairs) hold counts of the zero rows of Z and are dropped
            // (reductions without a return value: no dependent global load per entry; only this CTA touches the
            // unit's rows during the launch, other launches are stream-ordered)
            const int rows = min(kPhPairs, p.n_pairs - nt * kPhPairs);
            uint32_t* g = p.hist + ((int64_t)f * p.n_pairs + (int64_t)nt * kPhPairs) * p.n_bins;
            for (int r = t256 >> 7; r < rows; r += 2)
                for (int b = t256 & 127; b < p.n_bins; b += 128) {
                    const uint32_t v = hist_s[r * hp + b];
                    if (v) atomicAdd(g + r * p.n_bins + b, v);
                }
            if (p.below) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const int pc = nt * kPhPairs + ch * 32 + lane;
                    if (pc < p.n_pairs && below_r[ch]) atomicAdd(p.below + (int64_t)f * p.n_pairs + pc, below_r[ch]);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// Rank selection in per-row histograms: for every row (= one (f, i, j) pair) the bin that holds the value of 0-based
// rank k - the first bin whose running count, started at below[row], exceeds k - and the count below that bin.
// bin -1: the rank lies under the window (below > k); bin n_bins: over it (the counts never reach k; below is then
// the total).  One warp per row, coalesced reads, shuffle scan.
__global__ void __launch_bounds__(256)
hist_select_kernel(const uint32_t* __restrict__ hist, int64_t n_rows, int n_bins, int k, int32_t* __restrict__ below,
                   int32_t* __restrict__ bin_out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const uint32_t* h = hist + row * n_bins;
    int run = below[row];
    if (run > k) {
        if (lane == 0) bin_out[row] = -1;
        return;
    }
    int found_bin = n_bins, found_below = -1;
    for (int b0 = 0; b0 < n_bins && found_below < 0; b0 += 32) {
        const int b = b0 + lane;
        const int v = b < n_bins ? (int)h[b] : 0;
        int inc = v;                                           // inclusive scan over the 32 bins of this step
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, b < n_bins && run + inc > k);
        if (hit) {
            const int l = __ffs(hit) - 1;
            found_bin = b0 + l;
            found_below = run + __shfl_sync(0xffffffffu, inc - v, l);
        } else {
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    if (lane == 0) {
        bin_out[row] = found_bin;
        below[row] = found_below >= 0 ? found_below : run;
    }
}

__global__ void phase_gather_kernel(const uint32_t* __restrict__ max_u, int64_t n, float* __restrict__ max_stat) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) max_stat[i] = __uint_as_float(max_u[i]);
}

struct PhaseLayout {
    int KPb, KB, S_pad, MT, n_pairs, n_pairs_pad, R_pad, NT, f_chunk, nterms;
    int64_t off_max, off_A, off_Z, total;
};

static PhaseLayout phase_layout(int L, int F, int Ne, int Nm, int64_t n_surr) {
    PhaseLayout y;
    y.nterms = L <= kPhSplitMaxL ? 3 : 1;
    y.KPb = (int)align_up(2 * (int64_t)L * y.nterms, kPhKB);
    y.KB = y.KPb / kPhKB;
    y.S_pad = (int)align_up(n_surr > 0 ? n_surr : 1, kPhM);
    y.MT = y.S_pad / kPhM;
    y.n_pairs = Ne * Nm;
    y.n_pairs_pad = (int)align_up(y.n_pairs, kPhPairs);
    y.R_pad = 2 * y.n_pairs_pad;
    y.NT = y.n_pairs_pad / kPhPairs;
    // frequencies per launch: keep the generated operands of one chunk within ~3 GB
    const int64_t per_f = ((int64_t)y.S_pad + y.R_pad) * y.KPb * 2;
    int64_t fc = (3ll << 30) / (per_f > 0 ? per_f : 1);
    fc = fc < 4 ? 4 : fc;
    fc = fc / 4 * 4;                                            // Philox emits 4 frequencies per call
    y.f_chunk = (int)(fc > F ? align_up(F, 4) : fc);
    int64_t o = 0;
    y.off_max = o; o = align_up(o + (int64_t)y.S_pad * 4, 1024);
    y.off_A = o; o = align_up(o + (int64_t)y.f_chunk * y.S_pad * y.KPb * 2, 1024);
    y.off_Z = o; o = align_up(o + (int64_t)y.f_chunk * y.R_pad * y.KPb * 2, 1024);
    y.total = o;
    return y;
}

int64_t phase_workspace_bytes(int L, int F, int Ne, int Nm, int64_t n_surr) {
    return phase_layout(L, F, Ne, Nm, n_surr).total;
}

// Operand generation + GEMM launch per frequency chunk, shared by the exceedance null and the histogram pass.
// `launch(fc, f0, mA, mB, y)` enqueues the GEMM kernel of one chunk.  `reuse`: ws2 still holds the phase panel and
// the Z rows that an earlier call generated for the SAME (ws, seed, surrogate range, frequency range) - honoured
// when the range is one chunk (otherwise only the last chunk's operands survive and everything is regenerated).
template <typename Launch>
static int phase_run(const void* ws, int L, int F, int Ne, int Nm, uint64_t seed, int64_t s_begin, int64_t n, int f_begin,
                     int f_end, void* ws2, int64_t ws2_bytes, cudaStream_t st, const PhaseLayout& y, bool reuse,
                     Launch launch) {
    CMC_REQUIRE(n < (1ll << 31) - 256, "cmc_surrogate_null: too many surrogates in one call");
    const CsdLayout cy = csd_layout(L, F, Ne, Nm);
    if (ws2_bytes < y.total) {
        set_error("cmc_surrogate_null: workspace %lld < %lld bytes", (long long)ws2_bytes, (long long)y.total);
        return CMC_EWORKSPACE;
    }
    CMC_REQUIRE((reinterpret_cast<uintptr_t>(ws2) & 255) == 0, "cmc_surrogate_null: workspace must be 256-byte aligned");
    CMC_REQUIRE((size_t)L * 8 <= 48 * 1024, "cmc_surrogate_null: L=%d too long (2L <= 12288)", L);
    const uint32_t* table;
    int rc = get_phase_table(&table);
    if (rc) return rc;
    const unsigned char* w = static_cast<const unsigned char*>(ws);
    unsigned char* w2 = static_cast<unsigned char*>(ws2);
    __half* A = reinterpret_cast<__half*>(w2 + y.off_A);
    __half* Z = reinterpret_cast<__half*>(w2 + y.off_Z);
    const bool keep = reuse && f_end - f_begin <= y.f_chunk;
    for (int f0 = f_begin; f0 < f_end; f0 += y.f_chunk) {
        const int fc = f_end - f0 < y.f_chunk ? f_end - f0 : y.f_chunk;
        if (!keep && y.n_pairs_pad != y.n_pairs) {
            rc = check_cuda(cudaMemsetAsync(Z, 0, (size_t)fc * y.R_pad * y.KPb * 2, st), "memset(Z)");
            if (rc) return rc;
        }
        // Philox counters use GLOBAL frequency groups of 4; a chunk may start inside a group
        if (!keep) phase_gen_kernel<<<dim3((y.KPb / 2 + 255) / 256, y.S_pad, ((f0 + fc - 1) >> 2) - (f0 >> 2) + 1), 256, 0, st>>>(
            table, seed, s_begin, (int)n, y.S_pad, L, y.nterms, fc, f0, y.KPb, A);
        CMC_CHECK_LAUNCH("phase_gen_kernel");
        if (!keep) (y.nterms == 3 ? z_gen_kernel<3> : z_gen_kernel<1>)<<<dim3(fc, Ne), 256, (size_t)L * 8, st>>>(
            reinterpret_cast<const float*>(w + cy.off_ahi) + (int64_t)f0 * cy.MT * kTileM * cy.KP,
            reinterpret_cast<const float*>(w + cy.off_alo) + (int64_t)f0 * cy.MT * kTileM * cy.KP,
            reinterpret_cast<const float*>(w + cy.off_bhi) + (int64_t)f0 * cy.NT * kTileN * cy.KP,
            reinterpret_cast<const float*>(w + cy.off_blo) + (int64_t)f0 * cy.NT * kTileN * cy.KP,
            reinterpret_cast<const float*>(w + cy.off_pxx) + (int64_t)f0 * Ne,
            reinterpret_cast<const float*>(w + cy.off_pyy) + (int64_t)f0 * Nm, L, Ne, Nm, cy.MT, cy.NT, cy.KP, cy.KP,
            y.KPb, y.R_pad, Z);
        CMC_CHECK_LAUNCH("z_gen_kernel");
        CUtensorMap mA, mB;
        if ((rc = make_kmajor_map(&mA, A, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, y.KPb, (int64_t)fc * y.S_pad, kPhM))) return rc;
        if ((rc = make_kmajor_map(&mB, Z, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, y.KPb, (int64_t)fc * y.R_pad, kPhN))) return rc;
        if ((rc = launch(fc, f0, mA, mB))) return rc;
    }
    return CMC_OK;
}

static int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

int phase_surrogate_null(const void* ws, int L, int F, int Ne, int Nm, uint64_t seed, int64_t s_begin, int64_t s_end,
                         int f_begin, int f_end, const float* coh_obs, uint32_t* exceed, float* max_stat, void* ws2,
                         int64_t ws2_bytes, cudaStream_t st) {
    const int64_t n = s_end - s_begin;
    const PhaseLayout y = phase_layout(L, F, Ne, Nm, n);
    const size_t tail_bytes = kPhPairs * 8 + sizeof(PhaseBarriers) + 16;
    size_t smem = 1024 + (size_t)y.KB * kPhABytes + kPhStages * kPhBBytes + tail_bytes;
    const bool stream_a = smem > 227 * 1024;                    // K > 512: the panel no longer fits next to the ring
    if (stream_a) smem = 1024 + (size_t)kPhStagesStream * (kPhABytes + kPhBBytes) + tail_bytes;
    auto gemm = stream_a ? phase_gemm_kernel<true> : phase_gemm_kernel<false>;
    int rc = ensure_smem_attr(reinterpret_cast<const void*>(gemm), smem);
    if (rc) return rc;
    if (ws2_bytes < y.total) {
        set_error("cmc_surrogate_null: workspace %lld < %lld bytes", (long long)ws2_bytes, (long long)y.total);
        return CMC_EWORKSPACE;
    }
    uint32_t* max_u = reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(ws2) + y.off_max);
    rc = check_cuda(cudaMemsetAsync(max_u, 0, (size_t)y.S_pad * 4, st), "memset(max_u)");
    if (rc) return rc;
    const int sms = sm_count();
    rc = phase_run(ws, L, F, Ne, Nm, seed, s_begin, n, f_begin, f_end, ws2, ws2_bytes, st, y, false,
                   [&](int fc, int f0, const CUtensorMap& mA, const CUtensorMap& mB) -> int {
                       PhaseParams p{};
                       p.F = fc; p.MT = y.MT; p.NT = y.NT; p.KB = y.KB; p.n_local = (int)n; p.n_pairs = y.n_pairs;
                       p.S_pad = y.S_pad; p.R_pad = y.R_pad;
                       p.coh_obs = coh_obs + (int64_t)f0 * y.n_pairs;
                       p.exceed = exceed + (int64_t)f0 * y.n_pairs;
                       p.max_u = max_u;
                       const int n_panels = fc * y.MT;
                       gemm<<<n_panels < sms ? n_panels : sms, kPhThreads, smem, st>>>(mA, mB, p);
                       CMC_CHECK_LAUNCH("phase_gemm_kernel");
                       return CMC_OK;
                   });
    if (rc) return rc;
    phase_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(max_u, n, max_stat);
    CMC_CHECK_LAUNCH("phase_gather_kernel");
    return CMC_OK;
}

int phase_surrogate_hist(const void* ws, int L, int F, int Ne, int Nm, uint64_t seed, int64_t s_begin, int64_t s_end,
                         int f_begin, int f_end, int n_bins, const float* bin_lo, const float* bin_scale,
                         uint32_t* hist, uint32_t* below, void* ws2, int64_t ws2_bytes, bool reuse, cudaStream_t st) {
    const int64_t n = s_end - s_begin;
    CMC_REQUIRE(n_bins >= 2 && n_bins <= kPhHistMaxBins, "cmc_surrogate_null_hist: n_bins must be in [2, %d]",
                kPhHistMaxBins);
    const PhaseLayout y = phase_layout(L, F, Ne, Nm, n);
    const size_t smem = 1024 + (size_t)kPhHistStages * (kPhABytes + kPhBBytes) + (size_t)kPhPairs * (n_bins + 1) * 4 +
                        kPhPairs * 8 + sizeof(PhaseBarriers) + 16;
    int rc = ensure_smem_attr(reinterpret_cast<const void*>(phase_hist_kernel), smem);
    if (rc) return rc;
    const int sms = sm_count();
    return phase_run(ws, L, F, Ne, Nm, seed, s_begin, n, f_begin, f_end, ws2, ws2_bytes, st, y, reuse,
                     [&](int fc, int f0, const CUtensorMap& mA, const CUtensorMap& mB) -> int {
                         PhaseHistParams p{};
                         p.F = fc; p.MT = y.MT; p.NT = y.NT; p.KB = y.KB; p.n_local = (int)n; p.n_pairs = y.n_pairs;
                         p.S_pad = y.S_pad; p.R_pad = y.R_pad; p.n_bins = n_bins;
                         p.below = below ? below + (int64_t)f0 * y.n_pairs : nullptr;
                         p.bin_lo = bin_lo ? bin_lo + (int64_t)f0 * y.n_pairs : nullptr;
                         p.bin_scale = bin_scale ? bin_scale + (int64_t)f0 * y.n_pairs : nullptr;
                         p.hist = hist + (int64_t)f0 * y.n_pairs * n_bins;
                         const int n_units = fc * y.NT;
                         phase_hist_kernel<<<n_units < sms ? n_units : sms, kPhHistThreads, smem, st>>>(mA, mB, p);
                         CMC_CHECK_LAUNCH("phase_hist_kernel");
                         return CMC_OK;
                     });
}

}  // namespace cmc

extern "C" CMC_API int cmc_hist_select(const uint32_t* hist, int64_t n_rows, int n_bins, int k, int32_t* below,
                                       int32_t* bin_out, void* stream) {
    using namespace cmc;
    CMC_REQUIRE(hist && below && bin_out, "cmc_hist_select: null pointer");
    CMC_REQUIRE(n_rows >= 0 && n_bins >= 1 && k >= 0, "cmc_hist_select: bad shape");
    if (n_rows == 0) return CMC_OK;
    hist_select_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(hist, n_rows, n_bins, k,
                                                                                                  below, bin_out);
    CMC_CHECK_LAUNCH("hist_select_kernel");
    return CMC_OK;
}

// host copy of the kernel's FP16 phase table as float pairs (cos, sin): diagnostics of the operand rounding
extern "C" CMC_API int cmc_phase_table(float* out_host /* [4096][2] */) {
    if (!out_host) return CMC_EINVAL;
    for (int a = 0; a < cmc::kPhaseN; ++a) {
        uint32_t hi, lo;
        cmc::host_phase_entry(a, &hi, &lo);
        out_host[2 * a] = cmc::host_f16_to_float((uint16_t)(hi & 0xFFFFu));
        out_host[2 * a + 1] = cmc::host_f16_to_float((uint16_t)(hi >> 16));
    }
    return CMC_OK;
}

// K2w: per-window multitaper magnitude-squared coherence for all (EEG, EMG) pairs, with the
// leave-one-taper-out jackknife CI and the independence-threshold mask fused in.
//
// Replaces signal_features.py:750-796 (PSD / CSD sums, raw coherence, threshold mask) and
// jackknife_coherence_and_ci (:484-578).  The average here runs over only K (~5) tapers, so
// this is an outer product with a transcendental epilogue, not a GEMM: CUDA cores, one CTA
// per (window, frequency chunk), spectra of one (window, frequency) staged in shared memory,
// outputs written fully coalesced (the pair index i * Nm + j is the contiguous output index).
// Its floor is the HBM write of the (W, F, Ne, Nm) outputs (13 bytes per pair); measured, the
// jackknife sits on the MUFU and FMA pipes (17 MUFU per pair) at ~2.5x that floor.
#include "common.cuh"

namespace cmc {

constexpr int kMscThreads = 256;
constexpr int kFreqPerBlock = 4;

// Build-time variants of the jackknife path (scripts/k2w_variants.sh builds and scripts/time_msc_windows.py times them;
// profiles/r02y_k2w_variants.txt).  Defaults = the fastest measured combination.
#ifndef CMC_K2W_RCP_MUFU
#define CMC_K2W_RCP_MUFU 1      // 1: reciprocals on MUFU.RCP, 0: exponent-flip estimate + three Newton steps on FFMA2
#endif
#ifndef CMC_K2W_Y_SMEM
#define CMC_K2W_Y_SMEM 1        // 1: the EMG operands are re-read from shared memory for every EEG row (fewer registers)
#endif
#ifndef CMC_K2W_MINB            // CTAs per SM the register allocation is held to
#define CMC_K2W_MINB(K) ((K) <= 5 ? 4 : (K) <= 8 ? 3 : 1)
#define CMC_K2W_MINB_MAXEMG(K) ((K) <= 5 ? 3 : (K) <= 8 ? 2 : 1)
#else
#define CMC_K2W_MINB_MAXEMG(K) CMC_K2W_MINB(K)
#endif

// The transcendental epilogue dominates this kernel (6 Fisher transforms + 2 inverse transforms per output), so it
// is written on the bare MUFU approximations with flush-to-zero (no denormal can occur: the arguments are clamped
// into [6e-8, 3.4e7]) - the default __logf / __expf / sqrtf expansions carry a denormal rescue and a Newton step per
// call, which doubled the instruction count of an output:
//   z'(c)  = lg2((1 + c) / (1 - c))                 (= 2 z / ln 2)           LG2
//   z^-1   = tanh(z)^2,  tanh(z) = 1 - 2 / (2^z' + 1)                        EX2
// z errors ~3e-7 relative, far inside the 1e-4 gate on coherence / CI bounds.
__device__ __forceinline__ float mufu_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// The jackknife works in units of z' = 2 z / ln 2 = lg2((1 + c) / (1 - c)): exp(2 z) = 2^z', so neither direction
// needs a multiplication by a constant, and the variance / CI half width scale along (t_crit * se is linear in z).
// The lower clip of the reference (c >= 1e-10) moves z by < 2e-10 and is dropped; the upper clip is folded into the
// clamp of the replicate coherence (1 - 2^-24 instead of 1: 6e-8 on a coherence).  fisher_z2x2 / inv_fisher_z2x2 below.
__device__ __forceinline__ float msc_ratio(float re, float im, float sxx, float syy) {
    // clip(|sxy|^2 / max(sxx * syy, tiny), 0, 1) evaluated as |sxy / sqrt(sxx) / sqrt(syy)|^2 so that
    // the product of the auto-spectra cannot overflow / underflow in float32
    if (!(sxx > 0.f) || !(syy > 0.f)) return 0.f;
    const float r = rsqrtf(sxx) * rsqrtf(syy);
    const float a = re * r, b = im * r;
    return fminf(a * a + b * b, 1.0f);
}

// raw multitaper coherence of one pair (signal_features.py:750-796): no jackknife
template <int K>
__device__ __forceinline__ float pair_coherence(const float2 (&x)[K], const float2 (&y)[K]) {
    float sxx = 0.f, syy = 0.f, sre = 0.f, sim = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        // conj(x) * y
        const float2 c = make_float2(x[k].x * y[k].x + x[k].y * y[k].y, x[k].x * y[k].y - x[k].y * y[k].x);
        const float px = x[k].x * x[k].x + x[k].y * x[k].y;
        const float py = y[k].x * y[k].x + y[k].y * y[k].y;
        sxx += px;
        syy += py;
        sre += c.x;
        sim += c.y;
    }
    return msc_ratio(sre, sim, sxx, syy);
}

// ---- packed pairs of float32: FFMA2 / FADD2 / FMUL2 of sm_100 process two IEEE float32 lanes per issue slot.  The
// jackknife is issue bound, so every thread works on TWO (EEG, EMG) pairs at once: lane .lo = pair A, lane .hi = pair B.
// Each lane is rounded exactly like the scalar instruction, so a pair gets the same bits whichever lane (and whichever
// kernel) computes it - the fused EMG-argmax kernel returns exactly the values of the unfused one.
typedef unsigned long long f32x2;
constexpr f32x2 kOne2 = 0x3f8000003f800000ull;       // (1, 1)
constexpr f32x2 kMinusTwo2 = 0xc0000000c0000000ull;  // (-2, -2)
constexpr f32x2 kMinusOne2 = 0xbf800000bf800000ull;  // (-1, -1)
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// 1 / d of both lanes from -d (which the callers get for free).  Default: MUFU.RCP.  The alternative keeps the
// reciprocals off the MUFU pipe - exponent-flip estimate (5 % off) and three Newton steps y += y (1 - d y) on FFMA2, 2
// integer + 6 packed instructions for both lanes, relative error 1.2e-7 over d in [2^-24, 2^36] - and measured slower:
// the packed instructions only run on the heavy FMA pipe, which is as busy (58 %) as the MUFU pipe (63 %).
__device__ __forceinline__ f32x2 rcp2_of_neg(f32x2 nd) {
    float n0, n1;
    upk2(nd, n0, n1);
#if CMC_K2W_RCP_MUFU
    float r0, r1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(-n0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(-n1));
    return pk2(r0, r1);
#endif
    f32x2 y = pk2(__uint_as_float(0xFEF311C7u - __float_as_uint(n0)), __uint_as_float(0xFEF311C7u - __float_as_uint(n1)));
#pragma unroll
    for (int it = 0; it < 3; ++it) y = fma2(y, fma2(nd, y, kOne2), y);
    return y;
}
// z' = lg2((1 + c) / (1 - c)) of both lanes, c in [0, 1 - 2^-24]
__device__ __forceinline__ f32x2 fisher_z2x2(f32x2 c) {
    float q0, q1;
    upk2(mul2(add2(c, kOne2), rcp2_of_neg(sub2(c, kOne2))), q0, q1);
    return pk2(mufu_lg2(q0), mufu_lg2(q1));
}
// tanh(z)^2 of both lanes for z = z' ln 2 / 2: tanh(z) = 1 - 2 / (2^z' + 1); |z'| is capped where tanh saturates
__device__ __forceinline__ f32x2 inv_fisher_z2x2(f32x2 z2) {
    float a, b;
    upk2(z2, a, b);
    const f32x2 e = pk2(mufu_ex2(fminf(fmaxf(a, -34.6f), 34.6f)), mufu_ex2(fminf(fmaxf(b, -34.6f), 34.6f)));
    const f32x2 t = fma2(kMinusTwo2, rcp2_of_neg(sub2(kMinusOne2, e)), kOne2);
    return mul2(t, t);
}

// Spectra of one (window, frequency) as the jackknife consumes them, staged in shared memory so that every packed
// operand is ONE aligned 64-bit load (building a pair from two scalar registers costs a MOV per operand and use):
//   EEG side, both lanes equal (a broadcast load feeds both pairs of a thread):
//     x4[k * Ne + i] = (re, re, im, im),  xn2[k * Ne + i] = (-im, -im),  rx2[k * Ne + i] = (r, r)
//   EMG side, planes of floats with an even row length Nmp: a thread works on the ADJACENT columns jA = 2 c (lane .lo)
//   and jA + 1 (lane .hi), so (v[jA], v[jA + 1]) is the packed operand - and the two outputs are one 8-byte store:
//     yre, yim, ry      [k * Nmp + j]     (column Nm of an odd Nm is zero: that lane is never stored)
// r = 1 / sqrt(sum_{m != k} |.|^2) are the leave-one-taper-out factors (0 for a silent channel): the auto-spectra part
// of every replicate depends on one channel only, so it is computed once per (window, bin, channel), not once per pair.
// Without the jackknife: plain float2 rows sx[k * Ne + i], sy[k * Nm + j].
template <int K>
struct MscSmem {
    float4* x4;
    float2 *xn2, *rx2;
    float *yre, *yim, *ry;
    float2 *sx, *sy;
    int Nmp;
    __device__ __forceinline__ MscSmem(unsigned char* raw, int Ne, int Nm) {
        Nmp = (Nm + 1) & ~1;
        x4 = reinterpret_cast<float4*>(raw);
        xn2 = reinterpret_cast<float2*>(x4 + K * Ne);
        rx2 = xn2 + K * Ne;
        yre = reinterpret_cast<float*>(rx2 + K * Ne);
        yim = yre + K * Nmp;
        ry = yim + K * Nmp;
        sx = reinterpret_cast<float2*>(raw);
        sy = sx + K * Ne;
    }
};
static inline size_t msc_smem_bytes(int K, int Ne, int Nm, bool jk) {
    const int Nmp = (Nm + 1) & ~1;
    return jk ? (size_t)K * ((sizeof(float4) + 2 * sizeof(float2)) * Ne + 3 * sizeof(float) * Nmp)
              : sizeof(float2) * K * (Ne + Nm);
}

template <int K>
struct EmgPair {
    f32x2 re[K], im[K], r[K];
    __device__ __forceinline__ void load(const MscSmem<K>& sm, int jA) {      // jA even
#pragma unroll
        for (int k = 0; k < K; ++k) {
            re[k] = *reinterpret_cast<const f32x2*>(sm.yre + k * sm.Nmp + jA);
            im[k] = *reinterpret_cast<const f32x2*>(sm.yim + k * sm.Nmp + jA);
            r[k] = *reinterpret_cast<const f32x2*>(sm.ry + k * sm.Nmp + jA);
        }
    }
};

struct PairStats2 {
    f32x2 coh, lo, hi;
};

// Jackknife statistics of two pairs (signal_features.py:484-578): leave-one-out sums as prefix + suffix (no
// cancellation, unlike total - term), Fisher z in lg2 units, CI = tanh^2(z(mean) -/+ t_crit * se).
template <int K>
__device__ __forceinline__ PairStats2 pair_stats_jk2(const f32x2 (&xre)[K], const f32x2 (&xim)[K],
                                                     const f32x2 (&xnim)[K], const f32x2 (&rx)[K],
                                                     const EmgPair<K>& y, f32x2 t_crit2) {
    f32x2 cre[K], cim[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {                    // conj(x) * y
        cre[k] = fma2(xim[k], y.im[k], mul2(xre[k], y.re[k]));
        cim[k] = fma2(xnim[k], y.re[k], mul2(xre[k], y.im[k]));
    }
    f32x2 pre_re[K], pre_im[K];
    pre_re[0] = pre_im[0] = 0ull;
#pragma unroll
    for (int k = 1; k < K; ++k) {
        pre_re[k] = k == 1 ? cre[0] : add2(pre_re[k - 1], cre[k - 1]);
        pre_im[k] = k == 1 ? cim[0] : add2(pre_im[k - 1], cim[k - 1]);
    }
    f32x2 z[K];
    f32x2 csum = 0ull, zsum = 0ull, a = 0ull, b = 0ull;
#pragma unroll
    for (int k = K - 1; k >= 0; --k) {
        const f32x2 r = mul2(rx[k], y.r[k]);
        const f32x2 sre = k == K - 1 ? pre_re[k] : (k == 0 ? a : add2(pre_re[k], a));
        const f32x2 sim = k == K - 1 ? pre_im[k] : (k == 0 ? b : add2(pre_im[k], b));
        const f32x2 u = mul2(sre, r), v = mul2(sim, r);
        float c0, c1;
        upk2(fma2(u, u, mul2(v, v)), c0, c1);
        const f32x2 ck = pk2(fminf(c0, 0.99999994f), fminf(c1, 0.99999994f));
        z[k] = fisher_z2x2(ck);
        csum = k == K - 1 ? ck : add2(csum, ck);
        zsum = k == K - 1 ? z[k] : add2(zsum, z[k]);
        if (k > 0) {
            a = k == K - 1 ? cre[k] : add2(a, cre[k]);
            b = k == K - 1 ? cim[k] : add2(b, cim[k]);
        }
    }
    const f32x2 inv_k = pk2(1.0f / K, 1.0f / K);
    const f32x2 mean = mul2(csum, inv_k);                      // every ck is in [0, 1 - 2^-24]
    const f32x2 zbar = mul2(zsum, inv_k);
    f32x2 ss = 0ull;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const f32x2 d = sub2(z[k], zbar);
        ss = k == 0 ? mul2(d, d) : fma2(d, d, ss);
    }
    float s0, s1;
    upk2(mul2(ss, pk2((float)(K - 1) / (float)K, (float)(K - 1) / (float)K)), s0, s1);
    const f32x2 hw = mul2(t_crit2, pk2(mufu_sqrt(s0), mufu_sqrt(s1)));
    const f32x2 zc = fisher_z2x2(mean);
    float l0, l1, h0, h1, m0, m1;
    upk2(inv_fisher_z2x2(sub2(zc, hw)), l0, l1);
    upk2(inv_fisher_z2x2(add2(zc, hw)), h0, h1);
    upk2(mean, m0, m1);
    PairStats2 out;
    out.coh = mean;
    out.lo = pk2(fminf(l0, m0), fminf(l1, m1));
    out.hi = pk2(fmaxf(h0, m0), fmaxf(h1, m1));
    return out;
}

template <int K, bool JK>
__device__ __forceinline__ void stage_spectra(const MscSmem<K>& sm, const float2* __restrict__ X,
                                              const float2* __restrict__ Y, int w, int f, int F, int Ne,
                                              int Nm, int64_t ldx, int64_t ldy) {
    for (int q = threadIdx.x; q < K * (Ne + Nm); q += blockDim.x) {
        const int k = q / (Ne + Nm), c = q % (Ne + Nm);
        if (c < Ne) {
            const float2 v = __ldg(X + ((int64_t)(w * K + k) * F + f) * ldx + c);
            if (JK) {
                sm.x4[k * Ne + c] = make_float4(v.x, v.x, v.y, v.y);
                sm.xn2[k * Ne + c] = make_float2(-v.y, -v.y);
            } else {
                sm.sx[k * Ne + c] = v;
            }
        } else {
            const int j = c - Ne;
            const float2 v = __ldg(Y + ((int64_t)(w * K + k) * F + f) * ldy + j);
            if (JK) {
                sm.yre[k * sm.Nmp + j] = v.x;
                sm.yim[k * sm.Nmp + j] = v.y;
                if (j + 1 == Nm && (Nm & 1))
                    sm.yre[k * sm.Nmp + Nm] = sm.yim[k * sm.Nmp + Nm] = sm.ry[k * sm.Nmp + Nm] = 0.f;
            } else {
                sm.sy[k * Nm + j] = v;
            }
        }
    }
}

template <int K>
__device__ __forceinline__ void stage_loo_roots(const MscSmem<K>& sm, int Ne, int Nm) {
    for (int c = threadIdx.x; c < Ne + Nm; c += blockDim.x) {
        const bool eeg = c < Ne;
        const int cc = eeg ? c : c - Ne;
        float pw[K], pre[K];
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float re, im;
            if (eeg) {
                const float4 t = sm.x4[k * Ne + cc];
                re = t.x;
                im = t.z;
            } else {
                re = sm.yre[k * sm.Nmp + cc];
                im = sm.yim[k * sm.Nmp + cc];
            }
            pw[k] = re * re + im * im;
            pre[k] = acc;
            acc += pw[k];
        }
        acc = 0.f;
#pragma unroll
        for (int k = K - 1; k >= 0; --k) {
            const float loo = pre[k] + acc;
            const float r = loo > 0.f ? rsqrtf(loo) : 0.f;
            if (eeg) sm.rx2[k * Ne + cc] = make_float2(r, r);
            else sm.ry[k * sm.Nmp + cc] = r;
            acc += pw[k];
        }
    }
}

// EEG row i of the staged (window, frequency): both lanes equal
template <int K>
__device__ __forceinline__ void load_eeg_row(const MscSmem<K>& sm, int Ne, int i, f32x2 (&xre)[K], f32x2 (&xim)[K],
                                             f32x2 (&xnim)[K], f32x2 (&rx)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const ulonglong2 v = reinterpret_cast<const ulonglong2*>(sm.x4)[k * Ne + i];
        xre[k] = v.x;
        xim[k] = v.y;
        xnim[k] = reinterpret_cast<const f32x2*>(sm.xn2)[k * Ne + i];
        rx[k] = reinterpret_cast<const f32x2*>(sm.rx2)[k * Ne + i];
    }
}

template <int K, bool JK>
__global__ void __launch_bounds__(kMscThreads, CMC_K2W_MINB(K))
msc_windows_kernel(const float2* __restrict__ X, const float2* __restrict__ Y, int F, int Ne, int Nm,
                   int64_t ldx, int64_t ldy, const uint8_t* __restrict__ window_mask, float t_crit,
                   float it_threshold, int vec2, float* __restrict__ coh, float* __restrict__ ci_lo,
                   float* __restrict__ ci_hi, uint8_t* __restrict__ significant) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const MscSmem<K> sm(smem_raw, Ne, Nm);
    const int w = blockIdx.y;
    if (window_mask && !window_mask[w]) return;
    const int n_pairs = Ne * Nm;
    const int f_end = min(F, (int)(blockIdx.x + 1) * kFreqPerBlock);
    const f32x2 t_crit2 = pk2(t_crit, t_crit);
    for (int f = blockIdx.x * kFreqPerBlock; f < f_end; ++f) {
        __syncthreads();
        stage_spectra<K, JK>(sm, X, Y, w, f, F, Ne, Nm, ldx, ldy);
        __syncthreads();
        if (JK) {
            stage_loo_roots<K>(sm, Ne, Nm);
            __syncthreads();
        }
        const int64_t obase = ((int64_t)w * F + f) * n_pairs;
        if (JK) {
            float* const coh_f = coh + obase;
            float* const lo_f = ci_lo + obase;
            float* const hi_f = ci_hi + obase;
            uint8_t* const sig_f = significant ? significant + obase : nullptr;
            // outputs of the columns jA (lane .lo) and jA + 1 (lane .hi) of row i, o = i * Nm + jA
            auto store = [&](int o, bool has_b, const PairStats2& st) {
                float c0, c1, l0, l1, h0, h1;
                upk2(st.coh, c0, c1);
                upk2(st.lo, l0, l1);
                upk2(st.hi, h0, h1);
                if (vec2) {                    // Nm even and every output 8-byte aligned: one store per array
                    *reinterpret_cast<float2*>(coh_f + o) = make_float2(c0, c1);
                    *reinterpret_cast<float2*>(lo_f + o) = make_float2(l0, l1);
                    *reinterpret_cast<float2*>(hi_f + o) = make_float2(h0, h1);
                    if (sig_f)
                        *reinterpret_cast<uchar2*>(sig_f + o) =
                            make_uchar2(c0 > it_threshold ? 1 : 0, c1 > it_threshold ? 1 : 0);
                } else {
                    coh_f[o] = c0;
                    lo_f[o] = l0;
                    hi_f[o] = h0;
                    if (sig_f) sig_f[o] = c0 > it_threshold ? 1 : 0;
                    if (has_b) {
                        coh_f[o + 1] = c1;
                        lo_f[o + 1] = l1;
                        hi_f[o + 1] = h1;
                        if (sig_f) sig_f[o + 1] = c1 > it_threshold ? 1 : 0;
                    }
                }
            };
            const int cols = (Nm + 1) >> 1;                      // column pairs
            if (kMscThreads % cols == 0) {
                // a thread owns one column pair (its spectra and factors stay in registers) and walks the EEG rows,
                // one broadcast load per row
                const int jA = 2 * (threadIdx.x % cols);
                EmgPair<K> y;
                y.load(sm, jA);
                for (int i = threadIdx.x / cols; i < Ne; i += kMscThreads / cols) {
#if CMC_K2W_Y_SMEM
                    y.load(sm, jA);
#endif
                    f32x2 xre[K], xim[K], xnim[K], rx[K];
                    load_eeg_row<K>(sm, Ne, i, xre, xim, xnim, rx);
                    store(i * Nm + jA, jA + 1 < Nm, pair_stats_jk2<K>(xre, xim, xnim, rx, y, t_crit2));
                }
            } else {
                for (int q = threadIdx.x; q < Ne * cols; q += kMscThreads) {
                    const int i = q / cols, jA = 2 * (q - i * cols);
                    EmgPair<K> y;
                    y.load(sm, jA);
                    f32x2 xre[K], xim[K], xnim[K], rx[K];
                    load_eeg_row<K>(sm, Ne, i, xre, xim, xnim, rx);
                    store(i * Nm + jA, jA + 1 < Nm, pair_stats_jk2<K>(xre, xim, xnim, rx, y, t_crit2));
                }
            }
        } else {
            auto emit = [&](int p, const float2 (&x)[K], const float2 (&y)[K]) {
                const float c = pair_coherence<K>(x, y);
                coh[obase + p] = c;
                if (significant) significant[obase + p] = c > it_threshold ? 1 : 0;
            };
            if (kMscThreads % Nm == 0) {
                // the EMG column of a thread is fixed (p advances by a multiple of Nm): keep its K spectra in registers
                const int j = threadIdx.x % Nm;
                float2 y[K];
#pragma unroll
                for (int k = 0; k < K; ++k) y[k] = sm.sy[k * Nm + j];
                for (int i = threadIdx.x / Nm; i < Ne; i += kMscThreads / Nm) {
                    float2 x[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) x[k] = sm.sx[k * Ne + i];
                    emit(i * Nm + j, x, y);
                }
            } else {
                for (int p = threadIdx.x; p < n_pairs; p += kMscThreads) {
                    const int i = p / Nm, j = p - i * Nm;
                    float2 x[K], y[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        x[k] = sm.sx[k * Ne + i];
                        y[k] = sm.sy[k * Nm + j];
                    }
                    emit(p, x, y);
                }
            }
        }
    }
}

// fused EMG-argmax variant: one warp per (window, frequency, EEG channel); with the jackknife a lane takes the EMG
// columns 2 c and 2 c + 1 together (c = lane, lane + 32, ...)
template <int K, bool JK>
__global__ void __launch_bounds__(kMscThreads, CMC_K2W_MINB_MAXEMG(K))
msc_windows_maxemg_kernel(const float2* __restrict__ X, const float2* __restrict__ Y, int F, int Ne, int Nm,
                          int64_t ldx, int64_t ldy, const uint8_t* __restrict__ window_mask, float t_crit,
                          float it_threshold, int zero_nonsig, float* __restrict__ out_coh,
                          float* __restrict__ out_lo, float* __restrict__ out_hi,
                          int32_t* __restrict__ out_arg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const MscSmem<K> sm(smem_raw, Ne, Nm);
    const int w = blockIdx.y;
    if (window_mask && !window_mask[w]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f_end = min(F, (int)(blockIdx.x + 1) * kFreqPerBlock);
    const f32x2 t_crit2 = pk2(t_crit, t_crit);
    const int cols = (Nm + 1) >> 1;
    for (int f = blockIdx.x * kFreqPerBlock; f < f_end; ++f) {
        __syncthreads();
        stage_spectra<K, JK>(sm, X, Y, w, f, F, Ne, Nm, ldx, ldy);
        __syncthreads();
        if (JK) {
            stage_loo_roots<K>(sm, Ne, Nm);
            __syncthreads();
        }
        // up to 64 EMG channels the column pair of a lane is the same for every EEG row: loaded once
        const bool y_fixed = JK && cols <= 32 && !CMC_K2W_Y_SMEM;
        EmgPair<K> yf;
        if (y_fixed) yf.load(sm, 2 * min(lane, cols - 1));
        for (int i = warp; i < Ne; i += kMscThreads / 32) {
            float best = -1.f, blo = 0.f, bhi = 0.f;
            int bj = 0x7fffffff;
            auto take = [&](float c, float lo, float hi, int j) {
                float val = c;
                if (zero_nonsig && !(c > it_threshold)) val = 0.f;
                if (val > best) { best = val; blo = lo; bhi = hi; bj = j; }
            };
            if (JK) {
                f32x2 xre[K], xim[K], xnim[K], rx[K];
                load_eeg_row<K>(sm, Ne, i, xre, xim, xnim, rx);
                for (int c = lane; c < cols; c += 32) {
                    if (!y_fixed) yf.load(sm, 2 * c);
                    const PairStats2 st = pair_stats_jk2<K>(xre, xim, xnim, rx, yf, t_crit2);
                    float c0, c1, l0, l1, h0, h1;
                    upk2(st.coh, c0, c1);
                    upk2(st.lo, l0, l1);
                    upk2(st.hi, h0, h1);
                    take(c0, l0, h0, 2 * c);
                    if (2 * c + 1 < Nm) take(c1, l1, h1, 2 * c + 1);
                }
            } else {
                float2 x[K];
#pragma unroll
                for (int k = 0; k < K; ++k) x[k] = sm.sx[k * Ne + i];
                for (int j = lane; j < Nm; j += 32) {
                    float2 y[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) y[k] = sm.sy[k * Nm + j];
                    take(pair_coherence<K>(x, y), 0.f, 0.f, j);
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, off);
                const float ol = __shfl_xor_sync(0xffffffffu, blo, off);
                const float oh = __shfl_xor_sync(0xffffffffu, bhi, off);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
                if (ov > best || (ov == best && oj < bj)) { best = ov; blo = ol; bhi = oh; bj = oj; }
            }
            if (lane == 0) {
                const int64_t o = ((int64_t)w * F + f) * Ne + i;
                out_coh[o] = best;
                if (JK) { out_lo[o] = blo; out_hi[o] = bhi; }
                if (out_arg) out_arg[o] = bj;
            }
        }
    }
}

template <int K>
static int launch_msc(bool maxemg, const float2* X, const float2* Y, int W, int F, int Ne, int Nm,
                      int64_t ldx, int64_t ldy, const uint8_t* mask, int jk, float t_crit, float it,
                      int zero_nonsig, float* o0, float* o1, float* o2, void* o3, cudaStream_t st) {
    dim3 grid((F + kFreqPerBlock - 1) / kFreqPerBlock, W);
    const size_t smem = msc_smem_bytes(K, Ne, Nm, jk != 0);
    const void* fn = maxemg ? (jk ? (const void*)msc_windows_maxemg_kernel<K, true> : (const void*)msc_windows_maxemg_kernel<K, false>)
                            : (jk ? (const void*)msc_windows_kernel<K, true> : (const void*)msc_windows_kernel<K, false>);
    if (smem > 48 * 1024) {
        const int rc = ensure_smem_attr(fn, smem);
        if (rc != CMC_OK) return rc;
    }
    if (!maxemg) {
        // two adjacent outputs go out as one 8-byte store when every one of them is 8-byte aligned
        const int vec2 = !(Nm & 1) && !(((uintptr_t)o0 | (uintptr_t)o1 | (uintptr_t)o2) & 7) && !((uintptr_t)o3 & 1);
        if (jk)
            msc_windows_kernel<K, true><<<grid, kMscThreads, smem, st>>>(
                X, Y, F, Ne, Nm, ldx, ldy, mask, t_crit, it, vec2, o0, o1, o2, (uint8_t*)o3);
        else
            msc_windows_kernel<K, false><<<grid, kMscThreads, smem, st>>>(
                X, Y, F, Ne, Nm, ldx, ldy, mask, t_crit, it, vec2, o0, o1, o2, (uint8_t*)o3);
    } else {
        if (jk)
            msc_windows_maxemg_kernel<K, true><<<grid, kMscThreads, smem, st>>>(
                X, Y, F, Ne, Nm, ldx, ldy, mask, t_crit, it, zero_nonsig, o0, o1, o2, (int32_t*)o3);
        else
            msc_windows_maxemg_kernel<K, false><<<grid, kMscThreads, smem, st>>>(
                X, Y, F, Ne, Nm, ldx, ldy, mask, t_crit, it, zero_nonsig, o0, o1, o2, (int32_t*)o3);
    }
    CMC_CHECK_LAUNCH("msc_windows_kernel");
    return CMC_OK;
}

static int dispatch_msc(bool maxemg, const float* X, const float* Y, int W, int K, int F, int Ne, int Nm,
                        int64_t ldx, int64_t ldy, const uint8_t* mask, int jk, float t_crit, float it,
                        int zero_nonsig, float* o0, float* o1, float* o2, void* o3, void* stream) {
    CMC_REQUIRE(X && Y && o0, "cmc_msc_windows: null pointer");
    CMC_REQUIRE(W >= 0 && F >= 1 && Ne >= 1 && Nm >= 1 && ldx >= Ne && ldy >= Nm,
                "cmc_msc_windows: bad shape W=%d F=%d Ne=%d Nm=%d", W, F, Ne, Nm);
    CMC_REQUIRE(!jk || (o1 && o2), "cmc_msc_windows: jackknife needs ci_lo and ci_hi");
    CMC_REQUIRE(!jk || K >= 2, "cmc_msc_windows: jackknife needs K >= 2 tapers");
    CMC_REQUIRE(K >= 1 && K <= 15 && msc_smem_bytes(K, Ne, Nm, jk != 0) <= 200 * 1024,
                "cmc_msc_windows: K * (Ne + Nm) too large for the staging buffer");
    CMC_REQUIRE(W <= 65535, "cmc_msc_windows: more than 65535 windows per call");
    if (W == 0) return CMC_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float2* Xc = reinterpret_cast<const float2*>(X);
    const float2* Yc = reinterpret_cast<const float2*>(Y);
#define CMC_MSC_CASE(KK)                                                                          \
    case KK:                                                                                      \
        return launch_msc<KK>(maxemg, Xc, Yc, W, F, Ne, Nm, ldx, ldy, mask, jk, t_crit, it,        \
                              zero_nonsig, o0, o1, o2, o3, st)
    switch (K) {
        CMC_MSC_CASE(1); CMC_MSC_CASE(2); CMC_MSC_CASE(3); CMC_MSC_CASE(4); CMC_MSC_CASE(5);
        CMC_MSC_CASE(6); CMC_MSC_CASE(7); CMC_MSC_CASE(8); CMC_MSC_CASE(9); CMC_MSC_CASE(10);
        CMC_MSC_CASE(11); CMC_MSC_CASE(12); CMC_MSC_CASE(13); CMC_MSC_CASE(14); CMC_MSC_CASE(15);
    }
#undef CMC_MSC_CASE
    set_error("cmc_msc_windows: K=%d tapers unsupported (1..15)", K);
    return CMC_EUNSUPPORTED;
}

}  // namespace cmc

extern "C" int cmc_msc_windows(const float* X, const float* Y, int W, int K, int F, int Ne, int Nm,
                               int64_t ldx, int64_t ldy, const uint8_t* window_mask,
                               int jackknife, float t_crit, float it_threshold,
                               float* coh, float* ci_lo, float* ci_hi, uint8_t* significant,
                               void* stream) {
    return cmc::dispatch_msc(false, X, Y, W, K, F, Ne, Nm, ldx, ldy, window_mask, jackknife, t_crit,
                             it_threshold, 0, coh, ci_lo, ci_hi, !(it_threshold < 0.f) ? significant : nullptr,   // NaN: all-false mask
                             stream);
}

extern "C" int cmc_msc_windows_maxemg(const float* X, const float* Y, int W, int K, int F, int Ne, int Nm,
                                      int64_t ldx, int64_t ldy, const uint8_t* window_mask,
                                      int jackknife, float t_crit, float it_threshold,
                                      int zero_nonsignificant,
                                      float* out_coh, float* out_lo, float* out_hi, int32_t* out_arg,
                                      void* stream) {
    return cmc::dispatch_msc(true, X, Y, W, K, F, Ne, Nm, ldx, ldy, window_mask, jackknife, t_crit,
                             it_threshold, zero_nonsignificant, out_coh, out_lo, out_hi, out_arg, stream);
}

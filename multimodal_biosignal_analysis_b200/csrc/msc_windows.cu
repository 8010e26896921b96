// K2w: per-window multitaper magnitude-squared coherence for all (EEG, EMG) pairs, with the
// leave-one-taper-out jackknife CI and the independence-threshold mask fused in.
//
// Replaces signal_features.py:750-796 (PSD / CSD sums, raw coherence, threshold mask) and
// jackknife_coherence_and_ci (:484-578).  The average here runs over only K (~5) tapers, so
// this is an outer-product + transcendental epilogue bound by the HBM write of the
// (W, F, Ne, Nm) outputs, not a GEMM: CUDA cores, one CTA per (window, frequency chunk),
// spectra of one (window, frequency) staged in shared memory, outputs written fully
// coalesced (the pair index i * Nm + j is the contiguous output index).
#include "common.cuh"

namespace cmc {

constexpr int kMscThreads = 256;
constexpr int kFreqPerBlock = 4;

// The transcendental epilogue dominates this kernel (6 Fisher transforms + 2 inverse transforms per output) and the
// kernel is issue bound, so both are written on the bare MUFU approximations with flush-to-zero (no denormal can
// occur: the arguments are clamped into [6e-8, 3.4e7]) - the default __logf / __expf / sqrtf expansions carry a
// denormal rescue and a Newton step per call, which doubled the instruction count of an output (431 -> ~220):
//   z(c)   = 0.5 ln((1 + c) / (1 - c)) = (ln 2 / 2) * lg2((1 + c) * rcp(1 - c))      RCP + LG2, 2 ulp each
//   z^-1   = tanh(z)^2,  tanh(z) = 1 - 2 * rcp(ex2(2 z log2 e) + 1)                   EX2 + RCP
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
__device__ __forceinline__ float mufu_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fisher_z(float c) {
    // signal_features.py:459-462 with the clip bounds representable in float32
    c = fminf(fmaxf(c, 1e-10f), 0.99999994f);
    return 0.34657359028f * mufu_lg2((1.0f + c) * mufu_rcp(1.0f - c));
}
__device__ __forceinline__ float inv_fisher(float z) {
    // tanh(z)^2 with tanh(z) = 1 - 2 / (exp(2 z) + 1); |z| is capped where tanh saturates in float32
    const float e = mufu_ex2(2.88539008178f * fminf(fmaxf(z, -12.0f), 12.0f));
    const float t = fmaf(-2.0f, mufu_rcp(e + 1.0f), 1.0f);
    return t * t;
}
// The jackknife works in units of z' = 2 z / ln 2 = lg2((1 + c) / (1 - c)): exp(2 z) = 2^z', so neither direction
// needs a multiplication by a constant, and the variance / CI half width scale along (t_crit * se is linear in z).
// The lower clip of the reference (c >= 1e-10) moves z by < 2e-10 and is dropped; the upper clip is folded into the
// clamp of the replicate coherence (1 - 2^-24 instead of 1: 6e-8 on a coherence).
__device__ __forceinline__ float fisher_z2(float c) {          // c in [0, 1 - 2^-24]
    return mufu_lg2((1.0f + c) * mufu_rcp(1.0f - c));
}
__device__ __forceinline__ float inv_fisher_z2(float z2) {     // tanh(z)^2 for z = z2 ln 2 / 2
    const float e = mufu_ex2(fminf(fmaxf(z2, -34.6f), 34.6f));
    const float t = fmaf(-2.0f, mufu_rcp(e + 1.0f), 1.0f);
    return t * t;
}
__device__ __forceinline__ float msc_ratio(float re, float im, float sxx, float syy) {
    // clip(|sxy|^2 / max(sxx * syy, tiny), 0, 1) evaluated as |sxy / sqrt(sxx) / sqrt(syy)|^2 so that
    // the product of the auto-spectra cannot overflow / underflow in float32
    if (!(sxx > 0.f) || !(syy > 0.f)) return 0.f;
    const float r = rsqrtf(sxx) * rsqrtf(syy);
    const float a = re * r, b = im * r;
    return fminf(a * a + b * b, 1.0f);
}

struct PairStats {
    float coh, lo, hi;
};

template <int K, bool JK>
__device__ __forceinline__ PairStats pair_stats(const float2 (&x)[K], const float2 (&y)[K], float t_crit) {
    float2 c[K];
    float px[K], py[K];
    float sxx = 0.f, syy = 0.f, sre = 0.f, sim = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        // conj(x) * y
        c[k] = make_float2(x[k].x * y[k].x + x[k].y * y[k].y, x[k].x * y[k].y - x[k].y * y[k].x);
        px[k] = x[k].x * x[k].x + x[k].y * x[k].y;
        py[k] = y[k].x * y[k].x + y[k].y * y[k].y;
        sxx += px[k];
        syy += py[k];
        sre += c[k].x;
        sim += c[k].y;
    }
    PairStats out;
    if (!JK) {
        out.coh = msc_ratio(sre, sim, sxx, syy);
        out.lo = out.hi = 0.f;
        return out;
    }
    // leave-one-out sums as prefix + suffix (no cancellation, unlike total - term)
    float pre_re[K], pre_im[K], pre_x[K], pre_y[K];
    float a = 0.f, b = 0.f, cx = 0.f, cy = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        pre_re[k] = a; pre_im[k] = b; pre_x[k] = cx; pre_y[k] = cy;
        a += c[k].x; b += c[k].y; cx += px[k]; cy += py[k];
    }
    float z[K];
    float csum = 0.f, zsum = 0.f;
    a = b = cx = cy = 0.f;
#pragma unroll
    for (int k = K - 1; k >= 0; --k) {
        const float ck = msc_ratio(pre_re[k] + a, pre_im[k] + b, pre_x[k] + cx, pre_y[k] + cy);
        z[k] = fisher_z(ck);
        csum += ck;
        zsum += z[k];
        a += c[k].x; b += c[k].y; cx += px[k]; cy += py[k];
    }
    const float mean = fminf(fmaxf(csum * (1.0f / K), 0.f), 1.f);
    const float zbar = zsum * (1.0f / K);
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) ss += (z[k] - zbar) * (z[k] - zbar);
    const float se = mufu_sqrt(ss * ((float)(K - 1) / (float)K));
    const float zc = fisher_z(mean);
    out.coh = mean;
    out.lo = fminf(inv_fisher(zc - t_crit * se), mean);
    out.hi = fmaxf(inv_fisher(zc + t_crit * se), mean);
    return out;
}

// Jackknife statistics of one pair from the spectra and the per-channel leave-one-taper-out factors
// rx[k] = 1 / sqrt(sum_{m != k} |x_m|^2) (0 for a silent channel), same for ry: the auto-spectra part of every
// replicate depends on one channel only, so it is computed once per (window, bin, channel) instead of once per pair.
// Arithmetic and summation order are those of pair_stats<K, true>, the results are bit-identical.
template <int K>
__device__ __forceinline__ PairStats pair_stats_jk(const float2 (&x)[K], const float2 (&y)[K], const float (&rx)[K],
                                                   const float (&ry)[K], float t_crit) {
    float2 c[K];
#pragma unroll
    for (int k = 0; k < K; ++k)
        c[k] = make_float2(x[k].x * y[k].x + x[k].y * y[k].y, x[k].x * y[k].y - x[k].y * y[k].x);   // conj(x) * y
    float pre_re[K], pre_im[K];
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        pre_re[k] = a; pre_im[k] = b;
        a += c[k].x; b += c[k].y;
    }
    float z[K];
    float csum = 0.f, zsum = 0.f;
    a = b = 0.f;
#pragma unroll
    for (int k = K - 1; k >= 0; --k) {
        const float r = rx[k] * ry[k];
        const float u = (pre_re[k] + a) * r, v = (pre_im[k] + b) * r;
        const float ck = fminf(fmaf(u, u, v * v), 0.99999994f);
        z[k] = fisher_z2(ck);
        csum += ck;
        zsum += z[k];
        a += c[k].x; b += c[k].y;
    }
    const float mean = csum * (1.0f / K);                      // every ck is in [0, 1 - 2^-24]
    const float zbar = zsum * (1.0f / K);
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) ss = fmaf(z[k] - zbar, z[k] - zbar, ss);
    const float hw = t_crit * mufu_sqrt(ss * ((float)(K - 1) / (float)K));
    const float zc = fisher_z2(mean);
    PairStats out;
    out.coh = mean;
    out.lo = fminf(inv_fisher_z2(zc - hw), mean);
    out.hi = fmaxf(inv_fisher_z2(zc + hw), mean);
    return out;
}

// leave-one-out inverse roots of the staged spectra: rs[k * n + c] for the n = Ne + Nm channels (sx then sy)
template <int K>
__device__ __forceinline__ void stage_loo_roots(const float2* sx, const float2* sy, float* rs, int Ne, int Nm) {
    for (int c = threadIdx.x; c < Ne + Nm; c += blockDim.x) {
        const float2* s = c < Ne ? sx + c : sy + (c - Ne);
        const int n = c < Ne ? Ne : Nm;
        float pw[K], pre[K];
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float2 v = s[k * n];
            pw[k] = v.x * v.x + v.y * v.y;
            pre[k] = acc;
            acc += pw[k];
        }
        acc = 0.f;
#pragma unroll
        for (int k = K - 1; k >= 0; --k) {
            const float loo = pre[k] + acc;
            rs[k * (Ne + Nm) + c] = loo > 0.f ? rsqrtf(loo) : 0.f;
            acc += pw[k];
        }
    }
}

template <int K>
__device__ __forceinline__ void stage_spectra(float2* sx, float2* sy, const float2* __restrict__ X,
                                              const float2* __restrict__ Y, int w, int f, int F, int Ne,
                                              int Nm, int64_t ldx, int64_t ldy) {
    for (int q = threadIdx.x; q < K * (Ne + Nm); q += blockDim.x) {
        const int k = q / (Ne + Nm), c = q % (Ne + Nm);
        if (c < Ne)
            sx[k * Ne + c] = __ldg(X + ((int64_t)(w * K + k) * F + f) * ldx + c);
        else
            sy[k * Nm + (c - Ne)] = __ldg(Y + ((int64_t)(w * K + k) * F + f) * ldy + (c - Ne));
    }
}

template <int K, bool JK>
__global__ void __launch_bounds__(kMscThreads)
msc_windows_kernel(const float2* __restrict__ X, const float2* __restrict__ Y, int F, int Ne, int Nm,
                   int64_t ldx, int64_t ldy, const uint8_t* __restrict__ window_mask, float t_crit,
                   float it_threshold, float* __restrict__ coh, float* __restrict__ ci_lo,
                   float* __restrict__ ci_hi, uint8_t* __restrict__ significant) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sx = reinterpret_cast<float2*>(smem_raw);
    float2* sy = sx + K * Ne;
    float* rs = reinterpret_cast<float*>(sy + K * Nm);          // [K][Ne + Nm] leave-one-out inverse roots (JK)
    const int w = blockIdx.y;
    if (window_mask && !window_mask[w]) return;
    const int n_pairs = Ne * Nm;
    const int f_end = min(F, (int)(blockIdx.x + 1) * kFreqPerBlock);
    for (int f = blockIdx.x * kFreqPerBlock; f < f_end; ++f) {
        __syncthreads();
        stage_spectra<K>(sx, sy, X, Y, w, f, F, Ne, Nm, ldx, ldy);
        __syncthreads();
        if (JK) {
            stage_loo_roots<K>(sx, sy, rs, Ne, Nm);
            __syncthreads();
        }
        const int64_t obase = ((int64_t)w * F + f) * n_pairs;
        auto emit = [&](int p, const float2 (&x)[K], const float2 (&y)[K]) {
            const PairStats s = pair_stats<K, JK>(x, y, t_crit);
            coh[obase + p] = s.coh;
            if (JK) {
                ci_lo[obase + p] = s.lo;
                ci_hi[obase + p] = s.hi;
            }
            if (significant) significant[obase + p] = s.coh > it_threshold ? 1 : 0;
        };
        if (!JK && kMscThreads % Nm == 0) {
            // (no-jackknife variant only: with the CI epilogue the extra live registers cost more than the loads)
            // the EMG column of a thread is fixed (p advances by a multiple of Nm): keep its K spectra in registers
            const int j = threadIdx.x % Nm;
            float2 y[K];
#pragma unroll
            for (int k = 0; k < K; ++k) y[k] = sy[k * Nm + j];
            for (int i = threadIdx.x / Nm; i < Ne; i += kMscThreads / Nm) {
                float2 x[K];
#pragma unroll
                for (int k = 0; k < K; ++k) x[k] = sx[k * Ne + i];
                emit(i * Nm + j, x, y);
            }
        } else if (JK && kMscThreads % Nm == 0) {
            // the EMG column of a thread is fixed (p advances by a multiple of Nm): its spectra and leave-one-out
            // factors stay in registers, the EEG side is a broadcast load per row
            const int j = threadIdx.x % Nm;
            float2 y[K];
            float ry[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                y[k] = sy[k * Nm + j];
                ry[k] = rs[k * (Ne + Nm) + Ne + j];
            }
            for (int i = threadIdx.x / Nm; i < Ne; i += kMscThreads / Nm) {
                float2 x[K];
                float rx[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    x[k] = sx[k * Ne + i];
                    rx[k] = rs[k * (Ne + Nm) + i];
                }
                const PairStats st = pair_stats_jk<K>(x, y, rx, ry, t_crit);
                const int64_t o = obase + i * Nm + j;
                coh[o] = st.coh;
                ci_lo[o] = st.lo;
                ci_hi[o] = st.hi;
                if (significant) significant[o] = st.coh > it_threshold ? 1 : 0;
            }
        } else if (JK) {
            for (int p = threadIdx.x; p < n_pairs; p += kMscThreads) {
                const int i = p / Nm, j = p - i * Nm;
                float2 x[K], y[K];
                float rx[K], ry[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    x[k] = sx[k * Ne + i];
                    y[k] = sy[k * Nm + j];
                    rx[k] = rs[k * (Ne + Nm) + i];
                    ry[k] = rs[k * (Ne + Nm) + Ne + j];
                }
                const PairStats st = pair_stats_jk<K>(x, y, rx, ry, t_crit);
                coh[obase + p] = st.coh;
                ci_lo[obase + p] = st.lo;
                ci_hi[obase + p] = st.hi;
                if (significant) significant[obase + p] = st.coh > it_threshold ? 1 : 0;
            }
        } else {
            for (int p = threadIdx.x; p < n_pairs; p += kMscThreads) {
                const int i = p / Nm, j = p - i * Nm;
                float2 x[K], y[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    x[k] = sx[k * Ne + i];
                    y[k] = sy[k * Nm + j];
                }
                emit(p, x, y);
            }
        }
    }
}

// fused EMG-argmax variant: one warp per (window, frequency, EEG channel)
template <int K, bool JK>
__global__ void __launch_bounds__(kMscThreads)
msc_windows_maxemg_kernel(const float2* __restrict__ X, const float2* __restrict__ Y, int F, int Ne, int Nm,
                          int64_t ldx, int64_t ldy, const uint8_t* __restrict__ window_mask, float t_crit,
                          float it_threshold, int zero_nonsig, float* __restrict__ out_coh,
                          float* __restrict__ out_lo, float* __restrict__ out_hi,
                          int32_t* __restrict__ out_arg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sx = reinterpret_cast<float2*>(smem_raw);
    float2* sy = sx + K * Ne;
    float* rs = reinterpret_cast<float*>(sy + K * Nm);          // [K][Ne + Nm] leave-one-out inverse roots (JK)
    const int w = blockIdx.y;
    if (window_mask && !window_mask[w]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f_end = min(F, (int)(blockIdx.x + 1) * kFreqPerBlock);
    for (int f = blockIdx.x * kFreqPerBlock; f < f_end; ++f) {
        __syncthreads();
        stage_spectra<K>(sx, sy, X, Y, w, f, F, Ne, Nm, ldx, ldy);
        __syncthreads();
        if (JK) {
            stage_loo_roots<K>(sx, sy, rs, Ne, Nm);
            __syncthreads();
        }
        for (int i = warp; i < Ne; i += kMscThreads / 32) {
            float2 x[K];
            float rx[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                x[k] = sx[k * Ne + i];
                rx[k] = JK ? rs[k * (Ne + Nm) + i] : 0.f;
            }
            float best = -1.f, blo = 0.f, bhi = 0.f;
            int bj = 0x7fffffff;
            for (int j = lane; j < Nm; j += 32) {
                float2 y[K];
                float ry[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    y[k] = sy[k * Nm + j];
                    ry[k] = JK ? rs[k * (Ne + Nm) + Ne + j] : 0.f;
                }
                PairStats s;
                if (JK) s = pair_stats_jk<K>(x, y, rx, ry, t_crit);
                else s = pair_stats<K, false>(x, y, t_crit);
                float val = s.coh;
                if (zero_nonsig && !(s.coh > it_threshold)) val = 0.f;
                if (val > best) { best = val; blo = s.lo; bhi = s.hi; bj = j; }
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
    const size_t smem = (sizeof(float2) + sizeof(float)) * K * (Ne + Nm);      // spectra + leave-one-out roots
    if (!maxemg) {
        if (jk)
            msc_windows_kernel<K, true><<<grid, kMscThreads, smem, st>>>(
                X, Y, F, Ne, Nm, ldx, ldy, mask, t_crit, it, o0, o1, o2, (uint8_t*)o3);
        else
            msc_windows_kernel<K, false><<<grid, kMscThreads, smem, st>>>(
                X, Y, F, Ne, Nm, ldx, ldy, mask, t_crit, it, o0, o1, o2, (uint8_t*)o3);
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
    CMC_REQUIRE((size_t)K * (Ne + Nm) * (sizeof(float2) + sizeof(float)) <= 48 * 1024,
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

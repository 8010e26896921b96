// K1: fused detrend + taper + batched real FFT over Welch / multitaper segments.
//
// Replaces np.fft.rfft(window * taper) of the reference (signal_features.py:743-748,
// :412-420; scipy Welch internals for preprocessing.py:1228).
//
// Layout ("batch fastest"): one CTA owns one segment and a tile of CT adjacent channels.
// The time-first input x[t][c] makes the CT channels of one sample contiguous, so global
// loads are sector-exact, the shared-memory work array buf[point][channel] is bank-conflict
// free at every butterfly stride, and the twiddle of a butterfly is shared by the CT lanes
// that process the same point for different channels.
//
// Real FFT of length N through a complex FFT of length M = N/2 on z[m] = x[2m] + i x[2m+1]
// (Stockham autosort, radix 16/8 in registers, in-place in shared memory with register
// staging), then the split X[b] = E[b] - i W_N^b O[b] for the requested bins only.
#include "common.cuh"
#include "fft_common.cuh"
#include <stdlib.h>
#include <math.h>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace cmc {

constexpr int kElemsPerThread = 32;

// Stockham pass of radix R on buf[M][CT] where NS = product of the previous radices.
template <int M, int CT, int NT, int R, int NS>
__device__ __forceinline__ void pass_smem(float2* buf, const float2* __restrict__ twM, int tid) {
    constexpr int B = kElemsPerThread / R;
    constexpr int STRIDE = M / R;
    float2 v[B][R];
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int q = tid + b * NT;
        const int c = q % CT, j = q / CT;
#pragma unroll
        for (int r = 0; r < R; ++r) v[b][r] = buf[(j + r * STRIDE) * CT + c];
        const int k = j % NS;
        apply_twiddles<R>(v[b], twM, k * (M / (NS * R)));
        dft<R>(v[b]);
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int q = tid + b * NT;
        const int c = q % CT, j = q / CT;
        const int k = j % NS;
        const int j0 = (j / NS) * (NS * R) + k;
#pragma unroll
        for (int r = 0; r < R; ++r) buf[(j0 + r * NS) * CT + c] = v[b][r];
    }
    __syncthreads();
}

template <int M, int CT>
__global__ void __launch_bounds__(M * CT / kElemsPerThread, (M * CT / kElemsPerThread) <= 256 ? 2 : 1)
fft_segments_kernel(const float* __restrict__ x, int64_t n_samples, int n_ch, int64_t ld,
                    const int64_t* __restrict__ seg_starts,
                    const float* __restrict__ windows, int n_win, int detrend,
                    int bin_lo, int F,
                    float2* __restrict__ spec, int64_t spec_ld,
                    const float2* __restrict__ twM, const float2* __restrict__ twN) {
    constexpr int N = 2 * M;
    constexpr int NT = M * CT / kElemsPerThread;
    constexpr int R0 = Plan<M>::R0, R1 = Plan<M>::R1, R2 = Plan<M>::R2;
    constexpr int B0 = kElemsPerThread / R0;
    constexpr int S0 = M / R0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* buf = reinterpret_cast<float2*>(smem_raw);
    float* part = reinterpret_cast<float*>(smem_raw + sizeof(float2) * M * CT);  // [NT/32][32]
    float* mean_s = part + NT;                                                    // [CT]

    const int tid = threadIdx.x;
    const int seg = blockIdx.x;
    const int c0 = blockIdx.y * CT;
    const int64_t start = seg_starts[seg];
    const float* xs = x + start * ld + c0;

    for (int kw = 0; kw < n_win; ++kw) {
        const float* win = windows + (int64_t)kw * N;
        float2 v[B0][R0];
        // ---- load raw samples for this thread's first-pass butterflies ----
#pragma unroll
        for (int b = 0; b < B0; ++b) {
            const int q = tid + b * NT;
            const int c = q % CT, j = q / CT;
            const bool ok = (c0 + c) < n_ch;
#pragma unroll
            for (int r = 0; r < R0; ++r) {
                const int n0 = 2 * (j + r * S0);
                float a = 0.f, bb = 0.f;
                if (ok) {
                    a = __ldg(xs + (int64_t)n0 * ld + c);
                    bb = __ldg(xs + (int64_t)(n0 + 1) * ld + c);
                }
                v[b][r] = make_float2(a, bb);
            }
        }
        float mu = 0.f;
        if (detrend == CMC_DETREND_CONSTANT) {
            // deterministic per-channel mean: lane partials -> per-warp rows -> fixed-order sum
            if (kw == 0) {
                float s = 0.f;
#pragma unroll
                for (int b = 0; b < B0; ++b)
#pragma unroll
                    for (int r = 0; r < R0; ++r) s += v[b][r].x + v[b][r].y;
                // lanes with equal (lane % CT) hold the same channel (NT % CT == 0, CT | 32)
#pragma unroll
                for (int off = 16; off >= CT; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                if ((tid & 31) < CT) part[(tid >> 5) * CT + (tid & 31)] = s;
                __syncthreads();
                if (tid < CT) {
                    float t = 0.f;
                    for (int w = 0; w < NT / 32; ++w) t += part[w * CT + tid];
                    mean_s[tid] = t * (1.0f / N);
                }
                __syncthreads();
            }
            mu = mean_s[tid % CT];
        }
        // ---- taper and first Stockham pass (NS = 1: no twiddles) ----
#pragma unroll
        for (int b = 0; b < B0; ++b) {
            const int q = tid + b * NT;
            const int c = q % CT, j = q / CT;
#pragma unroll
            for (int r = 0; r < R0; ++r) {
                const int n0 = 2 * (j + r * S0);
                const float2 w = __ldg(reinterpret_cast<const float2*>(win + n0));
                v[b][r].x = (v[b][r].x - mu) * w.x;
                v[b][r].y = (v[b][r].y - mu) * w.y;
            }
            dft<R0>(v[b]);
#pragma unroll
            for (int r = 0; r < R0; ++r) buf[(j * R0 + r) * CT + c] = v[b][r];
        }
        __syncthreads();
        if (R1 > 1) pass_smem<M, CT, NT, (R1 > 1 ? R1 : 2), R0>(buf, twM, tid);
        if (R2 > 1) pass_smem<M, CT, NT, (R2 > 1 ? R2 : 2), R0 * R1>(buf, twM, tid);

        // ---- real-FFT split for the requested bins, coalesced over channels ----
        float2* out = spec + ((int64_t)(seg * n_win + kw) * F) * spec_ld + c0;
        for (int q = tid; q < F * CT; q += NT) {
            const int c = q % CT, bi = q / CT;
            const int b = bin_lo + bi;
            if (c0 + c < n_ch) {
                const float2 A = buf[(b & (M - 1)) * CT + c];
                const float2 Bz = buf[((M - b) & (M - 1)) * CT + c];
                const float2 E = make_float2(0.5f * (A.x + Bz.x), 0.5f * (A.y - Bz.y));
                const float2 O = make_float2(0.5f * (A.x - Bz.x), 0.5f * (A.y + Bz.y));
                const float2 T = cmul(__ldg(twN + b), O);
                float2 X = make_float2(E.x + T.y, E.y - T.x);
                if (b == 0 || b == M) X.y = 0.f;
                if (detrend == CMC_DETREND_POST_TAPER && b == 0) X.x = 0.f;
                out[(int64_t)bi * spec_ld + c] = X;
            }
        }
        __syncthreads();
    }
}

template <int M, int CT>
static int launch_fft(const float* x, int64_t n_samples, int n_ch, int64_t ld,
                      const int64_t* seg_starts, int n_seg, const float* windows, int n_win,
                      int detrend, int bin_lo, int F, float2* spec, int64_t spec_ld,
                      const float2* twM, const float2* twN, cudaStream_t st) {
    constexpr int NT = M * CT / kElemsPerThread;
    static_assert(NT >= 32 && NT <= 1024 && NT % 32 == 0, "bad CTA size");
    static_assert(32 % CT == 0, "CT must divide the warp");
    const size_t smem = sizeof(float2) * M * CT + sizeof(float) * (NT + CT);
    auto kern = fft_segments_kernel<M, CT>;
    int rc = ensure_smem_attr(reinterpret_cast<const void*>(kern), smem);
    if (rc) return rc;
    dim3 grid(n_seg, (n_ch + CT - 1) / CT);
    kern<<<grid, NT, smem, st>>>(x, n_samples, n_ch, ld, seg_starts, windows, n_win, detrend,
                                 bin_lo, F, spec, spec_ld, twM, twN);
    CMC_CHECK_LAUNCH("fft_segments_kernel");
    return CMC_OK;
}

// Arbitrary-length fallback (N not a supported power of two): direct O(N F) DFT per (segment, channel) with the
// same detrend / taper semantics.  Used by validation-style calls (e.g. Welch PSDs with nperseg = 4 fs on
// 500 Hz recordings); coalesced over channels, twiddles exp(-2 pi i q / N) from a per-device table.
__global__ void __launch_bounds__(256)
dft_direct_kernel(const float* __restrict__ x, int n_ch, int64_t ld, const int64_t* __restrict__ seg_starts,
                  const float* __restrict__ windows, int n_win, int N, int detrend, int bin_lo, int F,
                  float2* __restrict__ spec, int64_t spec_ld, const float2* __restrict__ twN) {
    const int c = blockIdx.y * 32 + threadIdx.x;
    const int seg = blockIdx.x / n_win, kw = blockIdx.x % n_win;
    if (c >= n_ch) return;
    const float* xs = x + seg_starts[seg] * ld + c;
    const float* win = windows + (int64_t)kw * N;
    float mu = 0.f;
    if (detrend == CMC_DETREND_CONSTANT) {
        float s = 0.f;
        for (int n = 0; n < N; ++n) s += xs[(int64_t)n * ld];
        mu = s / N;
    }
    float2* out = spec + ((int64_t)blockIdx.x * F) * spec_ld + c;
    for (int bi = threadIdx.y; bi < F; bi += blockDim.y) {
        const int b = bin_lo + bi;
        float re = 0.f, im = 0.f;
        int q = 0;                                   // (b * n) mod N
        for (int n = 0; n < N; ++n) {
            const float v = (xs[(int64_t)n * ld] - mu) * __ldg(win + n);
            const float2 w = __ldg(twN + q);
            re += v * w.x;
            im += v * w.y;
            q += b;
            if (q >= N) q -= N;
        }
        if (detrend == CMC_DETREND_POST_TAPER && b == 0) re = 0.f;
        if (b == 0 || 2 * b == N) im = 0.f;
        out[(int64_t)bi * spec_ld] = make_float2(re, im);
    }
}

// full-circle table exp(-2 pi i q / N), q in [0, N), for the direct kernel
static int get_full_twiddles(int N, const float2** tw, cudaStream_t st = nullptr) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, float2*> cache;
    int dev = 0;
    int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(dev, N);
    auto it = cache.find(key);
    if (it == cache.end()) {
        if ((rc = refuse_table_during_capture(st, "cmc_fft_segments (direct DFT)", N))) return rc;
        std::vector<float2> h(N);
        for (int q = 0; q < N; ++q) {
            const double a = -6.283185307179586476925286766559 * q / N;
            h[q] = make_float2((float)cos(a), (float)sin(a));
        }
        float2* d = nullptr;
        rc = check_cuda(cudaMalloc(&d, h.size() * sizeof(float2)), "cudaMalloc(dft table)");
        if (rc) return rc;
        rc = check_cuda(cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice), "cudaMemcpy(dft table)");
        if (rc) return rc;
        it = cache.emplace(key, d).first;
    }
    *tw = it->second;
    return CMC_OK;
}

int fft_segments_tma(const float* x, int64_t n_samples, int n_ch, int64_t ld, const int64_t* seg_starts, int n_seg,
                     const float* windows, int n_win, int N, int detrend, int bin_lo, int F, float2* spec,
                     int64_t spec_ld, const float2* twM, const float2* twN, cudaStream_t st, const float* x2 = nullptr,
                     int n_ch2 = 0, int64_t ld2 = 0, float2* spec2 = nullptr);

}  // namespace cmc

extern "C" int cmc_fft_segments(const float* x, int64_t n_samples, int n_ch, int64_t ld,
                                const int64_t* seg_starts, int n_seg,
                                const float* windows, int n_win, int N, int detrend,
                                int bin_lo, int bin_hi,
                                float* spec, int64_t spec_ld, void* stream) {
    using namespace cmc;
    CMC_REQUIRE(x && seg_starts && windows && spec, "cmc_fft_segments: null pointer");
    CMC_REQUIRE(n_ch >= 1 && ld >= n_ch && spec_ld >= n_ch, "cmc_fft_segments: bad channel pitch");
    CMC_REQUIRE(n_seg >= 0 && n_win >= 1, "cmc_fft_segments: bad segment/window count");
    CMC_REQUIRE(detrend >= 0 && detrend <= 2, "cmc_fft_segments: detrend must be 0, 1 or 2");
    CMC_REQUIRE(bin_lo >= 0 && bin_hi >= bin_lo && bin_hi <= N / 2,
                "cmc_fft_segments: bins [%d, %d] outside [0, %d]", bin_lo, bin_hi, N / 2);
    CMC_REQUIRE((reinterpret_cast<uintptr_t>(windows) & 7) == 0 && (reinterpret_cast<uintptr_t>(spec) & 7) == 0,
                "cmc_fft_segments: windows/spec must be 8-byte aligned");
    if (n_seg == 0) return CMC_OK;
    if (N < 128 || N > 8192 || (N & (N - 1))) {
        CMC_REQUIRE(N >= 2 && N <= (1 << 20), "cmc_fft_segments: N=%d outside [2, 2^20]", N);
        CMC_REQUIRE((int64_t)n_seg * n_win <= 2147483647ll, "cmc_fft_segments: too many segments");
        const float2* tw;
        int rc0 = get_full_twiddles(N, &tw, static_cast<cudaStream_t>(stream));
        if (rc0) return rc0;
        dft_direct_kernel<<<dim3((unsigned)(n_seg * n_win), (n_ch + 31) / 32), dim3(32, 8), 0,
                            static_cast<cudaStream_t>(stream)>>>(x, n_ch, ld, seg_starts, windows, n_win, N, detrend, bin_lo,
                                                                 bin_hi - bin_lo + 1, reinterpret_cast<float2*>(spec),
                                                                 spec_ld, tw);
        CMC_CHECK_LAUNCH("dft_direct_kernel");
        return CMC_OK;
    }
    const float2 *twM, *twN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = get_twiddles(N, &twM, &twN, st);
    if (rc) return rc;
    const int F = bin_hi - bin_lo + 1;
    float2* sp = reinterpret_cast<float2*>(spec);
    // fast path: TMA-staged two-channel kernel (fft_tma.cu); returns 1 when the layout does not qualify
    static const bool no_tma = getenv("CMC_FFT_NO_TMA") != nullptr;
    if (!no_tma) {
        rc = fft_segments_tma(x, n_samples, n_ch, ld, seg_starts, n_seg, windows, n_win, N, detrend, bin_lo, F, sp,
                              spec_ld, twM, twN, st);
        if (rc != 1) return rc;
    }
#define CMC_FFT_CASE(NN, CT)                                                                   \
    case NN:                                                                                   \
        return launch_fft<NN / 2, CT>(x, n_samples, n_ch, ld, seg_starts, n_seg, windows, n_win, \
                                      detrend, bin_lo, F, sp, spec_ld, twM, twN, st)
    switch (N) {
        CMC_FFT_CASE(128, 32);
        CMC_FFT_CASE(256, 32);
        CMC_FFT_CASE(512, 32);
        CMC_FFT_CASE(1024, 16);
        CMC_FFT_CASE(2048, 8);
        CMC_FFT_CASE(4096, 8);
        CMC_FFT_CASE(8192, 4);
    }
#undef CMC_FFT_CASE
    return CMC_EUNSUPPORTED;
}

// Creates the per-device tables cmc_fft_segments needs for segment length N (twiddles, or the full-circle table of the
// direct DFT for lengths the FFT kernels do not take) so that the first call for N may happen inside a stream capture.
extern "C" int cmc_fft_prepare(int N) {
    using namespace cmc;
    CMC_REQUIRE(N >= 2 && N <= (1 << 20), "cmc_fft_prepare: N=%d outside [2, 2^20]", N);
    if (N < 128 || N > 8192 || (N & (N - 1))) {
        const float2* tw;
        return get_full_twiddles(N, &tw);
    }
    const float2 *twM, *twN;
    return get_twiddles(N, &twM, &twN);
}

// Two recordings of equal length that share the segment table, the window rows and the bin range (the EEG and the
// EMG array of one subject-condition) in ONE launch of the pipelined K1 kernel when both qualify for it; any other
// case runs as two cmc_fft_segments calls with identical results.  spec1 / spec2 share the row pitch spec_ld
// (typically two channel ranges of one [n_seg][n_win][F][spec_ld] array).
extern "C" int cmc_fft_segments_pair(const float* x1, int n_ch1, int64_t ld1, float* spec1,
                                     const float* x2, int n_ch2, int64_t ld2, float* spec2,
                                     int64_t n_samples, const int64_t* seg_starts, int n_seg,
                                     const float* windows, int n_win, int N, int detrend,
                                     int bin_lo, int bin_hi, int64_t spec_ld, void* stream) {
    using namespace cmc;
    CMC_REQUIRE(x1 && x2 && seg_starts && windows && spec1 && spec2, "cmc_fft_segments_pair: null pointer");
    const bool fast = n_seg > 0 && n_ch1 >= 1 && n_ch2 >= 1 && ld1 >= n_ch1 && ld2 >= n_ch2 && spec_ld >= n_ch1 &&
                      spec_ld >= n_ch2 && n_win >= 1 && detrend >= 0 && detrend <= 2 && bin_lo >= 0 &&
                      bin_hi >= bin_lo && bin_hi <= N / 2 && (N == 512 || N == 1024 || N == 2048) &&
                      (reinterpret_cast<uintptr_t>(windows) & 7) == 0 && (reinterpret_cast<uintptr_t>(spec1) & 7) == 0 &&
                      (reinterpret_cast<uintptr_t>(spec2) & 7) == 0;
    static const bool no_pair = getenv("CMC_FFT_NO_PAIR") != nullptr || getenv("CMC_FFT_NO_TMA") != nullptr;
    if (fast && !no_pair) {
        const float2 *twM, *twN;
        int rc = get_twiddles(N, &twM, &twN, static_cast<cudaStream_t>(stream));
        if (rc) return rc;
        rc = fft_segments_tma(x1, n_samples, n_ch1, ld1, seg_starts, n_seg, windows, n_win, N, detrend, bin_lo,
                              bin_hi - bin_lo + 1, reinterpret_cast<float2*>(spec1), spec_ld, twM, twN,
                              static_cast<cudaStream_t>(stream), x2, n_ch2, ld2, reinterpret_cast<float2*>(spec2));
        if (rc != 1) return rc;
    }
    int rc = cmc_fft_segments(x1, n_samples, n_ch1, ld1, seg_starts, n_seg, windows, n_win, N, detrend, bin_lo, bin_hi,
                              spec1, spec_ld, stream);
    if (rc) return rc;
    return cmc_fft_segments(x2, n_samples, n_ch2, ld2, seg_starts, n_seg, windows, n_win, N, detrend, bin_lo, bin_hi,
                            spec2, spec_ld, stream);
}

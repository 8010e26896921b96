// Power spectral density epilogue on K1's spectra: mean over tapers / segments of |X|^2
// with the scipy density / one-sided conventions and the optional log10 of multitaper_psd.
#include "common.cuh"

namespace cmc {

__global__ void __launch_bounds__(256)
psd_kernel(const float2* __restrict__ spec, int K, int F, int n_ch, int64_t ld_in, float base_scale, int one_sided,
           int bin_lo, int N, int log_scale, float* __restrict__ out, int64_t ld_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y, w = blockIdx.z;
    if (c >= n_ch) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
        const float2 v = __ldg(spec + (((int64_t)w * K + k) * F + f) * ld_in + c);
        acc += v.x * v.x + v.y * v.y;
    }
    const int b = bin_lo + f;
    float s = base_scale / K;
    if (one_sided && b > 0 && 2 * b < N) s *= 2.f;
    float p = acc * s;
    if (log_scale) p = log10f(fabsf(p) + 1e-10f);
    out[((int64_t)w * F + f) * ld_out + c] = p;
}

}  // namespace cmc

extern "C" int cmc_psd_from_spectra(const float* spec, int W, int K, int F, int n_ch, int64_t ld_in, float base_scale,
                                    int one_sided, int bin_lo, int N, int log_scale, float* out, int64_t ld_out,
                                    void* stream) {
    using namespace cmc;
    CMC_REQUIRE(spec && out, "cmc_psd_from_spectra: null pointer");
    CMC_REQUIRE(W >= 0 && K >= 1 && F >= 1 && n_ch >= 1 && ld_in >= n_ch && ld_out >= n_ch && F <= 65535 && W <= 65535,
                "cmc_psd_from_spectra: bad shape");
    if (W == 0) return CMC_OK;
    psd_kernel<<<dim3((n_ch + 63) / 64, F, W), 64, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float2*>(spec), K, F, n_ch, ld_in, base_scale, one_sided, bin_lo, N, log_scale, out, ld_out);
    CMC_CHECK_LAUNCH("psd_kernel");
    return CMC_OK;
}

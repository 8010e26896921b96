// Error reporting, launch bookkeeping and per-device immutable tables of libcmc_b200.
#include "common.cuh"

#include <math.h>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <utility>
#include <vector>

namespace cmc {

static thread_local std::string t_last_error;
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_last_error = buf;
}

int ensure_smem_attr(const void* func, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> done;
    int dev = 0;
    int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(dev, func);
    auto it = done.find(key);
    if (it != done.end() && it->second >= bytes) return CMC_OK;
    rc = check_cuda(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes),
                    "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    if (rc) return rc;
    done[key] = bytes;
    return CMC_OK;
}

// A stream that is being captured cannot take the cudaMalloc + synchronous copy of a first-use table: fail with a
// message that says what to do instead of breaking the capture halfway.
int refuse_table_during_capture(cudaStream_t st, const char* what, int N) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
        set_error("%s: the table for N=%d does not exist yet and cannot be created while the stream is being captured "
                  "into a CUDA graph - call cmc_fft_prepare(N) (or the same function once, eagerly) before the capture",
                  what, N);
        return CMC_EINVAL;
    }
    return CMC_OK;
}

int get_twiddles(int N, const float2** twM, const float2** twN, cudaStream_t st) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, std::pair<float2*, float2*>> cache;
    int dev = 0;
    int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(dev, N);
    auto it = cache.find(key);
    if (it == cache.end()) {
        if ((rc = refuse_table_during_capture(st, "cmc_fft_segments", N))) return rc;
        const int M = N / 2;
        std::vector<float2> h(M + M + 1);
        const double two_pi = 6.283185307179586476925286766559;
        for (int q = 0; q < M; ++q) {
            double a = -two_pi * q / M;
            h[q] = make_float2((float)cos(a), (float)sin(a));
        }
        for (int b = 0; b <= M; ++b) {
            double a = -two_pi * b / N;
            h[M + b] = make_float2((float)cos(a), (float)sin(a));
        }
        float2* d = nullptr;
        rc = check_cuda(cudaMalloc(&d, h.size() * sizeof(float2)), "cudaMalloc(twiddles)");
        if (rc) return rc;
        // synchronous copy on the legacy stream: the table is complete before any kernel that
        // uses it is enqueued by this thread
        rc = check_cuda(cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice),
                        "cudaMemcpy(twiddles)");
        if (rc) return rc;
        it = cache.emplace(key, std::make_pair(d, d + M)).first;
    }
    *twM = it->second.first;
    *twN = it->second.second;
    return CMC_OK;
}

}  // namespace cmc

extern "C" int cmc_abi_version(void) { return CMC_ABI_VERSION; }
extern "C" const char* cmc_last_error(void) { return cmc::t_last_error.c_str(); }
extern "C" int64_t cmc_launch_count(void) { return cmc::g_launches.load(); }

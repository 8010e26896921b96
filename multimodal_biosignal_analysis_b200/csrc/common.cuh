// Shared helpers for libcmc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/cmc.h"

namespace cmc {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_cuda(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return CMC_ECUDA;
    }
    return CMC_OK;
}

#define CMC_CHECK_LAUNCH(name)                                              \
    do {                                                                    \
        cmc::g_launches.fetch_add(1, std::memory_order_relaxed);            \
        int _rc = cmc::check_cuda(cudaGetLastError(), name);                \
        if (_rc != CMC_OK) return _rc;                                      \
    } while (0)

#define CMC_REQUIRE(cond, ...)                                              \
    do {                                                                    \
        if (!(cond)) {                                                      \
            cmc::set_error(__VA_ARGS__);                                    \
            return CMC_EINVAL;                                              \
        }                                                                   \
    } while (0)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// Opt a kernel into > 48 KB of dynamic shared memory (once per device and function).
int ensure_smem_attr(const void* func, size_t bytes);

// Per-device immutable twiddle tables (created on first use, never freed/changed).
//   twM[q] = exp(-2 pi i q / M), q in [0, M);  twN[b] = exp(-2 pi i b / N), b in [0, N/2]
int get_twiddles(int N, const float2** twM, const float2** twN, cudaStream_t st = nullptr);
int refuse_table_during_capture(cudaStream_t st, const char* what, int N);

}  // namespace cmc

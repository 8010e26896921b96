// Micro-benchmark: FP32 throughput of scalar FADD/FFMA vs packed FADD2/FFMA2 on sm_100a, with and without
// competing integer instructions.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 f32x2_bench.cu -o f32x2_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float2 a[8], b = make_float2(s, s * 1.0001f), c = make_float2(0.999f, 1.001f);
    int acc = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 2) {          // scalar: two FFMA per element
                a[i].x = fmaf(a[i].x, c.x, b.x);
                a[i].y = fmaf(a[i].y, c.y, b.y);
            } else {                               // packed: one FFMA2 per element
                a[i] = __ffma2_rn(a[i], c, b);
            }
            if (MODE >= 2) {                       // one integer instruction per element competing for issue slots
                acc = acc * 3 + i;
                acc ^= it;
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + acc;
}

template <int MODE>
float run(float* d, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(d, 10, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(d, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
    const int iters = 20000;
    const double fma_per_launch = 148.0 * 4 * 256 * 16.0 * iters;
    const char* names[4] = {"scalar FFMA", "packed FFMA2", "scalar FFMA + 2 INT", "packed FFMA2 + 2 INT"};
    float ms[4] = {run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters)};
    for (int m = 0; m < 4; ++m)
        printf("%-22s %8.3f ms  %7.2f TFMA/s  (%.1f FMA lanes / clk / SM at 1.965 GHz)\n", names[m], ms[m],
               fma_per_launch / ms[m] / 1e9, fma_per_launch / (ms[m] * 1e-3) / 148 / 1.965e9);
    return 0;
}

// K1 fast path: TMA-staged, two channels per thread.
//
// Same arithmetic as fft_segments_kernel (fft.cu) with three changes that the ncu profile of the first
// version asked for (profiles/r01a_k1_fft_baseline.md):
//   * the raw segment tile [N samples][8 channels] is fetched by TMA (cp.async.bulk.tensor.2d, 32-byte rows)
//     straight into the FFT work buffer - no LDG, no register staging, no long-scoreboard stalls;
//   * the raw layout IS the first-pass input layout (sample 2m -> real part, 2m+1 -> imaginary part of point m),
//     so the transform runs in place: raw tile -> pass 0 -> interleaved complex rows -> remaining passes;
//   * every thread owns TWO adjacent channels of its butterflies: window values, twiddle powers and index
//     arithmetic are shared by both, shared-memory accesses are LDS.128 / STS.128, output stores 16 bytes.
// Shared-memory rows are XOR-swizzled so that all butterfly strides are bank-conflict free.
#include "fft_common.cuh"
#include "tc_common.cuh"
#include "csd_layout.cuh"
#include <stdlib.h>

namespace cmc {

using namespace tc;

constexpr int kTmaCT = 8;            // channels per CTA
constexpr int kPtsPerThread = 16;    // complex points per thread (x 2 channels)

// byte offset of raw sample (row n, channel pair cp) inside the TMA-written tile (SWIZZLE_64B: address bits
// [4,6) ^= bits [7,9))
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool SWZ>
__device__ __forceinline__ uint32_t raw_off(int n, int cp) {
    const uint32_t lin = (uint32_t)n * 32u + (uint32_t)cp * 8u;
    return SWZ ? (lin ^ (((lin >> 7) & 3u) << 4)) : lin;
}
// byte offset of complex point p (channel pair cp) in the interleaved work layout: 64-byte rows, rows of odd
// 16-row groups swapped pairwise so that first-pass stores (row stride 16) spread over all banks
__device__ __forceinline__ uint32_t pt_off(int p, int cp) {
    return (uint32_t)(p ^ ((p >> 4) & 1)) * 64u + (uint32_t)cp * 16u;
}

// Row swizzle p -> p ^ ((p >> 4) & 1) commutes with adding multiples of 32 rows, so it is applied once per
// butterfly (all strides used after the first pass are multiples of 32 rows for M >= 512).
template <int M, int R, int NS>
__device__ __forceinline__ void pass_pair(uint32_t sbuf, const float2* __restrict__ twM, int tid) {
    constexpr int NT = M / 4;
    constexpr int B = kPtsPerThread / R;
    constexpr int STRIDE = M / R;
    constexpr bool kHoist = (STRIDE % 32 == 0) && (NS % 32 == 0 || NS >= 32);
    float2 va[B][R], vb[B][R];
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int q = tid + b * NT;
        const int cp = q & 3, j = q >> 2;
        const uint32_t base = sbuf + pt_off(j, cp);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 t = lds128(STRIDE % 32 == 0 ? base + (uint32_t)(r * STRIDE) * 64u
                                                     : sbuf + pt_off(j + r * STRIDE, cp));
            va[b][r] = make_float2(t.x, t.y);
            vb[b][r] = make_float2(t.z, t.w);
        }
        const int k = j % NS;
        apply_twiddles2<R>(va[b], vb[b], twM, k * (M / (NS * R)));
        dft<R>(va[b]);
        dft<R>(vb[b]);
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int q = tid + b * NT;
        const int cp = q & 3, j = q >> 2;
        const int k = j % NS;
        const int j0 = (j / NS) * (NS * R) + k;
        const uint32_t base = sbuf + pt_off(j0, cp);
#pragma unroll
        for (int r = 0; r < R; ++r)
            sts128(NS % 32 == 0 ? base + (uint32_t)(r * NS) * 64u : sbuf + pt_off(j0 + r * NS, cp),
                   make_float4(va[b][r].x, va[b][r].y, vb[b][r].x, vb[b][r].y));
    }
    __syncthreads();
    (void)kHoist;
}

template <int M, bool SWZ>
__global__ void __launch_bounds__(M / 4, (M / 4) <= 256 ? 2 : 1)
fft_segments_tma_kernel(const __grid_constant__ CUtensorMap tmap, int n_ch,
                        const int64_t* __restrict__ seg_starts, const float* __restrict__ windows, int n_win,
                        int detrend, int bin_lo, int F, float2* __restrict__ spec, int64_t spec_ld,
                        const float2* __restrict__ twM, const float2* __restrict__ twN, int dbg) {
    constexpr int N = 2 * M;
    constexpr int NT = M / 4;
    constexpr int R0 = Plan<M>::R0, R1 = Plan<M>::R1, R2 = Plan<M>::R2;
    constexpr int B0 = kPtsPerThread / R0;
    constexpr int S0 = M / R0;
    constexpr int kRowsPerBox = 256;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* buf = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    float* part = reinterpret_cast<float*>(buf + N * 32);            // [NT / 32][8]
    float* mean_s = part + (NT / 32) * kTmaCT;                       // [8]
    uint64_t* bar = reinterpret_cast<uint64_t*>(mean_s + kTmaCT);

    const uint32_t sbuf = smem_u32(buf);
    const int tid = threadIdx.x;
    const int seg = blockIdx.x;
    const int c0 = blockIdx.y * kTmaCT;
    const int start = (int)seg_starts[seg];
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmap);
    }
    __syncthreads();

    for (int kw = 0; kw < n_win; ++kw) {
        // ---- TMA: raw tile -> shared memory (N rows of 32 bytes, boxes of 256 rows) ----
        if (tid == 0 && !(dbg & 1)) {
            fence_proxy_async();                       // earlier generic-proxy accesses before async-proxy writes
            mbar_arrive_expect_tx(bar, N * 32);
#pragma unroll 1
            for (int i = 0; i < N / kRowsPerBox; ++i)
                tma_load_2d(buf + i * kRowsPerBox * 32, &tmap, bar, c0, start + i * kRowsPerBox);
        }
        if (!(dbg & 1)) mbar_wait(bar, kw & 1);
        if (dbg & 2) { __syncthreads(); continue; }

        const float* win = windows + (int64_t)kw * N;
        float2 va[B0][R0], vb[B0][R0];
#pragma unroll
        for (int b = 0; b < B0; ++b) {
            const int q = tid + b * NT;
            const int cp = q & 3, j = q >> 2;
#pragma unroll
            for (int r = 0; r < R0; ++r) {
                const int n0 = 2 * (j + r * S0);
                const float2 re = lds64(sbuf + raw_off<SWZ>(n0, cp));       // x[2p][c], x[2p][c+1]
                const float2 im = lds64(sbuf + raw_off<SWZ>(n0 + 1, cp));   // x[2p+1][c], x[2p+1][c+1]
                va[b][r] = make_float2(re.x, im.x);
                vb[b][r] = make_float2(re.y, im.y);
            }
        }
        float mua = 0.f, mub = 0.f;
        if (detrend == CMC_DETREND_CONSTANT) {
            if (kw == 0) {
                float sa = 0.f, sb = 0.f;
#pragma unroll
                for (int b = 0; b < B0; ++b)
#pragma unroll
                    for (int r = 0; r < R0; ++r) {
                        sa += va[b][r].x + va[b][r].y;
                        sb += vb[b][r].x + vb[b][r].y;
                    }
                // lanes with equal (lane & 3) hold the same channel pair: fixed-order butterfly reduction
#pragma unroll
                for (int off = 16; off >= 4; off >>= 1) {
                    sa += __shfl_xor_sync(0xffffffffu, sa, off);
                    sb += __shfl_xor_sync(0xffffffffu, sb, off);
                }
                if ((tid & 31) < 4) {
                    part[(tid >> 5) * kTmaCT + 2 * (tid & 31)] = sa;
                    part[(tid >> 5) * kTmaCT + 2 * (tid & 31) + 1] = sb;
                }
                __syncthreads();
                if (tid < kTmaCT) {
                    float t = 0.f;
                    for (int w = 0; w < NT / 32; ++w) t += part[w * kTmaCT + tid];
                    mean_s[tid] = t * (1.0f / N);
                }
                __syncthreads();
            }
            mua = mean_s[2 * (tid & 3)];
            mub = mean_s[2 * (tid & 3) + 1];
        }
        // every thread has its raw samples in registers: the tile may now be overwritten in place
        __syncthreads();
#pragma unroll
        for (int b = 0; b < B0; ++b) {
            const int q = tid + b * NT;
            const int cp = q & 3, j = q >> 2;
#pragma unroll
            for (int r = 0; r < R0; ++r) {
                const float2 w = __ldg(reinterpret_cast<const float2*>(win + 2 * (j + r * S0)));
                va[b][r] = make_float2((va[b][r].x - mua) * w.x, (va[b][r].y - mua) * w.y);
                vb[b][r] = make_float2((vb[b][r].x - mub) * w.x, (vb[b][r].y - mub) * w.y);
            }
            dft<R0>(va[b]);
            dft<R0>(vb[b]);
            // rows j * R0 + r: the swizzle bit is (j & 1) when R0 == 16, and r only enters through its low bit
#pragma unroll
            for (int r = 0; r < R0; ++r)
                sts128(sbuf + pt_off(j * R0 + r, cp), make_float4(va[b][r].x, va[b][r].y, vb[b][r].x, vb[b][r].y));
        }
        __syncthreads();
        if (R1 > 1) pass_pair<M, (R1 > 1 ? R1 : 2), R0>(sbuf, twM, tid);
        if (R2 > 1) pass_pair<M, (R2 > 1 ? R2 : 2), R0 * R1>(sbuf, twM, tid);

        // ---- real-FFT split for the requested bins; 2 channels = 16 bytes per lane ----
        float2* out = spec + ((int64_t)(seg * n_win + kw) * F) * spec_ld + c0;
        for (int q = tid; q < F * 4; q += NT) {
            const int cp = q & 3, bi = q >> 2;
            const int b = bin_lo + bi;
            const float4 A = lds128(sbuf + pt_off(b & (M - 1), cp));
            const float4 Bz = lds128(sbuf + pt_off((M - b) & (M - 1), cp));
            const float2 w = __ldg(twN + b);
            float2 X[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float ax = h ? A.z : A.x, ay = h ? A.w : A.y, bx = h ? Bz.z : Bz.x, by = h ? Bz.w : Bz.y;
                const float2 E = make_float2(0.5f * (ax + bx), 0.5f * (ay - by));
                const float2 O = make_float2(0.5f * (ax - bx), 0.5f * (ay + by));
                const float2 T = cmul(w, O);
                X[h] = make_float2(E.x + T.y, E.y - T.x);
                if (b == 0 || b == M) X[h].y = 0.f;
                if (detrend == CMC_DETREND_POST_TAPER && b == 0) X[h].x = 0.f;
            }
            float2* o = out + (int64_t)bi * spec_ld + 2 * cp;
            const int c = c0 + 2 * cp;
            if (c + 1 < n_ch && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                *reinterpret_cast<float4*>(o) = make_float4(X[0].x, X[0].y, X[1].x, X[1].y);
            } else {
                if (c < n_ch) o[0] = X[0];
                if (c + 1 < n_ch) o[1] = X[1];
            }
        }
        __syncthreads();
    }
}

template <int M, bool SWZ>
static int launch_tma(const CUtensorMap& tmap, int n_ch, const int64_t* seg_starts, int n_seg, const float* windows,
                      int n_win, int detrend, int bin_lo, int F, float2* spec, int64_t spec_ld, const float2* twM,
                      const float2* twN, cudaStream_t st) {
    constexpr int NT = M / 4;
    const size_t smem = 1024 + (size_t)M * 64 + sizeof(float) * ((NT / 32) * kTmaCT + kTmaCT) + 16;
    auto kern = fft_segments_tma_kernel<M, SWZ>;
    int rc = ensure_smem_attr(reinterpret_cast<const void*>(kern), smem);
    if (rc) return rc;
    dim3 grid(n_seg, (n_ch + kTmaCT - 1) / kTmaCT);
    static const int dbg = getenv("CMC_FFT_DBG") ? atoi(getenv("CMC_FFT_DBG")) : 0;
    kern<<<grid, NT, smem, st>>>(tmap, n_ch, seg_starts, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN, dbg);
    CMC_CHECK_LAUNCH("fft_segments_tma_kernel");
    return CMC_OK;
}

// Returns 1 when the request does not qualify for the TMA path (caller falls back to fft_segments_kernel).
int fft_segments_tma(const float* x, int64_t n_samples, int n_ch, int64_t ld, const int64_t* seg_starts, int n_seg,
                     const float* windows, int n_win, int N, int detrend, int bin_lo, int F, float2* spec,
                     int64_t spec_ld, const float2* twM, const float2* twN, cudaStream_t st) {
    if (N != 512 && N != 1024 && N != 2048 && N != 4096) return 1;
    if ((ld & 3) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0 || n_samples >= (1ll << 31)) return 1;
    EncodeTiledFn enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    // SWIZZLE_64B with a 32-byte inner box faults on sm_100 (illegal memory access, found on hardware with
    // scripts/debug_fft.py), so the raw tile is loaded unswizzled: first-pass reads are then 2-way bank
    // conflicted (4 instead of 2 wavefronts per request), every other access is conflict free.
    constexpr bool swz = false;
    CUtensorMap tmap;
    cuuint64_t dims[2] = {(cuuint64_t)n_ch, (cuuint64_t)n_samples};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)kTmaCT, 256u};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(recording) failed with CUresult %d", (int)r);
        return CMC_ECUDA;
    }
    switch (N) {
        case 512: return launch_tma<256, false>(tmap, n_ch, seg_starts, n_seg, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN, st);
        case 1024: return launch_tma<512, false>(tmap, n_ch, seg_starts, n_seg, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN, st);
        case 2048: return launch_tma<1024, false>(tmap, n_ch, seg_starts, n_seg, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN, st);
        case 4096: return launch_tma<2048, false>(tmap, n_ch, seg_starts, n_seg, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN, st);
    }
    return 1;
}

}  // namespace cmc

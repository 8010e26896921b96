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

#include <map>
#include <mutex>
#include <utility>
#include "tc_common.cuh"
#include "csd_layout.cuh"
#include "tile_counter.cuh"
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

// barrier over `n` threads with hardware barrier `id` (id 0, n = blockDim.x is __syncthreads)
__device__ __forceinline__ void group_sync(int id, int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// CP = channel pairs per tile row: 4 (tiles of 8 channels, 32-byte raw rows) or 2 (tiles of 4 channels, 16-byte raw
// rows - half-size tiles for the four-worker pipeline).
template <bool SWZ, int CP = 4>
__device__ __forceinline__ uint32_t raw_off(int n, int cp) {
    const uint32_t lin = (uint32_t)n * (8u * CP) + (uint32_t)cp * 8u;
    return (SWZ && CP == 4) ? (lin ^ (((lin >> 7) & 3u) << 4)) : lin;
}
// byte offset of complex point p (channel pair cp) in the interleaved work layout: rows of CP * 16 bytes; the low
// row bits are XOR-ed with the 16-row group index so that first-pass stores (row stride 16) spread over all banks:
// one bit for 64-byte rows (two rows per 128 bytes), two bits for 32-byte rows (four rows per 128 bytes).  The
// swizzle commutes with adding multiples of kSwzPeriod rows.
template <int CP = 4>
__device__ __forceinline__ uint32_t pt_off(int p, int cp) {
    return (uint32_t)(p ^ ((p >> 4) & (CP == 4 ? 1 : 3))) * (16u * CP) + (uint32_t)cp * 16u;
}
template <int CP> struct SwzPeriod { static constexpr int value = CP == 4 ? 32 : 64; };

// The row swizzle commutes with adding multiples of SwzPeriod rows, so it is applied once per butterfly whenever the
// pass stride is such a multiple (all strides after the first pass for M >= 512 with 8-channel tiles).
template <int M, int R, int NS, int CP, class TW>
__device__ __forceinline__ void pass_pair(uint32_t sbuf, const TW& twM, int tid, int bar_id) {
    constexpr int NT = M * CP / 16;
    constexpr int B = kPtsPerThread / R;
    constexpr int STRIDE = M / R;
    constexpr int P = SwzPeriod<CP>::value;
    constexpr uint32_t RB = 16u * CP;              // bytes per point row
    float2 va[B][R], vb[B][R];
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int q = tid + b * NT;
        const int cp = q & (CP - 1), j = q / CP;
        const uint32_t base = sbuf + pt_off<CP>(j, cp);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 t = lds128(STRIDE % P == 0 ? base + (uint32_t)(r * STRIDE) * RB
                                                    : sbuf + pt_off<CP>(j + r * STRIDE, cp));
            va[b][r] = make_float2(t.x, t.y);
            vb[b][r] = make_float2(t.z, t.w);
        }
        const int k = j % NS;
        apply_twiddles2<R>(va[b], vb[b], twM, k * (M / (NS * R)));
        dft<R>(va[b]);
        dft<R>(vb[b]);
    }
    group_sync(bar_id, NT);
#pragma unroll
    for (int b = 0; b < B; ++b) {
        const int q = tid + b * NT;
        const int cp = q & (CP - 1), j = q / CP;
        const int k = j % NS;
        const int j0 = (j / NS) * (NS * R) + k;
        const uint32_t base = sbuf + pt_off<CP>(j0, cp);
#pragma unroll
        for (int r = 0; r < R; ++r)
            sts128(NS % P == 0 ? base + (uint32_t)(r * NS) * RB : sbuf + pt_off<CP>(j0 + r * NS, cp),
                   make_float4(va[b][r].x, va[b][r].y, vb[b][r].x, vb[b][r].y));
    }
    group_sync(bar_id, NT);
}

// Transforms the raw tile that sits at `sbuf` (N rows x 8 channels, written by TMA) in place and writes the
// requested bins of window row `kw`.  Executed by the NT = M * CP / 16 threads that synchronise on barrier `bar_id`.
#ifdef CMC_K1_PROFILE
// instrumented build only: per-phase clock64() totals of thread 0 of worker 0 of every CTA
__device__ unsigned long long g_k1_cycles[8];
#define K1_TICK(i) do { if (tid == 0 && bar_id == 1) { const long long now_ = clock64(); \
    atomicAdd(&g_k1_cycles[i], (unsigned long long)(now_ - tick_)); tick_ = now_; } } while (0)
#define K1_TICK_INIT long long tick_ = clock64()
#else
#define K1_TICK(i) do { } while (0)
#define K1_TICK_INIT do { } while (0)
#endif

struct NoPoll {
    __device__ __forceinline__ void operator()() const {}
};

// TAB: the window rows, twM and the requested entries of twN have been staged in shared memory at `tab_s`
// (layout: twM [M], twN [F], windows [n_win][N] floats); otherwise they are read from global memory.
template <int M, bool SWZ, class Poll, bool TAB = false, int CP = 4>
__device__ __forceinline__ void process_tile(const Poll& poll, uint32_t sbuf, float* part, float* mean_s, int tid, int bar_id, int kw,
                                             int seg, int c0, int n_ch, const float* __restrict__ windows, int n_win,
                                             int detrend, int bin_lo, int F, float2* __restrict__ spec, int64_t spec_ld,
                                             const float2* __restrict__ twM_g, const float2* __restrict__ twN_g,
                                             uint32_t tab_s = 0) {
    const Tab<TAB> twM{twM_g, tab_s};
    const Tab<TAB> twN{twN_g, tab_s + 8u * (uint32_t)(M - bin_lo)};                       // indexed by the bin number
    const Tab<TAB> win2{reinterpret_cast<const float2*>(windows + (int64_t)kw * (2 * M)),
                        tab_s + 8u * (uint32_t)(M + F) + 4u * (uint32_t)(kw * 2 * M)};    // pairs (w[2p], w[2p+1])
    constexpr int N = 2 * M;
    constexpr int NT = M * CP / 16;
    constexpr int CT = 2 * CP;
    constexpr int P = SwzPeriod<CP>::value;
    constexpr uint32_t RB = 16u * CP;
    constexpr int R0 = Plan<M>::R0, R1 = Plan<M>::R1, R2 = Plan<M>::R2;
    constexpr int B0 = kPtsPerThread / R0;
    constexpr int S0 = M / R0;
    K1_TICK_INIT;
    float2 va[B0][R0], vb[B0][R0];
#pragma unroll
    for (int b = 0; b < B0; ++b) {
        const int q = tid + b * NT;
        const int cp = q & (CP - 1), j = q / CP;
        // raw rows are CP * 8 bytes and a lane group of CP owns one row, so a half-warp that read the even rows of
        // consecutive points would hit only half of the banks: every other group of 8 / CP points reads its odd
        // row first
        const int sw = (j >> (CP == 4 ? 1 : 2)) & 1;
#pragma unroll
        for (int r = 0; r < R0; ++r) {
            const int n0 = 2 * (j + r * S0);
            const float2 first = lds64(sbuf + raw_off<SWZ, CP>(n0 + sw, cp));
            const float2 second = lds64(sbuf + raw_off<SWZ, CP>(n0 + 1 - sw, cp));
            const float2 re = sw ? second : first;                      // x[2p][c], x[2p][c+1]
            const float2 im = sw ? first : second;                      // x[2p+1][c], x[2p+1][c+1]
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
            // lanes with equal (lane % CP) hold the same channel pair: fixed-order butterfly reduction
#pragma unroll
            for (int off = 16; off >= CP; off >>= 1) {
                sa += __shfl_xor_sync(0xffffffffu, sa, off);
                sb += __shfl_xor_sync(0xffffffffu, sb, off);
            }
            if ((tid & 31) < CP) {
                part[(tid >> 5) * CT + 2 * (tid & 31)] = sa;
                part[(tid >> 5) * CT + 2 * (tid & 31) + 1] = sb;
            }
            group_sync(bar_id, NT);
            // every thread folds the warp partials of its own channel pair (same order everywhere): no second barrier
            float ta = 0.f, tb = 0.f;
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) {
                ta += part[w * CT + 2 * (tid & (CP - 1))];
                tb += part[w * CT + 2 * (tid & (CP - 1)) + 1];
            }
            mua = ta * (1.0f / N);
            mub = tb * (1.0f / N);
            if (tid < CP) {                                // kept for the other tapers of this segment
                mean_s[2 * tid] = mua;
                mean_s[2 * tid + 1] = mub;
            }
        } else {
            mua = mean_s[2 * (tid & (CP - 1))];
            mub = mean_s[2 * (tid & (CP - 1)) + 1];
        }
    }
    // every thread has its raw samples in registers: the tile may now be overwritten in place (the barrier of the mean
    // reduction already says so when it ran)
    if (!(detrend == CMC_DETREND_CONSTANT && kw == 0)) group_sync(bar_id, NT);
    K1_TICK(1);
    poll();
#pragma unroll
    for (int b = 0; b < B0; ++b) {
        const int q = tid + b * NT;
        const int cp = q & (CP - 1), j = q / CP;
#pragma unroll
        for (int r = 0; r < R0; ++r) {
            const float2 w = win2(j + r * S0);
            va[b][r] = make_float2((va[b][r].x - mua) * w.x, (va[b][r].y - mua) * w.y);
            vb[b][r] = make_float2((vb[b][r].x - mub) * w.x, (vb[b][r].y - mub) * w.y);
        }
        dft<R0>(va[b]);
        dft<R0>(vb[b]);
#pragma unroll
        for (int r = 0; r < R0; ++r)
            sts128(sbuf + pt_off<CP>(j * R0 + r, cp), make_float4(va[b][r].x, va[b][r].y, vb[b][r].x, vb[b][r].y));
    }
    group_sync(bar_id, NT);
    K1_TICK(2);
    poll();
    // ---- band-limited request: when every requested bin lies below the stride NSL = M / RL of the last pass, that
    // pass only has to produce output 0 of butterfly b (Z[b]) and output RL - 1 of butterfly NSL - b (Z[M - b]).
    // Both are plain twiddled sums of RL inputs, so the pass, its barrier, its stores and the reload of the split
    // collapse into one step: item = (bin, channel pair), Z[b] = sum_r W_M^(r b) y[b + r NSL] and
    // Z[M - b] = sum_r conj(W_M^(r b)) y[NSL - b + r NSL] accumulated with FMAs straight into the split.
    // (Pairing bins b and NSL - b in one item reads every row once instead of ~1.6 times but leaves one long item
    // per thread: 50.2 us against 48.7 us for this form, config 2.) ----
    constexpr int RL = R2 > 1 ? R2 : R1;
    constexpr int NSL = M / RL;
    float2* out = spec + ((int64_t)(seg * n_win + kw) * F) * spec_ld + c0;
    if (RL > 1 && bin_lo + F <= NSL) {
        if (R2 > 1) pass_pair<M, (R1 > 1 ? R1 : 2), R0, CP>(sbuf, twM, tid, bar_id);
        K1_TICK(3);
        poll();
        for (int q = tid; q < F * CP; q += NT) {
            const int cp = q & (CP - 1), bi = q / CP;
            const int b = bin_lo + bi;
            const int kk = (NSL - b) & (NSL - 1);
            const uint32_t pa = sbuf + pt_off<CP>(b, cp), pb = sbuf + pt_off<CP>(kk, cp);
            float4 A = lds128(pa), Bz = lds128(pb);
#pragma unroll
            for (int r = 1; r < RL; ++r) {
                const float2 w = twM(r * b);
                const float4 ya = lds128(NSL % P == 0 ? pa + (uint32_t)(r * NSL) * RB : sbuf + pt_off<CP>(b + r * NSL, cp));
                const float4 yb = lds128(NSL % P == 0 ? pb + (uint32_t)(r * NSL) * RB : sbuf + pt_off<CP>(kk + r * NSL, cp));
                A.x = fmaf(-w.y, ya.y, fmaf(w.x, ya.x, A.x));
                A.y = fmaf(w.y, ya.x, fmaf(w.x, ya.y, A.y));
                A.z = fmaf(-w.y, ya.w, fmaf(w.x, ya.z, A.z));
                A.w = fmaf(w.y, ya.z, fmaf(w.x, ya.w, A.w));
                Bz.x = fmaf(w.y, yb.y, fmaf(w.x, yb.x, Bz.x));
                Bz.y = fmaf(-w.y, yb.x, fmaf(w.x, yb.y, Bz.y));
                Bz.z = fmaf(w.y, yb.w, fmaf(w.x, yb.z, Bz.z));
                Bz.w = fmaf(-w.y, yb.z, fmaf(w.x, yb.w, Bz.w));
            }
            const float2 w = twN(b);
            float2 X[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float ax = h ? A.z : A.x, ay = h ? A.w : A.y;
                const float bx = h ? Bz.z : Bz.x, by = h ? Bz.w : Bz.y;
                const float2 E = make_float2(0.5f * (ax + bx), 0.5f * (ay - by));
                const float2 O = make_float2(0.5f * (ax - bx), 0.5f * (ay + by));
                const float2 T = cmul(w, O);
                X[h] = make_float2(E.x + T.y, E.y - T.x);
                if (b == 0) X[h].y = 0.f;
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
        group_sync(bar_id, NT);
        K1_TICK(5);
        return;
    }
    if (R1 > 1) pass_pair<M, (R1 > 1 ? R1 : 2), R0, CP>(sbuf, twM, tid, bar_id);
    K1_TICK(3);
    poll();
    if (R2 > 1) pass_pair<M, (R2 > 1 ? R2 : 2), R0 * R1, CP>(sbuf, twM, tid, bar_id);
    K1_TICK(4);
    poll();

    // ---- real-FFT split for the requested bins.  One item = (bin, half of the channel tile): two channel pairs
    // with independent loads and arithmetic, so that the F * 2 items of a band-limited request fit one round of
    // the worker's threads instead of a full round plus a mostly idle one ----
    for (int q = tid; q < F * (CP / 2); q += NT) {
        const int half = q % (CP / 2), bi = q / (CP / 2);
        const int b = bin_lo + bi;
        const float2 w = twN(b);
        float4 A[2], Bz[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            A[u] = lds128(sbuf + pt_off<CP>(b & (M - 1), 2 * half + u));
            Bz[u] = lds128(sbuf + pt_off<CP>((M - b) & (M - 1), 2 * half + u));
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int cp = 2 * half + u;
            float2 X[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float ax = h ? A[u].z : A[u].x, ay = h ? A[u].w : A[u].y;
                const float bx = h ? Bz[u].z : Bz[u].x, by = h ? Bz[u].w : Bz[u].y;
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
    }
    group_sync(bar_id, NT);
    K1_TICK(5);
}

#ifdef CMC_K1_PROFILE
}  // namespace cmc
extern "C" CMC_API int cmc_dbg_k1_cycles(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, cmc::g_k1_cycles, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {}; cudaMemcpyToSymbol(cmc::g_k1_cycles, z, sizeof(z)); }
    return 0;
}
namespace cmc {
#endif

// one thread: raw tile of (segment start, channel tile) -> shared memory, N rows of 32 bytes in boxes of 256 rows
template <int M, int CP = 4>
__device__ __forceinline__ void issue_tile_tma(unsigned char* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int start) {
    constexpr int N = 2 * M;
    constexpr int ROW = 8 * CP;                // bytes per raw row
    fence_proxy_async();                       // earlier generic-proxy accesses before async-proxy writes
    mbar_arrive_expect_tx(bar, N * ROW);
#pragma unroll 1
    for (int i = 0; i < N / 256; ++i) tma_load_2d(dst + i * 256 * ROW, tmap, bar, c0, start + i * 256);
}

// Basic kernel: one CTA per (segment, channel tile), 2 CTAs per SM.
template <int M, bool SWZ>
__global__ void __launch_bounds__(M / 4, (M / 4) <= 256 ? 2 : 1)
fft_segments_tma_kernel(const __grid_constant__ CUtensorMap tmap, int n_ch,
                        const int64_t* __restrict__ seg_starts, const float* __restrict__ windows, int n_win,
                        int detrend, int bin_lo, int F, float2* __restrict__ spec, int64_t spec_ld,
                        const float2* __restrict__ twM, const float2* __restrict__ twN) {
    constexpr int N = 2 * M;
    constexpr int NT = M / 4;
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
        if (tid == 0) issue_tile_tma<M>(buf, &tmap, bar, c0, start);
        mbar_wait(bar, kw & 1);
        process_tile<M, SWZ>(NoPoll(), sbuf, part, mean_s, tid, 0, kw, seg, c0, n_ch, windows, n_win, detrend, bin_lo, F,
                             spec, spec_ld, twM, twN);
    }
}

// Pipelined kernel: one persistent CTA per SM with NW workers of NT threads and NW + NW / 2 tile buffers.  Each worker
// transforms its current tile while an idle buffer receives the next tile of whichever worker claims it first; when
// a worker moves on to its prefetched buffer it hands the old one back.  The basic kernel serialises the TMA wait
// and the butterflies inside a CTA (53.9 us = 0.75 x (40.6 compute + 25.4 load), profiles/r01b); here the load of
// tile t + 1 overlaps the transform of tile t.
//   CP = 4: tiles of 8 channels (64 KB at N = 2048), two 256-thread workers, three buffers, one prefetch in flight;
//   CP = 2: tiles of 4 channels (32 KB), FOUR 128-thread workers, six buffers, two prefetches in flight - the
//           worker barriers span 4 warps instead of 8 and three other workers can issue while one waits
//           (profiles/r01b addendum 8: barrier stalls 1.9 and load waits 0.6 cycles per issued instruction).
template <int NB>
struct PipeCtrlT {
    uint64_t full[NB];
    unsigned free_mask;      // bit b set: buffer b is idle
    unsigned fills[NB];      // number of TMA fills issued into each buffer (parity of the next wait)
    int next_buf[4];         // per worker: prefetched buffer or -1
    unsigned next_parity[4];
    long long next_tile[4];  // per worker: tile claimed for the next round
};

// Tiles beyond the first one of every worker are claimed from a device-wide counter, so workers that start late
// (their SM was still busy with another kernel) or run slower simply take fewer tiles.  The last worker to leave
// resets the counter, so it is reusable by the next STREAM-ORDERED launch (and by CUDA-graph replays, whose kernel
// arguments are frozen).  Two launches that may overlap must never share a counter: eager launches own one slot per
// (device, stream) - launches of one stream are ordered - and every launch recorded during stream capture gets a
// fresh slot of its own (parallel graph branches, graphs replayed beside eager work); when a pool runs out the
// launch uses the fixed tile stride instead (tile_counter_for).
constexpr int kTileCounterEagerSlots = 64;       // distinct (device, stream) pairs with a claim counter
constexpr int kTileCounterCaptureSlots = 960;   // kernel nodes recorded into CUDA graphs
__device__ TileCounter g_tile_counters[kTileCounterEagerSlots + kTileCounterCaptureSlots];

// Host side of the rule above.  Returns nullptr (= fixed stride) when no private counter is available.
TileCounter* tile_counter_for(int dev, cudaStream_t st) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, int> eager;      // (device, stream) -> slot
    static std::map<int, int> n_eager, n_capture;                  // per device
    TileCounter* slots = nullptr;
    if (cudaGetSymbolAddress(reinterpret_cast<void**>(&slots), g_tile_counters) != cudaSuccess) return nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (cap != cudaStreamCaptureStatusNone) {
        int& n = n_capture[dev];
        if (n >= kTileCounterCaptureSlots) return nullptr;
        return slots + kTileCounterEagerSlots + n++;
    }
    auto it = eager.find({dev, st});
    if (it == eager.end()) {
        int& n = n_eager[dev];
        if (n >= kTileCounterEagerSlots) return nullptr;
        it = eager.emplace(std::make_pair(dev, st), n++).first;
    }
    return slots + it->second;
}

__device__ __forceinline__ void worker_leave(TileCounter* ctr, int tid, int n_workers) {
    if (tid == 0 && ctr) {
        __threadfence();
        if (atomicAdd(&ctr->done, 1u) == (unsigned)n_workers * gridDim.x - 1u) {   // every worker has claimed its last tile
            ctr->next = 0u;
            ctr->done = 0u;
            __threadfence();
        }
    }
}

// The kernel serves ONE or TWO recordings that share the segment table, the window rows and the bin range (EEG and
// EMG of one subject-condition): a tile index runs over (segment, channel tile of recording 0 | 1), so the two
// modalities share one launch - one prologue, one tail of partly idle SMs, one claim counter - instead of two.
struct PipeSecond {
    int n_ch;                // channels of the second recording (0 = none)
    float2* spec;            // its output base (channel 0 of the second recording)
};

__device__ __forceinline__ int claim_free(unsigned* mask) {
    unsigned m = *reinterpret_cast<volatile unsigned*>(mask);
    while (m) {
        const int b = __ffs(m) - 1;
        const unsigned old = atomicAnd(mask, ~(1u << b));
        if (old & (1u << b)) return b;
        m = old & ~(1u << b);
    }
    return -1;
}

template <int M, bool TAB, int CP>
__global__ void __launch_bounds__((M * CP / 16) * (8 / CP), 1)
fft_segments_tma_pipe_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap2,
                             int n_ch, const PipeSecond second, int n_seg,
                             const int64_t* __restrict__ seg_starts, const float* __restrict__ windows, int n_win,
                             int detrend, int bin_lo, int F, float2* __restrict__ spec, int64_t spec_ld,
                             const float2* __restrict__ twM, const float2* __restrict__ twN, TileCounter* ctr) {
    constexpr int N = 2 * M;
    constexpr int CT = 2 * CP;
    constexpr int NT = M * CP / 16;            // threads per worker
    constexpr int NW = 8 / CP;                 // workers per CTA (NW * NT = M / 2 threads)
    constexpr int NB = NW + NW / 2;            // tile buffers
    constexpr int kTileBytes = N * CT * 4;
    using PipeCtrl = PipeCtrlT<NB>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    float* part_all = reinterpret_cast<float*>(base + NB * kTileBytes);         // [NW][NT / 32][CT]
    float* mean_all = part_all + NW * (NT / 32) * CT;                           // [NW][CT]
    PipeCtrl* ctl = reinterpret_cast<PipeCtrl*>(mean_all + NW * CT);
    // TAB: twM [M], the requested twN entries [F] and the window rows [n_win][N] live in shared memory for the whole
    // launch (the CTA is persistent), so the per-tile table loads are shared-memory round trips
    float2* tab = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(ctl) + ((sizeof(PipeCtrl) + 15) & ~size_t(15)));
    const uint32_t tab_s = TAB ? smem_u32(tab) : 0u;
    const int worker = threadIdx.x / NT;
    const int tid = threadIdx.x - worker * NT;
    const int bar_id = 1 + worker;
    float* part = part_all + worker * (NT / 32) * CT;
    float* mean_s = mean_all + worker * CT;
    const int n_ct0 = (n_ch + CT - 1) / CT;
    const int n_ct = n_ct0 + (second.n_ch + CT - 1) / CT;
    const long long total = (long long)n_seg * n_ct;
    const long long first_wave = (long long)NW * gridDim.x;   // tiles handed out statically
    long long t = (long long)NW * blockIdx.x + worker;        // first tile is static, the rest come from the counter
    // tile -> (recording, first channel): channel tiles [0, n_ct0) belong to the first recording
    auto tile_map = [&](long long tile, int& c0) -> const CUtensorMap* {
        const int ct = (int)(tile % n_ct);
        c0 = (ct < n_ct0 ? ct : ct - n_ct0) * CT;
        return ct < n_ct0 ? &tmap : &tmap2;
    };

    if (threadIdx.x == 0) {
        for (int b = 0; b < NB; ++b) {
            mbar_init(&ctl->full[b], 1);
            ctl->fills[b] = 0;
        }
        ctl->free_mask = ((1u << NB) - 1u) & ~((1u << NW) - 1u);     // buffers NW .. NB - 1 start idle
        for (int w = 0; w < NW; ++w) ctl->next_buf[w] = -1;
        fence_barrier_init();
        tma_prefetch_desc(&tmap);
        if (second.n_ch) tma_prefetch_desc(&tmap2);
        // the first tile of every worker is on its way before the tables are staged
        for (int w = 0; w < NW; ++w) {
            const long long tw = (long long)NW * blockIdx.x + w;
            if (tw < total) {
                int c0f;
                const CUtensorMap* mf = tile_map(tw, c0f);
                issue_tile_tma<M, CP>(base + w * kTileBytes, mf, &ctl->full[w], c0f, (int)seg_starts[tw / n_ct]);
                ctl->fills[w] = 1;
            }
        }
    }
    if (TAB) {
        for (int i = threadIdx.x; i < M; i += blockDim.x) tab[i] = __ldg(twM + i);
        for (int i = threadIdx.x; i < F; i += blockDim.x) tab[M + i] = __ldg(twN + bin_lo + i);
        float* wtab = reinterpret_cast<float*>(tab + M + F);
        for (int i = threadIdx.x; i < n_win * N; i += blockDim.x) wtab[i] = __ldg(windows + i);
    }
    __syncthreads();
    if (t >= total) {                          // this worker has no tile; its buffer stays unused
        worker_leave(ctr, tid, NW);
        return;
    }

    int cur = worker;
    unsigned cur_parity = 0;
    long long tn = total, tnn = total;         // thread 0 of the worker: tiles claimed for the next two rounds
    if (tid == 0) tn = ctr ? first_wave + atomicAdd(&ctr->next, 1u) : t + first_wave;
    while (true) {
        const int seg = (int)(t / n_ct);
        int c0;
        const CUtensorMap* cur_map = tile_map(t, c0);
        const bool is2 = cur_map == &tmap2;
        const int cur_nch = is2 ? second.n_ch : n_ch;
        float2* cur_spec = is2 ? second.spec : spec;
        if (tid == 0) {
            ctl->next_tile[worker] = tn;       // read by the whole worker after the tile's last barrier
            // claim one round ahead: the atomic's round trip hides behind this tile
            tnn = !ctr ? tn + first_wave : (tn < total ? first_wave + atomicAdd(&ctr->next, 1u) : total);
        }
        for (int kw = 0; kw < n_win; ++kw) {
            if (kw > 0 && tid == 0) {          // the in-place transform consumed the raw tile: fetch it again
                cur_parity = ctl->fills[cur] & 1;
                ctl->fills[cur] += 1;
                issue_tile_tma<M, CP>(base + cur * kTileBytes, cur_map, &ctl->full[cur], c0, (int)seg_starts[seg]);
            }
            if (kw > 0) {
                group_sync(bar_id, NT);
                cur_parity = (ctl->fills[cur] - 1) & 1;
            }
#ifdef CMC_K1_PROFILE
            long long w0_ = clock64();
#endif
            mbar_wait(&ctl->full[cur], cur_parity);
#ifdef CMC_K1_PROFILE
            if (tid == 0 && worker == 0) atomicAdd(&g_k1_cycles[0], (unsigned long long)(clock64() - w0_));
#endif
            // thread 0 of the worker tries to claim an idle buffer for the worker's next tile - here and again after
            // every pass barrier, so that a buffer released by another worker mid-tile is picked up at once
            auto poll = [&]() {
                if (tid == 0 && kw == n_win - 1 && tn < total && ctl->next_buf[worker] < 0) {
                    const int fb = claim_free(&ctl->free_mask);
                    if (fb >= 0) {
                        ctl->next_parity[worker] = ctl->fills[fb] & 1;
                        ctl->fills[fb] += 1;
                        int c0n;
                        const CUtensorMap* mn = tile_map(tn, c0n);
                        issue_tile_tma<M, CP>(base + fb * kTileBytes, mn, &ctl->full[fb], c0n, (int)seg_starts[tn / n_ct]);
                        ctl->next_buf[worker] = fb;
                    }
                }
            };
            poll();
            process_tile<M, false, decltype(poll), TAB, CP>(poll, smem_u32(base + cur * kTileBytes), part, mean_s, tid,
                                                            bar_id, kw, seg, c0, cur_nch, windows, n_win, detrend, bin_lo,
                                                            F, cur_spec, spec_ld, twM, twN, tab_s);
        }
        // process_tile ended with a worker barrier: the buffer is no longer read and ctl->next_* is visible
        t = ctl->next_tile[worker];
        if (t >= total) {
            if (tid == 0) {
                __threadfence_block();
                atomicOr(&ctl->free_mask, 1u << cur);              // let another worker prefetch into it
            }
            worker_leave(ctr, tid, NW);
            break;
        }
        tn = tnn;
        const int nb = ctl->next_buf[worker];
        if (nb >= 0) {
            const unsigned np = ctl->next_parity[worker];
            group_sync(bar_id, NT);                                 // everyone has read next_* before it is reset
            if (tid == 0) {
                ctl->next_buf[worker] = -1;
                __threadfence_block();
                atomicOr(&ctl->free_mask, 1u << cur);              // hand the old buffer back
            }
            cur = nb;
            cur_parity = np;
        } else {
            // no prefetch happened (the other workers held the idle buffers): reload in place
            if (tid == 0) {
                ctl->fills[cur] += 1;
                int c0r;
                const CUtensorMap* mr = tile_map(t, c0r);
                issue_tile_tma<M, CP>(base + cur * kTileBytes, mr, &ctl->full[cur], c0r, (int)seg_starts[t / n_ct]);
            }
            group_sync(bar_id, NT);
            cur_parity = (ctl->fills[cur] - 1) & 1;
        }
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
    kern<<<grid, NT, smem, st>>>(tmap, n_ch, seg_starts, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN);
    CMC_CHECK_LAUNCH("fft_segments_tma_kernel");
    return CMC_OK;
}

template <int M, int CP>
static int launch_tma_pipe(const CUtensorMap& tmap, int n_ch, const CUtensorMap& tmap2, const PipeSecond& second,
                           const int64_t* seg_starts, int n_seg, const float* windows,
                           int n_win, int detrend, int bin_lo, int F, float2* spec, int64_t spec_ld, const float2* twM,
                           const float2* twN, cudaStream_t st) {
    constexpr int CT = 2 * CP, NT = M * CP / 16, NW = 8 / CP, NB = NW + NW / 2;
    const size_t smem_base = 1024 + (size_t)NB * (2 * M) * CT * 4 + sizeof(float) * NW * ((NT / 32) * CT + CT) +
                             sizeof(PipeCtrlT<NB>) + 32;
    // twiddles, requested split twiddles and window rows next to the tile buffers when they fit
    const size_t tab_bytes = ((size_t)M + F) * 8 + (size_t)n_win * 2 * M * 4;
    const bool tab = smem_base + tab_bytes <= 227 * 1024;
    const size_t smem = smem_base + (tab ? tab_bytes : 0);
    auto kern = tab ? fft_segments_tma_pipe_kernel<M, true, CP> : fft_segments_tma_pipe_kernel<M, false, CP>;
    int rc = ensure_smem_attr(reinterpret_cast<const void*>(kern), smem);
    if (rc) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (long long)n_seg * ((n_ch + CT - 1) / CT + (second.n_ch + CT - 1) / CT);
    const long long grid = (tiles + NW - 1) / NW < sms ? (tiles + NW - 1) / NW : sms;
    // claim counter private to this launch's stream / graph node (zero-initialised device memory, reset by the kernel)
    static const bool static_tiles = getenv("CMC_FFT_STATIC_TILES") != nullptr;      // fixed stride instead of claims
    TileCounter* ctr = static_tiles ? nullptr : tile_counter_for(dev, st);
    kern<<<(unsigned)grid, NW * NT, smem, st>>>(tmap, tmap2, n_ch, second, n_seg, seg_starts, windows, n_win, detrend,
                                                bin_lo, F, spec, spec_ld, twM, twN, ctr);
    CMC_CHECK_LAUNCH("fft_segments_tma_pipe_kernel");
    return CMC_OK;
}

static bool tma_layout_ok(const float* x, int64_t n_samples, int64_t ld) {
    return (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && n_samples < (1ll << 31);
}

static int make_recording_map(CUtensorMap* tmap, const float* x, int64_t n_samples, int n_ch, int64_t ld, int ct) {
    EncodeTiledFn enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    // SWIZZLE_64B with a 32-byte inner box faults on sm_100 (illegal memory access, found on hardware with
    // a probe script in round 1), so the raw tile is loaded unswizzled: first-pass reads are then 2-way bank
    // conflicted (4 instead of 2 wavefronts per request), every other access is conflict free.
    cuuint64_t dims[2] = {(cuuint64_t)n_ch, (cuuint64_t)n_samples};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)ct, 256u};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(recording) failed with CUresult %d", (int)r);
        return CMC_ECUDA;
    }
    return CMC_OK;
}

// Returns 1 when the request does not qualify for the TMA path (caller falls back to fft_segments_kernel).
// x2 != nullptr: a second recording (n_ch2 channels, pitch ld2, same length / segments / windows / bins) transformed
// by the same launch into spec2; only the pipelined kernel takes pairs (returns 1 otherwise, the caller then issues
// two single calls).
int fft_segments_tma(const float* x, int64_t n_samples, int n_ch, int64_t ld, const int64_t* seg_starts, int n_seg,
                     const float* windows, int n_win, int N, int detrend, int bin_lo, int F, float2* spec,
                     int64_t spec_ld, const float2* twM, const float2* twN, cudaStream_t st, const float* x2, int n_ch2,
                     int64_t ld2, float2* spec2) {
    if (N != 512 && N != 1024 && N != 2048 && N != 4096) return 1;
    if (!tma_layout_ok(x, n_samples, ld)) return 1;
    if (x2 && (!tma_layout_ok(x2, n_samples, ld2) || N == 4096)) return 1;
    // CMC_FFT_CT4=1: four half-size workers for N = 2048 (tiles of 4 channels).  Measured 93.8 us against 90.1 us for
    // the default two 8-channel workers on config 2 (both modalities in one launch): twice the tiles, 16-byte TMA
    // rows and half-width output stores cost more than the cheaper barriers and the second prefetch gain.
    static const bool ct4 = getenv("CMC_FFT_CT4") != nullptr;
    static const bool no_pipe = getenv("CMC_FFT_NO_PIPE") != nullptr;
    const int ct = (N == 2048 && ct4 && !no_pipe) ? 4 : kTmaCT;
    const long long n_tiles = (long long)n_seg * ((n_ch + ct - 1) / ct + (x2 ? (n_ch2 + ct - 1) / ct : 0));
    const bool pipe = !no_pipe && N <= 2048 && n_tiles >= 64 * (kTmaCT / ct);
    CUtensorMap tmap, tmap2;
    int rc = make_recording_map(&tmap, x, n_samples, n_ch, ld, pipe ? ct : kTmaCT);
    if (rc) return rc;
    PipeSecond second{0, nullptr};
    if (x2) {
        if (!pipe) return 1;
        if ((rc = make_recording_map(&tmap2, x2, n_samples, n_ch2, ld2, ct))) return rc;
        second.n_ch = n_ch2;
        second.spec = spec2;
    } else {
        tmap2 = tmap;
    }
    // pipelined persistent kernel whenever the tile buffers fit into shared memory (N <= 2048) and there is enough work
    if (pipe) {
        switch (N) {
            case 512:  return launch_tma_pipe<256, 4>(tmap, n_ch, tmap2, second, seg_starts, n_seg, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN, st);
            case 1024: return launch_tma_pipe<512, 4>(tmap, n_ch, tmap2, second, seg_starts, n_seg, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN, st);
            case 2048:
                if (ct == 4) return launch_tma_pipe<1024, 2>(tmap, n_ch, tmap2, second, seg_starts, n_seg, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN, st);
                return launch_tma_pipe<1024, 4>(tmap, n_ch, tmap2, second, seg_starts, n_seg, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN, st);
            default: break;
        }
    }
    if (x2) return 1;
    switch (N) {
        case 512: return launch_tma<256, false>(tmap, n_ch, seg_starts, n_seg, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN, st);
        case 1024: return launch_tma<512, false>(tmap, n_ch, seg_starts, n_seg, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN, st);
        case 2048: return launch_tma<1024, false>(tmap, n_ch, seg_starts, n_seg, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN, st);
        case 4096: return launch_tma<2048, false>(tmap, n_ch, seg_starts, n_seg, windows, n_win, detrend, bin_lo, F, spec, spec_ld, twM, twN, st);
    }
    return 1;
}

}  // namespace cmc

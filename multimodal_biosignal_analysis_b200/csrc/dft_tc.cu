// K1t: band-limited hann-windowed Welch spectra on the tensor cores (tcgen05.mma kind::f16, BF16 x 3).
// This file holds the unfolded kernel (dft_hann_tc_kernel, also as CTA pairs) and, further down, the folded kernel
// that is the default (dft_hann_fold_kernel); the header describes what both share.
//
// The Welch path of the reference is scipy.signal.coherence with its defaults (preprocessing.py:1228-1230): periodic
// hann window, 50 % overlap, per-segment constant detrend - and the callers only keep a narrow band (1 - 100 Hz =
// 100 of the 1,025 bins at nperseg 2048).  A full FFT per segment computes ten times the bins that are kept and
// transforms every sample twice.  This kernel computes exactly what is kept, and every sample once:
//
//   * half blocks.  With hop N / 2 a segment is two half blocks of N / 2 samples, and each half block belongs to two
//     segments.  For the rectangular-window DFT, R_s[b] = P_h[b] + (-1)^b P_{h+1}[b] with the HALF-BLOCK sums
//     P_h[b] = sum_{n < N/2} x[h N/2 + n] exp(-2 pi i b n / N)        (one GEMM row per half block).
//   * hann in the frequency domain.  w[n] = 1/2 - 1/2 cos(2 pi n / N) is a three-tap filter over bins:
//     X_s[b] = 1/2 R_s[b] - 1/4 (R_s[b-1] + R_s[b+1]).  Removing the segment mean only changes R_s[0] (-> 0).
//   * so ONE GEMM per half block, D[channel][(bin, re/im)] = sum_n x[n][channel] W[(bin, re/im)][n] with the
//     constant table W = (cos, -sin)(2 pi b n / N), M = 128 channels (64 EEG + 64 EMG), N = 208 columns = 104 bins,
//     K = N / 2 samples, and an epilogue that emits  1/2 P[b] -+ 1/4 (P[b-1] + P[b+1])  into the two segments the
//     half block belongs to (second-half emission carries the sign (-1)^b).
//
// Precision.  Operands are split into BF16 hi + lo (16 significand bits) and contracted with three MMAs
// (lo*hi + hi*lo + hi*hi, FP32 accumulation in TMEM): relative error ~1e-5 of the spectrum's rms, i.e. <= ~1e-6 on a
// coherence - inside the 1e-4 gate.  Before the split every channel is shifted by the first sample of its chain of
// overlapping segments (y = x - c; exact for b != 0 because both halves of a segment use the same c, undone for
// b = 0), so a DC offset does not eat the 16 bits, and by a short mean of the half block itself (c_h, added back per
// odd bin in the epilogue), so the 1 / b leakage of a half-block mean does not sit in the accumulators.
//
// Determinism without a zero-fill.  Every output element receives exactly two contributions, from two consecutive
// half blocks of a chain.  Half blocks of even chain position STORE, odd ones ADD (red.global.add.v2.f32); add-phase
// epilogues wait on a device counter until every store-phase unit of the launch has finished (store units come first
// in every CTA's static schedule, all CTAs are co-resident: no deadlock).  store + one add is order independent.
//
// Roles (512 threads, one persistent CTA per SM): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-7 epilogue (TMEM lane = channel), warps 8-15 converters.  Unfolded kernel: the raw FP32 tile
// [64 samples][128 channels] lands where the A operand will live (ring of three 32 KB stages); the converters read it
// into registers, meet at a barrier and write the K-major BF16 hi / lo planes (manual 128-byte swizzle) in place; W
// hi / lo k-blocks (52 KB) arrive by TMA on a ring of two.  Two 208-column accumulators (TMEM columns 0 and 256)
// overlap the epilogue with the next unit.  The folded kernel (its own comment further down) halves K and the table.
#include "common.cuh"
#include "csd_layout.cuh"
#include "tc_common.cuh"
#include "tile_counter.cuh"

#include <cuda_bf16.h>
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <numeric>
#include <vector>

namespace cmc {

using namespace tc;

constexpr int kDtThreads = 512;
constexpr int kDtConvThreads = 256;                          // warps 8-15
constexpr int kDtStagesA = 3;                                // raw tile / A operand ring
constexpr int kDtStagesB = 2;                                // W k-block ring
constexpr int kDtStagesBPair = 4;                            // W ring of the CTA-pair variant (each CTA stages half a k-block)
constexpr int kDtKB = 64;                                    // samples per k-block: 128-byte BF16 rows (SWIZZLE_128B).  32 (64-byte rows,
                                                             // SWIZZLE_64B, rings of 5 + 4) is also implemented and measured: same speed
constexpr int kDtRowB = kDtKB * 2;                           // bytes per operand row of a k-block
constexpr int kDtSpt = kDtKB / 2;                            // samples per converter thread and k-block
constexpr int kDtM = 128;                                    // channels per unit (two 64-channel groups)
constexpr int kDtBins = 104;                                 // accumulator bins
constexpr int kDtCols = 2 * kDtBins;                         // 208 accumulator columns (re, im interleaved)
constexpr int kDtPlaneA = kDtM * kDtKB * 2;                  // 16 KB: one BF16 plane of the A k-block
constexpr int kDtABytes = 2 * kDtPlaneA;                     // 32 KB = the raw FP32 tile [64][128]
constexpr int kDtPlaneB = kDtCols * kDtKB * 2;               // one BF16 plane of the W k-block
constexpr int kDtBBytes = 2 * kDtPlaneB;                     // 52 KB
constexpr int kDtChainCap = 32;                              // segments per chain (bounds the drift y = x - c sees)
constexpr int kDtPrefetch = 4;                               // k-blocks the L2 prefetch of the raw tiles runs ahead
constexpr int kDtSlots = 4;                                  // units whose per-channel offsets are alive at once

struct DtItem {           // one half block
    int x_row;            // first sample
    int c_row;            // sample whose value is subtracted first (first sample of the chain: exact for b != 0)
    int seg_a;            // segment whose FIRST half this is, or -1
    int seg_b;            // segment whose SECOND half this is, or -1
    int phase;            // 0: emissions are stores, 1: emissions are adds
    int pad[3];
};

struct DtParams {
    const DtItem* items;
    const float* e1im;             // [104] unfolded kernel: Im of E1[b] = sum_{n < N/2} exp(-2 pi i b n / N) for odd b (Re = 1),
                                   // 0 for even b; folded kernel: E1c[b] = sum_{k < N/4} 2 cos(theta_b (k + 1/2))
    const float* rot;              // [104][2] folded kernel: (cos phi_b, sin phi_b), phi_b = theta_b (N/4 - 1/2)
    int n_items, n_gp;             // half blocks, channel-group pairs per half block (units = n_items * n_gp)
    int n_store_units;             // units with phase 0 (they come first)
    int KB, N, bin_lo, F, b0, detrend;
    const float* x[2];
    long long ld[2];
    int n_ch[2];
    int n_grp0;                    // 64-channel groups of recording 0 (groups of recording 1 follow)
    float2* spec[2];
    long long spec_ld;
    TileCounter* ctr;
    int pf;                        // k-blocks the L2 prefetch of the raw tiles runs ahead (0 = off)
    int dbg;                       // developer switches (CMC_DT_DBG): 1 no conversion, 2 no MMAs, 4 no emissions, 8 no W loads after the first two
};

struct __align__(8) DtBarriers {
    uint64_t full_a[kDtStagesA];
    uint64_t conv[kDtStagesA];
    uint64_t empty_a[kDtStagesA];
    uint64_t full_b[kDtStagesBPair];
    uint64_t empty_b[kDtStagesBPair];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

// UMMA shared-memory descriptor of a K-major BF16 operand k-block: rows of kDtRowB bytes, 8-row swizzle atoms
__device__ __forceinline__ uint64_t dt_desc(uint32_t smem_addr) {
    if (kDtKB == 64) return make_smem_desc_k_sw128(smem_addr);
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);          // start address  [0,14)
    d |= static_cast<uint64_t>(1) << 16;                             // leading byte offset (ignored)
    d |= static_cast<uint64_t>(512 >> 4) << 32;                      // stride byte offset: 8 rows x 64 bytes
    d |= static_cast<uint64_t>(1) << 46;                             // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(4) << 61;                             // SWIZZLE_64B
    return d;
}

// ------------------------------------------------------------------ CTA pair (cta_group::2) primitives
__device__ __forceinline__ uint32_t dt_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `addr` (a shared::cta window address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dt_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void dt_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void dt_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are signalled on an mbarrier of the pair's LEADER (cluster address `mbar`)
__device__ __forceinline__ void dt_tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t mbar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void dt_umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all earlier MMAs of this thread have completed) on the barrier at the same offset in BOTH CTAs
__device__ __forceinline__ void dt_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void dt_tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dt_tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// developer trace (CMC_DT_DBG & 32): globaltimer stamps of CTA 0, read back with cmc_dbg_dt_trace
__device__ unsigned long long g_dt_trace[64];
__device__ __forceinline__ void dt_stamp(const DtParams& p, int slot) {
    if ((p.dbg & 32) && blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_dt_trace[slot] = t;
    }
}

__device__ __forceinline__ float dt_lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void dt_sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned dt_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void dt_emit(float2* o, float re, float im, int add) {
    if (add)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(o), "f"(re), "f"(im) : "memory");
    else
        asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(o), "f"(re), "f"(im) : "memory");
}
__device__ __forceinline__ void dt_prefetch_l2(const CUtensorMap* m, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
// predicated emission (no branch): `on` != 0 -> store or add
template <bool ADD>
__device__ __forceinline__ void dt_emit_if(float2* o, float re, float im, int on) {
    if (ADD)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %3, 0;\n@p red.global.add.v2.f32 [%0], {%1, %2};\n}\n"
                     ::"l"(o), "f"(re), "f"(im), "r"(on) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %3, 0;\n@p st.global.v2.f32 [%0], {%1, %2};\n}\n"
                     ::"l"(o), "f"(re), "f"(im), "r"(on) : "memory");
}
__device__ __forceinline__ float4 dt_lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
// BF16 hi / lo split of two values: hi = rn(v), lo = rn(v - hi); element 0 in the low half-word
__device__ __forceinline__ void dt_split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - __low2float(h), v1 - __high2float(h));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// group g of the launch -> (recording, first channel); groups past the last read as zeros (rec 0, channels out of range)
__device__ __forceinline__ void dt_group(const DtParams& p, int g, int& rec, int& c0) {
    const int n_grp1 = (p.n_ch[1] + 63) >> 6;
    if (g < p.n_grp0) { rec = 0; c0 = g * 64; }
    else if (g < p.n_grp0 + n_grp1) { rec = 1; c0 = (g - p.n_grp0) * 64; }
    else { rec = 0; c0 = p.n_grp0 * 64; }
}

// 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void dt_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// tcgen05.wait::ld that the compiler cannot move uses of `r` across (the registers are operands of the wait)
__device__ __forceinline__ void dt_tmem_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// P[j] += ch * E1[b0 + j] for the 8 bins of accumulator chunk cq; e1s[j] = (1, Im E1) for odd bins, (0, 0) for even
__device__ __forceinline__ void dt_fixup(uint32_t (&v)[16], int cq, float ch, const float2* e1s) {
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const float2 e = e1s[8 * cq + jj];
        v[2 * jj] = __float_as_uint(fmaf(ch, e.x, __uint_as_float(v[2 * jj])));
        v[2 * jj + 1] = __float_as_uint(fmaf(ch, e.y, __uint_as_float(v[2 * jj + 1])));
    }
}

// Three-tap hann over the 104 accumulator bins of one TMEM lane and the two emissions of the half block: a rolled
// loop over 8-bin chunks (the fully unrolled form was 100 KB of straight-line code and ran at instruction-fetch
// speed: 15 - 25 us per unit).  oA / oB point at output bin j = 0 of the two target segments (or are null).
template <bool ADD>
__device__ __forceinline__ void dt_epilogue_bins(uint32_t taddr, float chh, const float2* e1s, int j_lo, int j_hi,
                                                 float dc_add, bool dc_is_bin0, bool dc_zero_all, bool post_taper,
                                                 float sgn_even, float2* oA, float2* oB, long long ld, bool no_emit) {
    uint32_t cur[16], nxt[16];
    dt_tmem_ld16(taddr, cur);
    dt_tmem_wait16(cur);
    dt_fixup(cur, 0, chh, e1s);
    float pm_re = 0.f, pm_im = 0.f;              // P[j - 1] of the first bin of the chunk
    if (dc_is_bin0) {                            // accumulator bin 0 is the DC bin
        cur[0] = dc_zero_all ? 0u : __float_as_uint(__uint_as_float(cur[0]) + dc_add);
        cur[1] = 0u;
        pm_re = __uint_as_float(cur[2]);         // P[-1] = conj(P[1])
        pm_im = -__uint_as_float(cur[3]);
    }
    constexpr int kChunks = kDtBins / 8;         // 13
    auto bin = [&](int cq, int jj) {
        const int j = 8 * cq + jj;
        const float pc_re = __uint_as_float(cur[2 * jj]), pc_im = __uint_as_float(cur[2 * jj + 1]);
        const float pp_re = __uint_as_float(jj < 7 ? cur[2 * jj + 2] : nxt[0]);
        const float pp_im = __uint_as_float(jj < 7 ? cur[2 * jj + 3] : nxt[1]);
        const float qm_re = jj > 0 ? __uint_as_float(cur[2 * jj - 2]) : pm_re;
        const float qm_im = jj > 0 ? __uint_as_float(cur[2 * jj - 1]) : pm_im;
        const float s_re = 0.25f * (qm_re + pp_re), s_im = 0.25f * (qm_im + pp_im);
        float a_re = fmaf(0.5f, pc_re, -s_re), a_im = fmaf(0.5f, pc_im, -s_im);
        const float sg = (jj & 1) ? -sgn_even : sgn_even;                 // (-1)^b of the second-half emission
        float b_re = sg * fmaf(0.5f, pc_re, s_re), b_im = sg * fmaf(0.5f, pc_im, s_im);
        if (jj == 0 && cq == 0 && dc_is_bin0 && post_taper) a_re = b_re = 0.f;   // periodogram: DC bin zeroed
        if (j >= j_lo && j < j_hi && !no_emit) {
            if (oA) dt_emit(oA, a_re, a_im, ADD);
            if (oB) dt_emit(oB, b_re, b_im, ADD);
        }
        if (oA) oA += ld;
        if (oB) oB += ld;
    };
#pragma unroll 1
    for (int cq = 0; cq < kChunks; ++cq) {
        // the load of the next chunk is in flight while bins 0 - 6 of this one are emitted
        if (cq < kChunks - 1) dt_tmem_ld16(taddr + 16 * (cq + 1), nxt);
#pragma unroll
        for (int jj = 0; jj < 7; ++jj) bin(cq, jj);
        if (cq < kChunks - 1) {
            dt_tmem_wait16(nxt);
            dt_fixup(nxt, cq + 1, chh, e1s);
        }
        bin(cq, 7);
        pm_re = __uint_as_float(cur[14]);
        pm_im = __uint_as_float(cur[15]);
#pragma unroll
        for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
    }
}

// PAIR: two CTAs of a cluster (one TPC) run their units in lockstep as ONE tcgen05.mma.cta_group::2 of M = 256: each
// CTA converts its own raw tile (A rows of its unit) and stages HALF of the W k-block (104 of the 208 rows); the
// tensor cores of the pair share the two halves.  Per CTA and k-block that is 26 KB instead of 52 KB of W written by
// TMA and 39 KB instead of 78 KB of W read by the tensor core - the main loop is bound by shared-memory bandwidth.
// The leader (cluster rank 0) issues the MMAs; its barriers collect the pair's arrivals (conv: 512 converter threads,
// full_b: the bytes of both halves, tmem_empty: 8 epilogue warps); tcgen05.commit multicasts to both CTAs.
template <bool PAIR>
__global__ void __launch_bounds__(kDtThreads, 1)
dft_hann_tc_kernel(const __grid_constant__ CUtensorMap mX0, const __grid_constant__ CUtensorMap mX1,
                   const __grid_constant__ CUtensorMap mW, const DtParams p) {
    constexpr int kStagesB = PAIR ? kDtStagesBPair : kDtStagesB;
    constexpr int kPlaneB = PAIR ? kDtPlaneB / 2 : kDtPlaneB;       // W rows this CTA stages per plane and k-block
    constexpr int kBBytes = 2 * kPlaneB;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = base;                                        // [kDtStagesA][32 KB]
    unsigned char* sB = base + kDtStagesA * kDtABytes;               // [kStagesB][kBBytes]
    float* csum = reinterpret_cast<float*>(sB + kStagesB * kBBytes); // [2][128] partial sums of the first k-block
    float* coff = csum + 2 * kDtM;                                   // [kDtSlots][128] per-unit channel offsets c_h
    float2* e1s = reinterpret_cast<float2*>(coff + kDtSlots * kDtM); // [112] E1 of the accumulator bins (odd bins only)
    DtBarriers* bars = reinterpret_cast<DtBarriers*>(e1s + 112);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_units = p.n_items * p.n_gp;
    const uint32_t rank = PAIR ? dt_cta_rank() : 0u;                 // 0 = leader of the pair
    // rounds of this CTA (pair): q = q0, q0 + q_step, ...; its unit in round q is 2 q + rank (PAIR) or q
    const int q0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int q_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int n_rounds = PAIR ? (n_units + 1) / 2 : n_units;
    if (threadIdx.x == 0) dt_stamp(p, 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kDtStagesA; ++s) {
            mbar_init(&bars->full_a[s], 1);
            mbar_init(&bars->conv[s], PAIR ? 2 * kDtConvThreads : kDtConvThreads);
            mbar_init(&bars->empty_a[s], 1);
        }
        for (int s = 0; s < kStagesB; ++s) {
            mbar_init(&bars->full_b[s], 1);
            mbar_init(&bars->empty_b[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->tmem_full[a], 1);
            mbar_init(&bars->tmem_empty[a], PAIR ? 8 : 4);
        }
        fence_barrier_init();
        tma_prefetch_desc(&mX0);
        tma_prefetch_desc(&mX1);
        tma_prefetch_desc(&mW);
    }
    if (warp == 2) {
        if (PAIR) dt_tmem_alloc_pair(&bars->tmem_base, 512);
        else {
            tmem_alloc(&bars->tmem_base, 512);
            tmem_relinquish();
        }
    }
    if (threadIdx.x >= 128 && threadIdx.x < 128 + 112) {
        const int j = threadIdx.x - 128;
        const bool odd = j < kDtBins && ((p.b0 + j) & 1);
        e1s[j] = odd ? make_float2(1.0f, __ldg(p.e1im + j)) : make_float2(0.f, 0.f);
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) dt_cluster_sync();                 // the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    if (threadIdx.x == 0) dt_stamp(p, 1);

    if (warp == 0) {
        // ===================== TMA producer (raw tiles and W k-blocks on separate rings) =====================
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            for (int q = q0; q < n_rounds; q += q_step) {
                const int u = PAIR ? 2 * q + (int)rank : q;
                const bool live = u < n_units;                       // the last pair may run one empty unit
                const DtItem it = p.items[(live ? u : 0) / p.n_gp];
                const int gp = u % p.n_gp;
                int ra, ca, rb, cb;
                dt_group(p, 2 * gp, ra, ca);
                dt_group(p, 2 * gp + 1, rb, cb);
                const CUtensorMap* ma = ra ? &mX1 : &mX0;
                const CUtensorMap* mb = rb ? &mX1 : &mX0;
                for (int kb = 0; kb < p.KB; ++kb) {
                    const int row = it.x_row + kb * kDtKB;
                    if (live && p.pf && kb + p.pf < p.KB) {          // optional L2 prefetch (CMC_DT_PF)
                        dt_prefetch_l2(ma, ca, row + p.pf * kDtKB);
                        dt_prefetch_l2(mb, cb, row + p.pf * kDtKB);
                    }
                    mbar_wait(&bars->empty_a[sa], pa ^ 1);
                    if (q == q0 && kb < 16) dt_stamp(p, 2 + kb);
                    if (live) {
                        mbar_arrive_expect_tx(&bars->full_a[sa], kDtABytes);
                        unsigned char* st = sA + sa * kDtABytes;
                        tma_load_2d(st, ma, &bars->full_a[sa], ca, row);
                        tma_load_2d(st + kDtPlaneA, mb, &bars->full_a[sa], cb, row);
                    } else {
                        mbar_arrive(&bars->full_a[sa]);
                    }
                    if (++sa == kDtStagesA) { sa = 0; pa ^= 1; }
                    mbar_wait(&bars->empty_b[sb], pb ^ 1);
                    unsigned char* sw = sB + sb * kBBytes;
                    if (PAIR) {
                        // both CTAs load their half of the W k-block; the bytes of both are counted on the leader's barrier
                        if (rank == 0) mbar_arrive_expect_tx(&bars->full_b[sb], 2 * kBBytes);
                        const uint32_t lead_bar = dt_mapa(smem_u32(&bars->full_b[sb]), 0);
                        dt_tma_load_2d_pair(sw, &mW, lead_bar, kb * kDtKB, (int)rank * (kDtCols / 2));
                        dt_tma_load_2d_pair(sw + kPlaneB, &mW, lead_bar, kb * kDtKB, kDtCols + (int)rank * (kDtCols / 2));
                    } else if ((p.dbg & 8) && (q != q0 || kb >= kStagesB)) {
                        mbar_arrive(&bars->full_b[sb]);            // timing experiment: stale W, no load
                    } else {
                        mbar_arrive_expect_tx(&bars->full_b[sb], kBBytes);
                        tma_load_2d(sw, &mW, &bars->full_b[sb], kb * kDtKB, 0);
                        tma_load_2d(sw + kPlaneB, &mW, &bars->full_b[sb], kb * kDtKB, kDtCols);
                    }
                    if (++sb == kStagesB) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * kDtM : kDtM, kDtCols);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0, n = 0;
            for (int q = q0; q < n_rounds; q += q_step) {
                const uint32_t acc = n & 1, accphase = (n >> 1) & 1;
                mbar_wait(&bars->tmem_empty[acc], accphase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * 256;
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(&bars->full_b[sb], pb);              // W planes (async proxy) have landed
                    mbar_wait(&bars->conv[sa], pa);                // A planes written by the converters
                    if (q == q0 && kb < 16) dt_stamp(p, 18 + kb);
                    tc_fence_after();
                    const uint32_t ahi = smem_u32(sA + sa * kDtABytes), alo = ahi + kDtPlaneA;
                    const uint32_t bhi = smem_u32(sB + sb * kBBytes), blo = bhi + kPlaneB;
#pragma unroll
                    for (int k = 0; k < kDtKB / 16; ++k) {
                        if (p.dbg & 2) break;
                        const uint64_t dah = dt_desc(ahi + k * 32), dal = dt_desc(alo + k * 32);
                        const uint64_t dbh = dt_desc(bhi + k * 32), dbl = dt_desc(blo + k * 32);
                        if (PAIR) {
                            dt_umma_f16_pair(d, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);
                            dt_umma_f16_pair(d, dah, dbl, idesc, 1u);
                            dt_umma_f16_pair(d, dah, dbh, idesc, 1u);
                        } else {
                            umma_f16(d, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);      // small terms first
                            umma_f16(d, dah, dbl, idesc, 1u);
                            umma_f16(d, dah, dbh, idesc, 1u);
                        }
                    }
                    if (PAIR) {
                        dt_commit_pair(&bars->empty_a[sa]);
                        dt_commit_pair(&bars->empty_b[sb]);
                    } else {
                        umma_commit(&bars->empty_a[sa]);
                        umma_commit(&bars->empty_b[sb]);
                    }
                    if (++sa == kDtStagesA) { sa = 0; pa ^= 1; }
                    if (++sb == kStagesB) { sb = 0; pb ^= 1; }
                }
                if (PAIR) dt_commit_pair(&bars->tmem_full[acc]);
                else umma_commit(&bars->tmem_full[acc]);
                ++n;
            }
        }
    } else if (warp >= 8) {
        // ===================== converters: y = x - c0 - c_h, BF16 hi / lo planes in place =====================
        const int t = threadIdx.x - 256;
        const int m = t & 127;                   // A row = channel of the unit
        const int g = t >> 7;                    // samples 32 g .. 32 g + 31 of the k-block
        int sa = 0;
        uint32_t pa = 0, n = 0;
        for (int q = q0; q < n_rounds; q += q_step) {
            const int u = PAIR ? 2 * q + (int)rank : q;
            const bool live = u < n_units;
            const DtItem it = p.items[(live ? u : 0) / p.n_gp];
            const int gp = u % p.n_gp;
            int rec, c0;
            dt_group(p, 2 * gp + (m >> 6), rec, c0);
            const int ch = c0 + (m & 63);
            const float* xr = rec ? p.x[1] : p.x[0];
            const long long ldr = rec ? p.ld[1] : p.ld[0];
            // c0: first sample of the chain (the same for both halves of every segment, so it drops out of every bin
            // but DC).  c_h: mean of the half block's first 64 samples of x - c0, added back in the epilogue; it keeps
            // the 1 / b leakage of a half-block mean out of the accumulators.
            float c = ch < (rec ? p.n_ch[1] : p.n_ch[0]) ? __ldg(xr + (long long)it.c_row * ldr + ch) : 0.f;
            for (int kb = 0; kb < p.KB; ++kb) {
                mbar_wait(&bars->full_a[sa], pa);
                if (t == 0 && q == q0 && kb < 16) dt_stamp(p, 34 + kb);
                if ((p.dbg & 1) || !live) {
                    if (PAIR) dt_arrive_cluster(dt_mapa(smem_u32(&bars->conv[sa]), 0));
                    else mbar_arrive(&bars->conv[sa]);
                    if (++sa == kDtStagesA) { sa = 0; pa ^= 1; }
                    continue;
                }
                const uint32_t sbase = smem_u32(sA + sa * kDtABytes);
                const uint32_t src = sbase + (uint32_t)((m >> 6) * kDtPlaneA + (m & 63) * 4 + g * kDtSpt * 256);
                float v[kDtSpt];
#pragma unroll
                for (int i = 0; i < kDtSpt; ++i) v[i] = dt_lds32(src + i * 256) - c;
                if (kb == 0) {
                    float s = 0.f;
#pragma unroll
                    for (int i = 0; i < kDtSpt; ++i) s += v[i];
                    csum[g * kDtM + m] = s;
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");       // every raw value is in a register
                if (kb == 0) {
                    const float chh = (csum[m] + csum[kDtM + m]) * (1.0f / kDtKB);
                    if (g == 0) coff[(n & (kDtSlots - 1)) * kDtM + m] = chh;
                    c += chh;
#pragma unroll
                    for (int i = 0; i < kDtSpt; ++i) v[i] -= chh;
                }
                // K-major row of channel m: 8-row atoms; 16-byte chunk index XOR (row & 7) for 128-byte rows
                // (SWIZZLE_128B), XOR ((row >> 1) & 3) for 64-byte rows (SWIZZLE_64B)
                const uint32_t row = sbase + (uint32_t)((m >> 3) * (8 * kDtRowB) + (m & 7) * kDtRowB);
                const int sw = kDtKB == 64 ? (m & 7) : ((m >> 1) & 3);
#pragma unroll
                for (int q = 0; q < kDtSpt / 8; ++q) {
                    uint4 hi, lo;
                    dt_split2(v[8 * q], v[8 * q + 1], hi.x, lo.x);
                    dt_split2(v[8 * q + 2], v[8 * q + 3], hi.y, lo.y);
                    dt_split2(v[8 * q + 4], v[8 * q + 5], hi.z, lo.z);
                    dt_split2(v[8 * q + 6], v[8 * q + 7], hi.w, lo.w);
                    const uint32_t off = (uint32_t)((((kDtSpt / 8) * g + q) ^ sw) << 4);
                    dt_sts128(row + off, hi);
                    dt_sts128(row + kDtPlaneA + off, lo);
                }
                fence_proxy_async();             // generic-proxy writes -> visible to the MMA's async-proxy reads
                if (PAIR) dt_arrive_cluster(dt_mapa(smem_u32(&bars->conv[sa]), 0));    // the leader issues the pair's MMAs
                else mbar_arrive(&bars->conv[sa]);
                if (++sa == kDtStagesA) { sa = 0; pa ^= 1; }
            }
            ++n;
        }
    } else if (warp >= 4) {
        // ===================== epilogue: three-tap hann, two emissions per half block =====================
        const int q = warp - 4;                  // TMEM lane quadrant
        const int m = threadIdx.x - 128;         // accumulator lane = channel of the unit
        uint32_t n = 0;
        const unsigned store_target = 4u * (unsigned)p.n_store_units;
        for (int q = q0; q < n_rounds; q += q_step) {
            const int u = PAIR ? 2 * q + (int)rank : q;
            const bool live = u < n_units;
            const DtItem it = p.items[(live ? u : 0) / p.n_gp];
            const int gp = u % p.n_gp;
            int rec, c0;
            dt_group(p, 2 * gp + (m >> 6), rec, c0);
            const int ch = c0 + (m & 63);
            const bool valid = live && ch < (rec ? p.n_ch[1] : p.n_ch[0]);
            float2* sp = rec ? p.spec[1] : p.spec[0];
            float2* outA = (valid && it.seg_a >= 0) ? sp + (long long)it.seg_a * p.F * p.spec_ld + ch : nullptr;
            float2* outB = (valid && it.seg_b >= 0) ? sp + (long long)it.seg_b * p.F * p.spec_ld + ch : nullptr;
            float c_first = 0.f;                 // c0 of this channel (only the DC bin of the non-detrended modes needs it)
            if (p.b0 == 0 && p.detrend != CMC_DETREND_CONSTANT && valid)
                c_first = __ldg((rec ? p.x[1] : p.x[0]) + (long long)it.c_row * (rec ? p.ld[1] : p.ld[0]) + ch);
            const uint32_t acc = n & 1, accphase = (n >> 1) & 1;
            mbar_wait(&bars->tmem_full[acc], accphase);
            if (m == 0 && q == q0) dt_stamp(p, 50);
            tc_fence_after();
            const float chh = coff[(n & (kDtSlots - 1)) * kDtM + m];
            if (live && it.phase) {
                // adds may only start once every store of the launch is visible
                if (lane == 0) {
                    const long long t0 = clock64();
                    while (dt_ld_acquire(&p.ctr->next) < store_target) {
                        __nanosleep(64);
                        if (clock64() - t0 > 4000000000LL) {
                            printf("cmc: dft_hann_tc store-phase wait timed out (block %d)\n", blockIdx.x);
                            __trap();
                        }
                    }
                }
                __syncwarp();
            }
            const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
            const int j_lo = p.bin_lo - p.b0;
            const long long back = (long long)j_lo * p.spec_ld;                   // oA / oB address accumulator bin 0
            const bool dc0 = p.b0 == 0;
            const float dc_add = 0.5f * (float)p.N * (c_first + chh);
            const float sgn_even = (p.b0 & 1) ? -1.0f : 1.0f;
            if (it.phase)
                dt_epilogue_bins<true>(taddr, chh, e1s, j_lo, j_lo + p.F, dc_add, dc0, p.detrend == CMC_DETREND_CONSTANT,
                                       p.detrend == CMC_DETREND_POST_TAPER, sgn_even, outA ? outA - back : nullptr,
                                       outB ? outB - back : nullptr, p.spec_ld, (p.dbg & 4) != 0);
            else
                dt_epilogue_bins<false>(taddr, chh, e1s, j_lo, j_lo + p.F, dc_add, dc0, p.detrend == CMC_DETREND_CONSTANT,
                                        p.detrend == CMC_DETREND_POST_TAPER, sgn_even, outA ? outA - back : nullptr,
                                        outB ? outB - back : nullptr, p.spec_ld, (p.dbg & 4) != 0);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {                                      // accumulator may be overwritten
                if (PAIR) dt_arrive_cluster(dt_mapa(smem_u32(&bars->tmem_empty[acc]), 0));
                else mbar_arrive(&bars->tmem_empty[acc]);
            }
            if (m == 0 && q == q0) dt_stamp(p, 51);
            if (live && !it.phase) {
                __threadfence();                 // this thread's stores before the warp's arrival
                __syncwarp();
                if (lane == 0) atomicAdd(&p.ctr->next, 1u);
            }
            ++n;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) dt_cluster_sync();                 // no arrival on the peer's barriers is in flight when a CTA leaves
    if (warp == 2) {
        if (PAIR) dt_tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
    if (threadIdx.x == 0) dt_stamp(p, 52);
    if (threadIdx.x == 0) {
        // the last CTA to leave hands the counter back at zero (stream-ordered launches and graph replays reuse it)
        __threadfence();
        if (atomicAdd(&p.ctr->done, 1u) == gridDim.x - 1u) {
            p.ctr->next = 0u;
            p.ctr->done = 0u;
            __threadfence();
        }
    }
}

// ====================================================================================================================
// K1t, folded form (default).  About its centre a half block splits into an even and an odd part,
//     P_h[b] = e^{-i phi_b} (Cp[b] - i Sm[b]),   phi_b = theta_b (N/4 - 1/2),   theta_b = 2 pi b / N,
//     Cp[b] = sum_{k < N/4} (y[N/4 + k] + y[N/4 - 1 - k]) cos(theta_b (k + 1/2)),
//     Sm[b] = sum_{k < N/4} (y[N/4 + k] - y[N/4 - 1 - k]) sin(theta_b (k + 1/2)),
// so the converters fold the raw tile (one add, one subtract per sample pair, in FP32, before the BF16 split) and the
// GEMM runs over K = N / 4 with the cos table for the sums and the sin table for the differences: two MMAs of N = 112
// instead of one of N = 208 over twice the K - half the tensor work and half the table bytes per sample.  The
// epilogue rotates (Cp, Sm) by phi_b (four FMAs per bin) and continues as the unfolded kernel does.  Stages carry 32
// folded samples (64 raw): raw tile / A planes 32 KB (sum hi, lo, difference hi, lo; 64-byte rows, SWIZZLE_64B),
// table k-block 28 KB; rings of 4 + 3 stages hold 256 / 192 raw samples in flight against 192 / 128 before.
constexpr int kFdKB = 32;                                    // folded samples per k-block (64-byte BF16 rows)
constexpr int kFdRows = 112;                                 // table rows / accumulator columns per set (104 bins + 8 zero rows)
constexpr int kFdStagesL = 2;                                // raw (landing) ring
constexpr int kFdStagesA = 2;                                // operand ring (A planes)
constexpr int kFdStagesB = 3;                                // table ring
constexpr int kFdPlaneA = kDtM * kFdKB * 2;                  // 8 KB
constexpr int kFdABytes = 4 * kFdPlaneA;                     // 32 KB: sum hi, sum lo, difference hi, difference lo
constexpr int kFdPlaneB = kFdRows * kFdKB * 2;               // 7 KB: one set of one plane
constexpr int kFdBBytes = 4 * kFdPlaneB;                     // 28 KB: [hi: cos rows, sin rows][lo: cos rows, sin rows]
constexpr int kFdSpt = kFdKB / 2;                            // folded samples per converter thread and k-block

struct __align__(8) FdBarriers {
    uint64_t full_l[kFdStagesL];     // raw tile landed (TMA bytes)
    uint64_t empty_l[kFdStagesL];    // raw tile read into registers by all 8 converter warps
    uint64_t conv[kFdStagesA];       // A planes written (8 converter warps)
    uint64_t empty_a[kFdStagesA];    // A planes consumed (MMA commit)
    uint64_t full_b[kFdStagesB];
    uint64_t empty_b[kFdStagesB];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

// K-major BF16 operand k-block with 64-byte rows: 8-row atoms of 512 bytes, SWIZZLE_64B
__device__ __forceinline__ uint64_t fd_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(512 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(4) << 61;
    return d;
}

// 32 lanes x 8 consecutive 32-bit columns
__device__ __forceinline__ void fd_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

// (Cp, Sm) of the 8 bins of chunk cq -> (Re P, Im P) with the odd-bin constant added back: Cp += ch * E1c, then the
// rotation by phi.  tab_s: shared-window address of the (cos phi, sin phi, E1c, -) table.
__device__ __forceinline__ void fd_load_rotate(uint32_t taddr, int cq, float (&re)[8], float (&im)[8], float ch,
                                               uint32_t tab_s) {
    uint32_t cp[8], sm[8];
    fd_tmem_ld8(taddr + 8 * cq, cp);
    fd_tmem_ld8(taddr + kFdRows + 8 * cq, sm);
    tmem_ld_wait();
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const float4 t = dt_lds_f4(tab_s + 16u * (uint32_t)(8 * cq + jj));
        const float c = fmaf(ch, t.z, __uint_as_float(cp[jj]));
        const float s = __uint_as_float(sm[jj]);
        re[jj] = fmaf(t.x, c, -t.y * s);
        im[jj] = -fmaf(t.y, c, t.x * s);
    }
}

// oA / oB: output pointers of accumulator bin 0 (any valid address when the emission is off), on_a / on_b: emission on.
// Chunks [c_begin, c_end) of the 13 eight-bin chunks: the four epilogue warps take all of them, except for the LAST
// unit of a CTA, whose epilogue nothing overlaps - there the eight converter warps (idle by then) take two thirds.
template <bool ADD>
__device__ __forceinline__ void fd_epilogue_bins(uint32_t taddr, float chh, uint32_t tab_s, int j_lo, int j_hi,
                                                 float dc_add, bool dc_is_bin0, bool dc_zero_all, bool post_taper,
                                                 float sgn_even, float2* oA, float2* oB, int on_a, int on_b, long long ld,
                                                 int c_begin, int c_end) {
    float cre[8], cim[8], nre[8] = {}, nim[8] = {};
    float pm_re = 0.f, pm_im = 0.f;              // P[j - 1] of the first bin of the chunk
    if (c_begin > 0) {
        fd_load_rotate(taddr, c_begin - 1, cre, cim, chh, tab_s);
        pm_re = cre[7];
        pm_im = cim[7];
        oA += (long long)(8 * c_begin) * ld;
        oB += (long long)(8 * c_begin) * ld;
    }
    fd_load_rotate(taddr, c_begin, cre, cim, chh, tab_s);
    if (dc_is_bin0 && c_begin == 0) {            // accumulator bin 0 is the DC bin
        cre[0] = dc_zero_all ? 0.f : cre[0] + dc_add;
        cim[0] = 0.f;
        pm_re = cre[1];                          // P[-1] = conj(P[1])
        pm_im = -cim[1];
    }
    constexpr int kChunks = kDtBins / 8;         // 13
#pragma unroll 1
    for (int cq = c_begin; cq < c_end; ++cq) {
        if (cq < kChunks - 1) fd_load_rotate(taddr, cq + 1, nre, nim, chh, tab_s);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = 8 * cq + jj;
            const float pp_re = jj < 7 ? cre[jj < 7 ? jj + 1 : 7] : nre[0];
            const float pp_im = jj < 7 ? cim[jj < 7 ? jj + 1 : 7] : nim[0];
            const float qm_re = jj > 0 ? cre[jj > 0 ? jj - 1 : 0] : pm_re;
            const float qm_im = jj > 0 ? cim[jj > 0 ? jj - 1 : 0] : pm_im;
            const float s_re = 0.25f * (qm_re + pp_re), s_im = 0.25f * (qm_im + pp_im);
            float a_re = fmaf(0.5f, cre[jj], -s_re), a_im = fmaf(0.5f, cim[jj], -s_im);
            const float sg = (jj & 1) ? -sgn_even : sgn_even;                 // (-1)^b of the second-half emission
            float b_re = sg * fmaf(0.5f, cre[jj], s_re), b_im = sg * fmaf(0.5f, cim[jj], s_im);
            if (jj == 0 && cq == 0 && dc_is_bin0 && post_taper) a_re = b_re = 0.f;   // periodogram: DC bin zeroed
            const int in_band = (j >= j_lo && j < j_hi) ? 1 : 0;
            dt_emit_if<ADD>(oA, a_re, a_im, in_band & on_a);
            dt_emit_if<ADD>(oB, b_re, b_im, in_band & on_b);
            oA += ld;
            oB += ld;
        }
        pm_re = cre[7];
        pm_im = cim[7];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            cre[i] = nre[i];
            cim[i] = nim[i];
        }
    }
}

// Everything of one unit's epilogue that depends on the lane: output pointers, the DC terms, the wait for the store
// phase, then the chunks [c_begin, c_end).  Called by the epilogue warps for every unit and by the converter warps for
// the last unit of the CTA.
__device__ __forceinline__ void fd_unit_epilogue(const DtParams& p, const DtItem& it, int gp, int m, int lane, uint32_t taddr,
                                                 float chh, uint32_t tab_s, unsigned store_target, int c_begin, int c_end) {
    int rec, c0;
    dt_group(p, 2 * gp + (m >> 6), rec, c0);
    const int ch = c0 + (m & 63);
    const bool valid = ch < (rec ? p.n_ch[1] : p.n_ch[0]);
    float2* sp = rec ? p.spec[1] : p.spec[0];
    float2* outA = (valid && it.seg_a >= 0) ? sp + (long long)it.seg_a * p.F * p.spec_ld + ch : nullptr;
    float2* outB = (valid && it.seg_b >= 0) ? sp + (long long)it.seg_b * p.F * p.spec_ld + ch : nullptr;
    float c_first = 0.f;
    if (p.b0 == 0 && p.detrend != CMC_DETREND_CONSTANT && valid && c_begin == 0)
        c_first = __ldg((rec ? p.x[1] : p.x[0]) + (long long)it.c_row * (rec ? p.ld[1] : p.ld[0]) + ch);
    if (it.phase) {
        // adds may only start once every store of the launch is visible
        if (lane == 0) {
            const long long t0 = clock64();
            while (dt_ld_acquire(&p.ctr->next) < store_target) {
                __nanosleep(64);
                if (clock64() - t0 > 4000000000LL) {
                    printf("cmc: dft_hann_fold store-phase wait timed out (block %d)\n", blockIdx.x);
                    __trap();
                }
            }
        }
        __syncwarp();
    }
    const int j_lo = p.bin_lo - p.b0;
    const long long back = (long long)j_lo * p.spec_ld;
    const bool dc0 = p.b0 == 0;
    const float dc_add = 0.5f * (float)p.N * c_first;        // the c_h part came in through E1c[0] = N / 2
    const float sgn_even = (p.b0 & 1) ? -1.0f : 1.0f;
    // emissions that are off (no such segment, channel out of range) keep a valid pointer and a zero predicate
    const int on_a = (outA != nullptr && !(p.dbg & 4)) ? 1 : 0, on_b = (outB != nullptr && !(p.dbg & 4)) ? 1 : 0;
    float2* pa_ = (outA ? outA : sp) - back;
    float2* pb_ = (outB ? outB : sp) - back;
    if (it.phase)
        fd_epilogue_bins<true>(taddr, chh, tab_s, j_lo, j_lo + p.F, dc_add, dc0, p.detrend == CMC_DETREND_CONSTANT,
                               p.detrend == CMC_DETREND_POST_TAPER, sgn_even, pa_, pb_, on_a, on_b, p.spec_ld, c_begin, c_end);
    else
        fd_epilogue_bins<false>(taddr, chh, tab_s, j_lo, j_lo + p.F, dc_add, dc0, p.detrend == CMC_DETREND_CONSTANT,
                                p.detrend == CMC_DETREND_POST_TAPER, sgn_even, pa_, pb_, on_a, on_b, p.spec_ld, c_begin, c_end);
}

// MC: the two CTAs of a cluster share every table k-block.  Each CTA still runs its own units with its own MMAs
// (cta_group::1); only the table ring is common: k-block g is loaded ONCE, by the CTA of rank g & 1, with a TMA multicast
// that lands it in both CTAs and signals each CTA's own full_b barrier; a table stage is free when the MMAs of BOTH
// CTAs have consumed it (tcgen05.commit multicast to both empty_b barriers, count 2).  Table traffic from L2 halves.
__device__ __forceinline__ void fd_tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    const uint16_t mask = 3;
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void fd_commit_mc(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

template <bool MC>
__global__ void __launch_bounds__(kDtThreads, 1)
dft_hann_fold_kernel(const __grid_constant__ CUtensorMap mX0, const __grid_constant__ CUtensorMap mX1,
                     const __grid_constant__ CUtensorMap mW, const DtParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    // the raw tiles land in their own ring: the converters need no barrier between reading a tile and writing the
    // planes (in place, all 256 threads had to meet in the middle of every k-block and the eight warps ran in lockstep)
    unsigned char* sL = base;                                        // [kFdStagesL][32 KB] raw FP32 tiles
    unsigned char* sA = sL + kFdStagesL * kFdABytes;                 // [kFdStagesA][32 KB] BF16 planes
    unsigned char* sB = sA + kFdStagesA * kFdABytes;                 // [kFdStagesB][28 KB]
    float* csum = reinterpret_cast<float*>(sB + kFdStagesB * kFdBBytes);   // [2][128] partial sums of the first k-block
    float* coff = csum + 2 * kDtM;                                   // [kDtSlots][128] per-unit channel offsets c_h
    float4* tab = reinterpret_cast<float4*>(coff + kDtSlots * kDtM); // [112] (cos phi, sin phi, E1c, 0) per accumulator bin
    FdBarriers* bars = reinterpret_cast<FdBarriers*>(tab + kFdRows);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_units = p.n_items * p.n_gp;
    const int q4 = p.N / 4;
    const int rank = MC ? (int)dt_cta_rank() : 0;
    // rounds of this CTA (pair): q = q0, q0 + q_step, ...; its unit in round q is 2 q + rank (MC) or q.  The last
    // round of an odd unit count leaves the rank-1 CTA without a unit: it still takes part in the table ring.
    const int q0 = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int q_step = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int n_rounds = MC ? (n_units + 1) / 2 : n_units;
    auto unit_of = [&](int q) { return MC ? 2 * q + rank : q; };
    auto is_last = [&](int q) { return q + q_step >= n_rounds || unit_of(q + q_step) >= n_units; };
    if (threadIdx.x == 0) dt_stamp(p, 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kFdStagesL; ++s) {
            mbar_init(&bars->full_l[s], 1);
            mbar_init(&bars->empty_l[s], kDtConvThreads / 32);
        }
        for (int s = 0; s < kFdStagesA; ++s) {
            mbar_init(&bars->conv[s], kDtConvThreads / 32);
            mbar_init(&bars->empty_a[s], 1);
        }
        for (int s = 0; s < kFdStagesB; ++s) {
            mbar_init(&bars->full_b[s], 1);
            mbar_init(&bars->empty_b[s], MC ? 2 : 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->tmem_full[a], 1);
            mbar_init(&bars->tmem_empty[a], 4);
        }
        fence_barrier_init();
        tma_prefetch_desc(&mX0);
        tma_prefetch_desc(&mX1);
        tma_prefetch_desc(&mW);
    }
    if (warp == 2) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    if (threadIdx.x >= 128 && threadIdx.x < 128 + kFdRows) {
        const int j = threadIdx.x - 128;
        tab[j] = j < kDtBins ? make_float4(__ldg(p.rot + 2 * j), __ldg(p.rot + 2 * j + 1), __ldg(p.e1im + j), 0.f)
                             : make_float4(1.f, 0.f, 0.f, 0.f);
    }
    tc_fence_before();
    __syncthreads();
    if (MC) dt_cluster_sync();                   // the peer's barriers exist before a multicast can signal them
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    if (threadIdx.x == 0) dt_stamp(p, 1);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            int gk = 0;                                      // k-blocks of the table ring so far (same in both CTAs)
            for (int q = q0; q < n_rounds; q += q_step) {
                const int u = unit_of(q);
                const bool live = u < n_units;
                const DtItem it = p.items[(live ? u : 0) / p.n_gp];
                const int gp = u % p.n_gp;
                int ra, ca, rb, cb;
                dt_group(p, 2 * gp, ra, ca);
                dt_group(p, 2 * gp + 1, rb, cb);
                const CUtensorMap* ma = ra ? &mX1 : &mX0;
                const CUtensorMap* mb = rb ? &mX1 : &mX0;
                for (int kb = 0; kb < p.KB; ++kb) {
                    // forward tile: samples N/4 + 32 kb ...; mirrored tile: samples N/4 - 32 (kb + 1) ... (read backwards)
                    const int row_f = it.x_row + q4 + kb * kFdKB;
                    const int row_r = it.x_row + q4 - (kb + 1) * kFdKB;
                    if (live) {
                        mbar_wait(&bars->empty_l[sa], pa ^ 1);
                        if (q == q0 && kb < 16) dt_stamp(p, 2 + kb);
                        mbar_arrive_expect_tx(&bars->full_l[sa], kFdABytes);
                        unsigned char* st = sL + sa * kFdABytes;
                        tma_load_2d(st, ma, &bars->full_l[sa], ca, row_f);
                        tma_load_2d(st + kFdPlaneA, mb, &bars->full_l[sa], cb, row_f);
                        tma_load_2d(st + 2 * kFdPlaneA, ma, &bars->full_l[sa], ca, row_r);
                        tma_load_2d(st + 3 * kFdPlaneA, mb, &bars->full_l[sa], cb, row_r);
                        if (++sa == kFdStagesL) { sa = 0; pa ^= 1; }
                    }
                    mbar_wait(&bars->empty_b[sb], pb ^ 1);         // MC: the MMAs of BOTH CTAs have consumed the stage
                    mbar_arrive_expect_tx(&bars->full_b[sb], kFdBBytes);
                    unsigned char* sw = sB + sb * kFdBBytes;
                    if (!MC) {
                        tma_load_2d(sw, &mW, &bars->full_b[sb], kb * kFdKB, 0);                    // hi: cos rows, sin rows
                        tma_load_2d(sw + 2 * kFdPlaneB, &mW, &bars->full_b[sb], kb * kFdKB, 2 * kFdRows);   // lo
                    } else if ((gk & 1) == rank) {                 // this CTA's turn: one load feeds both CTAs
                        fd_tma_load_2d_mc(sw, &mW, &bars->full_b[sb], kb * kFdKB, 0);
                        fd_tma_load_2d_mc(sw + 2 * kFdPlaneB, &mW, &bars->full_b[sb], kb * kFdKB, 2 * kFdRows);
                    }
                    ++gk;
                    if (++sb == kFdStagesB) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(kDtM, kFdRows);
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0, n = 0;
            for (int q = q0; q < n_rounds; q += q_step) {
                if (unit_of(q) >= n_units) {                 // no unit in the pair's last round: keep the table ring turning
                    for (int kb = 0; kb < p.KB; ++kb) {
                        mbar_wait(&bars->full_b[sb], pb);
                        fd_commit_mc(&bars->empty_b[sb]);
                        if (++sb == kFdStagesB) { sb = 0; pb ^= 1; }
                    }
                    break;
                }
                const uint32_t acc = n & 1, accphase = (n >> 1) & 1;
                mbar_wait(&bars->tmem_empty[acc], accphase ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(&bars->full_b[sb], pb);
                    mbar_wait(&bars->conv[sa], pa);
                    if (q == q0 && kb < 16) dt_stamp(p, 18 + kb);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + sa * kFdABytes);
                    const uint32_t b0 = smem_u32(sB + sb * kFdBBytes);
#pragma unroll
                    for (int set = 0; set < 2; ++set) {              // 0: sums x cos table, 1: differences x sin table
                        const uint32_t d = tmem_base + acc * 256 + set * kFdRows;
                        const uint32_t ahi = a0 + set * 2 * kFdPlaneA, alo = ahi + kFdPlaneA;
                        const uint32_t bhi = b0 + set * kFdPlaneB, blo = bhi + 2 * kFdPlaneB;
#pragma unroll
                        for (int k = 0; k < kFdKB / 16; ++k) {
                            const uint64_t dah = fd_desc(ahi + k * 32), dal = fd_desc(alo + k * 32);
                            const uint64_t dbh = fd_desc(bhi + k * 32), dbl = fd_desc(blo + k * 32);
                            umma_f16(d, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);      // small terms first
                            umma_f16(d, dah, dbl, idesc, 1u);
                            umma_f16(d, dah, dbh, idesc, 1u);
                        }
                    }
                    umma_commit(&bars->empty_a[sa]);
                    if (MC) fd_commit_mc(&bars->empty_b[sb]);
                    else umma_commit(&bars->empty_b[sb]);
                    if (++sa == kFdStagesA) { sa = 0; pa ^= 1; }
                    if (++sb == kFdStagesB) { sb = 0; pb ^= 1; }
                }
                umma_commit(&bars->tmem_full[acc]);
                ++n;
            }
        }
    } else if (warp >= 8) {
        // ===================== converters: fold, y = x - c0 - c_h, BF16 hi / lo planes in place =====================
        const int t = threadIdx.x - 256;
        const int m = t & 127;                   // A row = channel of the unit
        const int g = t >> 7;                    // folded samples 16 g .. 16 g + 15 of the k-block
        int sl = 0, sa = 0;
        uint32_t pl = 0, pa = 0, n = 0;
        int u_last = -1;
        for (int q = q0; q < n_rounds; q += q_step) {
            const int u = unit_of(q);
            if (u >= n_units) break;
            u_last = u;
            const DtItem it = p.items[u / p.n_gp];
            const int gp = u % p.n_gp;
            int rec, c0;
            dt_group(p, 2 * gp + (m >> 6), rec, c0);
            const int ch = c0 + (m & 63);
            const float* xr = rec ? p.x[1] : p.x[0];
            const long long ldr = rec ? p.ld[1] : p.ld[0];
            float c = ch < (rec ? p.n_ch[1] : p.n_ch[0]) ? __ldg(xr + (long long)it.c_row * ldr + ch) : 0.f;
            for (int kb = 0; kb < p.KB; ++kb) {
                mbar_wait(&bars->full_l[sl], pl);
                if (t == 0 && q == q0 && kb < 16) dt_stamp(p, 34 + kb);
                const uint32_t lbase = smem_u32(sL + sl * kFdABytes);
                const uint32_t src_f = lbase + (uint32_t)((m >> 6) * kFdPlaneA + (m & 63) * 4 + g * kFdSpt * 256);
                // the mirror of forward row kk is row 31 - kk of the mirrored tile
                const uint32_t src_r = lbase + (uint32_t)(2 * kFdPlaneA + (m >> 6) * kFdPlaneA + (m & 63) * 4 +
                                                          (kFdKB - 1 - g * kFdSpt) * 256);
                float sum[kFdSpt], dif[kFdSpt];
                const float c2 = 2.0f * c;
#pragma unroll
                for (int i = 0; i < kFdSpt; ++i) {
                    const float f = dt_lds32(src_f + i * 256), r = dt_lds32(src_r - i * 256);
                    sum[i] = (f + r) - c2;
                    dif[i] = f - r;
                }
                __syncwarp();                    // every lane holds its samples: the raw tile may be overwritten
                if (t == 0 && q == q0 && kb >= 4 && kb < 7) dt_stamp(p, 53 + (kb - 4));
                if (lane == 0) mbar_arrive(&bars->empty_l[sl]);
                if (++sl == kFdStagesL) { sl = 0; pl ^= 1; }
                if (kb == 0) {
                    // c_h: mean of the 64 samples around the centre of the half block (added back in the epilogue)
                    float s = 0.f;
#pragma unroll
                    for (int i = 0; i < kFdSpt; ++i) s += sum[i];
                    csum[g * kDtM + m] = s;
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                    const float chh = (csum[m] + csum[kDtM + m]) * (1.0f / (2 * kFdKB));
                    if (g == 0) coff[(n & (kDtSlots - 1)) * kDtM + m] = chh;
                    c += chh;
#pragma unroll
                    for (int i = 0; i < kFdSpt; ++i) sum[i] -= 2.0f * chh;
                    asm volatile("bar.sync 2, 256;" ::: "memory");   // csum may be rewritten by the next unit
                }
                mbar_wait(&bars->empty_a[sa], pa ^ 1);             // the MMAs that read these planes have completed
                if (t == 0 && q == q0 && kb >= 4 && kb < 7) dt_stamp(p, 56 + (kb - 4));
                // K-major rows of 64 bytes: 8-row atoms of 512 bytes, 16-byte chunk index XOR ((row >> 1) & 3)
                const uint32_t row = smem_u32(sA + sa * kFdABytes) + (uint32_t)((m >> 3) * 512 + (m & 7) * 64);
                const int sw = (m >> 1) & 3;
#pragma unroll
                for (int q = 0; q < kFdSpt / 8; ++q) {
                    uint4 hi, lo;
                    const uint32_t off = (uint32_t)((((kFdSpt / 8) * g + q) ^ sw) << 4);
                    dt_split2(sum[8 * q], sum[8 * q + 1], hi.x, lo.x);
                    dt_split2(sum[8 * q + 2], sum[8 * q + 3], hi.y, lo.y);
                    dt_split2(sum[8 * q + 4], sum[8 * q + 5], hi.z, lo.z);
                    dt_split2(sum[8 * q + 6], sum[8 * q + 7], hi.w, lo.w);
                    dt_sts128(row + off, hi);
                    dt_sts128(row + kFdPlaneA + off, lo);
                    dt_split2(dif[8 * q], dif[8 * q + 1], hi.x, lo.x);
                    dt_split2(dif[8 * q + 2], dif[8 * q + 3], hi.y, lo.y);
                    dt_split2(dif[8 * q + 4], dif[8 * q + 5], hi.z, lo.z);
                    dt_split2(dif[8 * q + 6], dif[8 * q + 7], hi.w, lo.w);
                    dt_sts128(row + 2 * kFdPlaneA + off, hi);
                    dt_sts128(row + 3 * kFdPlaneA + off, lo);
                }
                fence_proxy_async();             // generic-proxy writes -> visible to the MMA's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->conv[sa]);       // one arrival per warp
                if (t == 0 && q == q0 && kb >= 4 && kb < 7) dt_stamp(p, 59 + (kb - 4));
                if (++sa == kFdStagesA) { sa = 0; pa ^= 1; }
            }
            ++n;
        }
        if (n > 0) {
            // nothing is left to convert: take two thirds of the last unit's epilogue (nothing overlaps it otherwise)
            const DtItem it = p.items[u_last / p.n_gp];
            const uint32_t acc = (n - 1) & 1, accphase = ((n - 1) >> 1) & 1;
            mbar_wait(&bars->tmem_full[acc], accphase);
            tc_fence_after();
            const float chh = coff[((n - 1) & (kDtSlots - 1)) * kDtM + m];
            const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(((warp - 8) & 3) * 32) << 16);
            const int c_begin = warp < 12 ? 5 : 9, c_end = warp < 12 ? 9 : kDtBins / 8;
            fd_unit_epilogue(p, it, u_last % p.n_gp, m, lane, taddr, chh, smem_u32(tab), 4u * (unsigned)p.n_store_units,
                             c_begin, c_end);
            tc_fence_before();
            if (!it.phase) __threadfence();
            asm volatile("bar.sync 3, 384;" ::: "memory");
        }
    } else if (warp >= 4) {
        // ===================== epilogue: rotation, three-tap hann, two emissions per half block =====================
        const int q = warp - 4;                  // TMEM lane quadrant
        const int m = threadIdx.x - 128;         // accumulator lane = channel of the unit
        uint32_t n = 0;
        const unsigned store_target = 4u * (unsigned)p.n_store_units;
        const uint32_t tab_s = smem_u32(tab);
        for (int q = q0; q < n_rounds; q += q_step) {
            const int u = unit_of(q);
            if (u >= n_units) break;
            const DtItem it = p.items[u / p.n_gp];
            const bool last = is_last(q);                            // the converter warps take chunks 5 - 12 of the last unit
            const uint32_t acc = n & 1, accphase = (n >> 1) & 1;
            mbar_wait(&bars->tmem_full[acc], accphase);
            if (m == 0 && q == q0) dt_stamp(p, 50);
            tc_fence_after();
            const float chh = coff[(n & (kDtSlots - 1)) * kDtM + m];
            const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
            fd_unit_epilogue(p, it, u % p.n_gp, m, lane, taddr, chh, tab_s, store_target, 0, last ? 5 : kDtBins / 8);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
            if (m == 0 && q == q0) dt_stamp(p, 51);
            if (!it.phase) __threadfence();      // this thread's stores before the warp's arrival on the store counter
            if (last) asm volatile("bar.sync 3, 384;" ::: "memory");   // ... and the helpers' stores (they fence too)
            if (!it.phase) {
                __syncwarp();
                if (lane == 0) atomicAdd(&p.ctr->next, 1u);
            }
            ++n;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (MC) dt_cluster_sync();                   // no multicast into, and no commit onto, a CTA that has left
    if (warp == 2) tmem_dealloc(tmem_base, 512);
    if (threadIdx.x == 0) dt_stamp(p, 52);
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&p.ctr->done, 1u) == gridDim.x - 1u) {
            p.ctr->next = 0u;
            p.ctr->done = 0u;
            __threadfence();
        }
    }
}

// Folded table: rows [plane][set][112][N/4]: set 0 = cos(theta_b (k + 1/2)), set 1 = sin(theta_b (k + 1/2)), b = b0 + j
// for j < 104 (zero rows above); plane 0 = BF16 hi, plane 1 = lo.
__global__ void dft_w_fold_table_kernel(__nv_bfloat16* W, int Kw, int N, int b0) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;                    // set * 112 + j
    if (k >= Kw) return;
    const int set = r / kFdRows, j = r - set * kFdRows;
    double v = 0.0;
    if (j < kDtBins) {
        const long long q = ((long long)(b0 + j) * (2 * k + 1)) % (2LL * N);     // angle = pi q / N
        double s, c;
        sincospi((double)q / (double)N, &s, &c);
        v = set ? s : c;
    }
    const __nv_bfloat16 hi = __double2bfloat16(v);
    const __nv_bfloat16 lo = __double2bfloat16(v - (double)__bfloat162float(hi));
    W[(long long)r * Kw + k] = hi;
    W[(long long)(2 * kFdRows + r) * Kw + k] = lo;
}

// W[r][n], r = 2 j + part: part 0 = cos, part 1 = -sin of 2 pi (b0 + j) n / N; rows [0, 208) hi, [208, 416) lo
__global__ void dft_w_table_kernel(__nv_bfloat16* W, int Kw, int N, int b0) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (n >= Kw) return;
    const int b = b0 + (r >> 1);
    const long long q = ((long long)b * n) % N;
    double s, c;
    sincospi(2.0 * (double)q / (double)N, &s, &c);
    const double v = (r & 1) ? -s : c;
    const __nv_bfloat16 hi = __double2bfloat16(v);
    const __nv_bfloat16 lo = __double2bfloat16(v - (double)__bfloat162float(hi));
    W[(long long)r * Kw + n] = hi;
    W[(long long)(kDtCols + r) * Kw + n] = lo;
}

struct WelchHannPlan {
    int dev, N, bin_lo, F, b0, KB, n_seg, n_items, n_store;
    long long max_row_end;
    DtItem* d_items;
    float* d_e1im;
    __nv_bfloat16* d_W;
    CUtensorMap mW;          // box = 208 rows: one CTA stages a whole W k-block
    CUtensorMap mW_half;     // box = 104 rows: each CTA of a pair stages half of it
    // folded kernel
    __nv_bfloat16* d_Wf;     // [2 planes][2 sets][112][N / 4]
    float* d_e1c;            // [104]
    float* d_rot;            // [104][2]
    CUtensorMap mWf;         // box = 32 x 224 rows (cos rows + sin rows of one plane), SWIZZLE_64B
};

static int make_raw_map(CUtensorMap* m, const float* x, int64_t n_samples, int n_ch, int64_t ld, int box_rows = kDtKB) {
    EncodeTiledFn enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    cuuint64_t dims[2] = {(cuuint64_t)n_ch, (cuuint64_t)n_samples};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(recording, tensor-core DFT) failed with CUresult %d", (int)r);
        return CMC_ECUDA;
    }
    return CMC_OK;
}

// W planes [2 * 208][Kw] BF16, K-major: box = one k-block (kDtKB elements) x 208 rows, swizzle = row bytes
static int make_w_map(CUtensorMap* m, const __nv_bfloat16* W, int Kw, int box_rows = kDtCols) {
    EncodeTiledFn enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    cuuint64_t dims[2] = {(cuuint64_t)Kw, (cuuint64_t)(2 * kDtCols)};
    cuuint64_t strides[1] = {(cuuint64_t)Kw * 2};
    cuuint32_t box[2] = {(cuuint32_t)kDtKB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(W), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, kDtKB == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(W table) failed with CUresult %d", (int)r);
        return CMC_ECUDA;
    }
    return CMC_OK;
}

// folded table [2 * 224][N / 4] BF16: box = 32 folded samples x 224 rows (cos rows + sin rows of one plane)
static int make_wf_map(CUtensorMap* m, const __nv_bfloat16* W, int Kw) {
    EncodeTiledFn enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    cuuint64_t dims[2] = {(cuuint64_t)Kw, (cuuint64_t)(4 * kFdRows)};
    cuuint64_t strides[1] = {(cuuint64_t)Kw * 2};
    cuuint32_t box[2] = {(cuuint32_t)kFdKB, (cuuint32_t)(2 * kFdRows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(W), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(folded W table) failed with CUresult %d", (int)r);
        return CMC_ECUDA;
    }
    return CMC_OK;
}

}  // namespace cmc

using namespace cmc;

extern "C" int cmc_welch_hann_plan_create(const int64_t* seg_starts_host, int n_seg, int N, int bin_lo, int bin_hi,
                                          void** plan_out) {
    CMC_REQUIRE(seg_starts_host && plan_out, "cmc_welch_hann_plan_create: null pointer");
    *plan_out = nullptr;
    CMC_REQUIRE(n_seg >= 1, "cmc_welch_hann_plan_create: no segments");
    CMC_REQUIRE(bin_lo >= 0 && bin_hi >= bin_lo && bin_hi <= N / 2,
                "cmc_welch_hann_plan_create: bins [%d, %d] outside [0, %d]", bin_lo, bin_hi, N / 2);
    const int F = bin_hi - bin_lo + 1;
    const int b0 = bin_lo > 0 ? bin_lo - 1 : 0;
    // accumulator bins b0 .. b0 + 103 must cover bin_lo - 1 .. bin_hi + 1 and stay below the Nyquist bin
    if (N < 256 || N > 16384 || (N % 128) != 0 || bin_hi + 1 - b0 > kDtBins - 1 || b0 + kDtBins > N / 2) {
        set_error("cmc_welch_hann_plan_create: N=%d bins [%d, %d] outside the tensor-core kernel (N %% 128 == 0, "
                  "256 <= N <= 16384, at most %d bins, band below Nyquist)", N, bin_lo, bin_hi, kDtBins - 2);
        return CMC_EUNSUPPORTED;
    }
    for (int s = 0; s < n_seg; ++s)
        CMC_REQUIRE(seg_starts_host[s] >= 0 && seg_starts_host[s] + N < (1ll << 31),
                    "cmc_welch_hann_plan_create: segment start %lld out of range", (long long)seg_starts_host[s]);
    // chains of segments that overlap by exactly N / 2 share their half blocks
    std::vector<int> order(n_seg);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return seg_starts_host[a] < seg_starts_host[b]; });
    std::vector<DtItem> items;
    long long max_end = 0;
    for (int i = 0; i < n_seg;) {
        int j = i;
        while (j + 1 < n_seg && j + 1 - i < kDtChainCap &&
               seg_starts_host[order[j + 1]] == seg_starts_host[order[j]] + N / 2)
            ++j;
        const int len = j - i + 1;
        const long long first = seg_starts_host[order[i]];
        for (int h = 0; h <= len; ++h) {
            DtItem it{};
            it.x_row = (int)(first + (long long)h * (N / 2));
            it.c_row = (int)first;
            it.seg_a = h < len ? order[i + h] : -1;
            it.seg_b = h >= 1 ? order[i + h - 1] : -1;
            it.phase = h & 1;
            items.push_back(it);
        }
        max_end = std::max(max_end, first + (long long)(len + 1) * (N / 2));
        i = j + 1;
    }
    std::stable_sort(items.begin(), items.end(), [](const DtItem& a, const DtItem& b) { return a.phase < b.phase; });
    auto* pl = new WelchHannPlan();
    pl->N = N; pl->bin_lo = bin_lo; pl->F = F; pl->b0 = b0; pl->KB = N / 2 / kDtKB; pl->n_seg = n_seg;
    pl->n_items = (int)items.size();
    pl->n_store = (int)std::count_if(items.begin(), items.end(), [](const DtItem& a) { return a.phase == 0; });
    pl->max_row_end = max_end;
    pl->d_items = nullptr; pl->d_W = nullptr; pl->d_e1im = nullptr;
    pl->d_Wf = nullptr; pl->d_e1c = nullptr; pl->d_rot = nullptr;
    // folded kernel: E1c[b] = sum_{k < N/4} 2 cos(theta_b (k + 1/2)) = sin(pi b / 2) / sin(pi b / N) (N / 2 for b = 0),
    // rotation (cos phi_b, sin phi_b) with phi_b = theta_b (N/4 - 1/2) = pi b / 2 - pi b / N
    std::vector<float> e1c(kDtBins, 0.f), rot(2 * kDtBins, 0.f);
    for (int j = 0; j < kDtBins; ++j) {
        const int b = b0 + j;
        const double pi = 3.14159265358979323846;
        const double quarter[4] = {0.0, 1.0, 0.0, -1.0};     // sin(pi b / 2), exact
        e1c[j] = b == 0 ? (float)(N / 2) : (float)(quarter[b & 3] / sin(pi * (double)b / (double)N));
        const double phi = pi * (double)(b & 3) / 2.0 - pi * (double)b / (double)N;   // pi b / 2 reduced mod 2 pi
        rot[2 * j] = (float)cos(phi);
        rot[2 * j + 1] = (float)sin(phi);
    }
    // E1[b] = sum_{n < N/2} exp(-2 pi i b n / N) = 1 - i cot(pi b / N) for odd b (0 for even b != 0)
    std::vector<float> e1(kDtBins, 0.f);
    for (int j = 0; j < kDtBins; ++j) {
        const int b = b0 + j;
        if (b & 1) e1[j] = (float)(-1.0 / tan(3.14159265358979323846 * (double)b / (double)N));
    }
    int rc = check_cuda(cudaGetDevice(&pl->dev), "cudaGetDevice");
    const int Kw = N / 2;
    if (!rc) rc = check_cuda(cudaMalloc(&pl->d_items, items.size() * sizeof(DtItem)), "cudaMalloc(plan items)");
    if (!rc) rc = check_cuda(cudaMemcpy(pl->d_items, items.data(), items.size() * sizeof(DtItem), cudaMemcpyHostToDevice),
                             "cudaMemcpy(plan items)");
    if (!rc) rc = check_cuda(cudaMalloc(&pl->d_e1im, kDtBins * sizeof(float)), "cudaMalloc(plan E1)");
    if (!rc) rc = check_cuda(cudaMemcpy(pl->d_e1im, e1.data(), kDtBins * sizeof(float), cudaMemcpyHostToDevice),
                             "cudaMemcpy(plan E1)");
    if (!rc) rc = check_cuda(cudaMalloc(&pl->d_W, (size_t)2 * kDtCols * Kw * sizeof(__nv_bfloat16)), "cudaMalloc(plan W)");
    if (!rc) {
        dft_w_table_kernel<<<dim3((Kw + 127) / 128, kDtCols), 128>>>(pl->d_W, Kw, N, b0);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        rc = check_cuda(cudaGetLastError(), "dft_w_table_kernel");
    }
    if (!rc) rc = check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize(plan)");
    if (!rc) rc = make_w_map(&pl->mW, pl->d_W, Kw);
    if (!rc) rc = make_w_map(&pl->mW_half, pl->d_W, Kw, kDtCols / 2);
    const int Kf = N / 4;
    if (!rc) rc = check_cuda(cudaMalloc(&pl->d_e1c, kDtBins * sizeof(float)), "cudaMalloc(plan E1c)");
    if (!rc) rc = check_cuda(cudaMemcpy(pl->d_e1c, e1c.data(), kDtBins * sizeof(float), cudaMemcpyHostToDevice), "cudaMemcpy(plan E1c)");
    if (!rc) rc = check_cuda(cudaMalloc(&pl->d_rot, 2 * kDtBins * sizeof(float)), "cudaMalloc(plan rot)");
    if (!rc) rc = check_cuda(cudaMemcpy(pl->d_rot, rot.data(), 2 * kDtBins * sizeof(float), cudaMemcpyHostToDevice), "cudaMemcpy(plan rot)");
    if (!rc) rc = check_cuda(cudaMalloc(&pl->d_Wf, (size_t)4 * kFdRows * Kf * sizeof(__nv_bfloat16)), "cudaMalloc(plan folded W)");
    if (!rc) {
        dft_w_fold_table_kernel<<<dim3((Kf + 127) / 128, 2 * kFdRows), 128>>>(pl->d_Wf, Kf, N, b0);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        rc = check_cuda(cudaGetLastError(), "dft_w_fold_table_kernel");
    }
    if (!rc) rc = check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize(plan, folded table)");
    if (!rc) rc = make_wf_map(&pl->mWf, pl->d_Wf, Kf);
    if (rc) {
        cudaFree(pl->d_items);
        cudaFree(pl->d_e1im);
        cudaFree(pl->d_W);
        cudaFree(pl->d_Wf);
        cudaFree(pl->d_e1c);
        cudaFree(pl->d_rot);
        delete pl;
        return rc;
    }
    *plan_out = pl;
    return CMC_OK;
}

extern "C" CMC_API int cmc_dbg_dt_trace(unsigned long long* out) {
    return (int)cudaMemcpyFromSymbol(out, cmc::g_dt_trace, sizeof(unsigned long long) * 64);
}

extern "C" int cmc_welch_hann_plan_destroy(void* plan) {
    if (!plan) return CMC_OK;
    auto* pl = static_cast<WelchHannPlan*>(plan);
    cudaFree(pl->d_items);
    cudaFree(pl->d_e1im);
    cudaFree(pl->d_W);
    cudaFree(pl->d_Wf);
    cudaFree(pl->d_e1c);
    cudaFree(pl->d_rot);
    delete pl;
    return CMC_OK;
}

extern "C" int cmc_welch_hann_plan_info(const void* plan, int* n_half_blocks, int* n_segments, int* n_bins) {
    CMC_REQUIRE(plan, "cmc_welch_hann_plan_info: null plan");
    const auto* pl = static_cast<const WelchHannPlan*>(plan);
    if (n_half_blocks) *n_half_blocks = pl->n_items;
    if (n_segments) *n_segments = pl->n_seg;
    if (n_bins) *n_bins = pl->F;
    return CMC_OK;
}

extern "C" int cmc_welch_hann_spectra(const void* plan, const float* x1, int n_ch1, int64_t ld1, float* spec1,
                                      const float* x2, int n_ch2, int64_t ld2, float* spec2, int64_t n_samples,
                                      int detrend, int64_t spec_ld, void* stream) {
    CMC_REQUIRE(plan && x1 && spec1, "cmc_welch_hann_spectra: null pointer");
    const auto* pl = static_cast<const WelchHannPlan*>(plan);
    if (!x2) n_ch2 = 0;
    CMC_REQUIRE(n_ch1 >= 1 && ld1 >= n_ch1 && n_ch2 >= 0 && (n_ch2 == 0 || (spec2 && ld2 >= n_ch2)) &&
                spec_ld >= n_ch1 && spec_ld >= n_ch2, "cmc_welch_hann_spectra: bad channel count / pitch");
    CMC_REQUIRE(detrend >= 0 && detrend <= 2, "cmc_welch_hann_spectra: detrend must be 0, 1 or 2");
    CMC_REQUIRE(pl->max_row_end <= n_samples, "cmc_welch_hann_spectra: segment outside the recording (%lld > %lld)",
                pl->max_row_end, (long long)n_samples);
    auto tma_ok = [](const float* x, int64_t ld) { return (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0; };
    if (!tma_ok(x1, ld1) || (n_ch2 && !tma_ok(x2, ld2)) || (reinterpret_cast<uintptr_t>(spec1) & 7) ||
        (n_ch2 && (reinterpret_cast<uintptr_t>(spec2) & 7))) {
        set_error("cmc_welch_hann_spectra: recordings need 16-byte aligned rows (channel pitch %% 4 == 0), spectra 8-byte alignment");
        return CMC_EUNSUPPORTED;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    CMC_REQUIRE(dev == pl->dev, "cmc_welch_hann_spectra: plan built on device %d, current device %d", pl->dev, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TileCounter* ctr = tile_counter_for(dev, st);
    if (!ctr) {
        set_error("cmc_welch_hann_spectra: no launch counter available (too many captured launches)");
        return CMC_EUNSUPPORTED;
    }
    // default: the folded kernel; CMC_DT_UNFOLD=1 / CMC_DT_PAIR=1 select the unfolded forms
    const bool fold = getenv("CMC_DT_UNFOLD") == nullptr && getenv("CMC_DT_PAIR") == nullptr;
    CUtensorMap m0, m1;
    int rc = make_raw_map(&m0, x1, n_samples, n_ch1, ld1, fold ? kFdKB : kDtKB);
    if (rc) return rc;
    if (n_ch2) { if ((rc = make_raw_map(&m1, x2, n_samples, n_ch2, ld2, fold ? kFdKB : kDtKB))) return rc; }
    else m1 = m0;
    DtParams p{};
    p.items = pl->d_items;
    p.e1im = pl->d_e1im;
    p.n_items = pl->n_items;
    p.n_grp0 = (n_ch1 + 63) / 64;
    const int n_grp = p.n_grp0 + (n_ch2 + 63) / 64;
    p.n_gp = (n_grp + 1) / 2;
    p.n_store_units = pl->n_store * p.n_gp;
    p.KB = pl->KB; p.N = pl->N; p.bin_lo = pl->bin_lo; p.F = pl->F; p.b0 = pl->b0; p.detrend = detrend;
    p.x[0] = x1; p.x[1] = n_ch2 ? x2 : x1;
    p.ld[0] = ld1; p.ld[1] = n_ch2 ? ld2 : ld1;
    p.n_ch[0] = n_ch1; p.n_ch[1] = n_ch2;
    p.spec[0] = reinterpret_cast<float2*>(spec1);
    p.spec[1] = reinterpret_cast<float2*>(n_ch2 ? spec2 : spec1);
    p.spec_ld = spec_ld;
    p.ctr = ctr;
    { const char* e = getenv("CMC_DT_DBG"); p.dbg = e ? atoi(e) : 0; }
    { const char* e = getenv("CMC_DT_PF"); p.pf = e ? atoi(e) : kDtPrefetch; }
    const long long n_units = (long long)p.n_items * p.n_gp;
    if (fold) {
        p.KB = pl->N / 4 / kFdKB;
        p.e1im = pl->d_e1c;
        p.rot = pl->d_rot;
        const size_t smem_f = 1024 + (size_t)(kFdStagesL + kFdStagesA) * kFdABytes + (size_t)kFdStagesB * kFdBBytes +
                              (2 + kDtSlots) * kDtM * sizeof(float) + kFdRows * sizeof(float4) + sizeof(FdBarriers) + 16;
        // CMC_DT_MC=1: clusters of two CTAs share every table k-block by TMA multicast
        if (getenv("CMC_DT_MC") != nullptr && sms >= 2) {
            auto kern = dft_hann_fold_kernel<true>;
            rc = ensure_smem_attr(reinterpret_cast<const void*>(kern), smem_f);
            if (rc) return rc;
            const long long n_rounds = (n_units + 1) / 2;
            const long long pairs = n_rounds < sms / 2 ? n_rounds : sms / 2;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)(2 * pairs));
            cfg.blockDim = dim3(kDtThreads);
            cfg.dynamicSmemBytes = smem_f;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            rc = check_cuda(cudaLaunchKernelEx(&cfg, kern, m0, m1, pl->mWf, p), "cudaLaunchKernelEx(dft_hann_fold_kernel<mc>)");
            if (rc) return rc;
            g_launches.fetch_add(1, std::memory_order_relaxed);
            return CMC_OK;
        }
        rc = ensure_smem_attr(reinterpret_cast<const void*>(dft_hann_fold_kernel<false>), smem_f);
        if (rc) return rc;
        const unsigned grid_f = (unsigned)(n_units < sms ? n_units : sms);
        dft_hann_fold_kernel<false><<<grid_f, kDtThreads, smem_f, st>>>(m0, m1, pl->mWf, p);
        CMC_CHECK_LAUNCH("dft_hann_fold_kernel");
        return CMC_OK;
    }
    // CMC_DT_PAIR=1: CTA pairs (tcgen05 cta_group::2) share every W k-block.  Bit-identical results; measured SLOWER
    // (77 us against 54 us for config 2: the pair's barrier round trips lengthen every stage cycle by ~2 us while the
    // raw / A ring, which bounds the loop, is no deeper), so it is not the default.
    const bool pair = getenv("CMC_DT_PAIR") != nullptr && kDtKB == 64 && sms >= 2;
    const size_t smem = 1024 + (size_t)kDtStagesA * kDtABytes +
                        (pair ? (size_t)kDtStagesBPair * (kDtBBytes / 2) : (size_t)kDtStagesB * kDtBBytes) +
                        (2 + kDtSlots) * kDtM * sizeof(float) + 112 * sizeof(float2) + sizeof(DtBarriers) + 16;
    if (pair) {
        auto kern = dft_hann_tc_kernel<true>;
        rc = ensure_smem_attr(reinterpret_cast<const void*>(kern), smem);
        if (rc) return rc;
        const long long n_rounds = (n_units + 1) / 2;
        const long long pairs = n_rounds < sms / 2 ? n_rounds : sms / 2;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(2 * pairs));
        cfg.blockDim = dim3(kDtThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        rc = check_cuda(cudaLaunchKernelEx(&cfg, kern, m0, m1, pl->mW_half, p), "cudaLaunchKernelEx(dft_hann_tc_kernel<pair>)");
        if (rc) return rc;
    } else {
        auto kern = dft_hann_tc_kernel<false>;
        rc = ensure_smem_attr(reinterpret_cast<const void*>(kern), smem);
        if (rc) return rc;
        const unsigned grid = (unsigned)(n_units < sms ? n_units : sms);
        kern<<<grid, kDtThreads, smem, st>>>(m0, m1, pl->mW, p);
    }
    CMC_CHECK_LAUNCH("dft_hann_tc_kernel");
    return CMC_OK;
}

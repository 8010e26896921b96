// K3 (shift mode), batched: four circular-shift surrogates per tensor-core tile, error-compensated 3xTF32.
//
// One tile = (frequency, 64 x 64 channel tile, group of four distinct shifts).  The A operand ([X; i X], 128 rows,
// K-major, TF32 hi and lo planes) is staged ONCE per k-block and contracted against four views of the doubled B rows
// (B_dbl / B_odd and their lo-plane twins, csd_layout.cuh) read at four K offsets - TMA boxes that land as one
// 256-row B tile per plane, three tcgen05.mma of N = 256 per k-step (lo*hi + hi*lo + hi*hi, small terms first): the
// surrogate coherences carry the same ~2e-6 error as the observed pass, so exceedance counts are compared against
// the fp64 definition inside the +-1e-4 band of the north star instead of the 4e-4 .. 2e-3 a single TF32 term gave.
// Stage = {A_hi, A_lo, 4 x B_hi, 4 x B_lo} = 96 KB, two stages.  Epilogue per shift: coherence from the 64-column
// slice, exceedance counts weighted by the shift's multiplicity, per-shift running maximum; the observed coherence
// and the counters of a thread's 32 fixed (i, j) positions live in REGISTERS for all shifts of a channel tile
// (one-shift-per-tile with global loads per output: 1.00 ms for config 3; shared-memory tables: 0.44 ms).
#include "common.cuh"
#include "tc_common.cuh"
#include "csd_layout.cuh"

namespace cmc {

using namespace tc;

constexpr int kShQ = 4;                                   // shifts per tile
constexpr int kShStages = 2;
constexpr int kShABytes = kTileM * kKBlock * 4;           // 16 KB per plane
constexpr int kShBBytes = kTileN * kKBlock * 4;           // 8 KB per shift and plane
constexpr int kShBOff = 2 * kShABytes;                    // B_hi views
constexpr int kShBLoOff = kShBOff + kShQ * kShBBytes;     // B_lo views
constexpr int kShStageBytes = 2 * kShABytes + 2 * kShQ * kShBBytes;   // 96 KB
constexpr int kShPitch = kTileN + 1;
constexpr int kShThreads = 256;

struct ShiftParams {
    int F, f0, MT, NT, Ne, Nm, KB, n_pos, n_groups;
    const int32_t* shift_off;     // [n_pos] K offset (floats) into the doubled B rows
    const uint32_t* shift_mult;   // [n_pos] surrogates that use this shift (0 = unused)
    const float* pxx;             // [F][Ne]
    const float* pyy;             // [F][Nm]
    const float* coh_obs;         // [F][Ne][Nm]
    uint32_t* exceed;             // [F][Ne][Nm]
    uint32_t* max_u;              // [n_pos] float bits
    long long total_tiles;        // F * MT * NT * n_groups
};

struct ShTile {
    int f, mt, nt, g;
};
// Work distribution.  The tiles of one (frequency, channel tile) are cut into chunks of kShChunk consecutive shift
// groups and the chunks are dealt to the CTAs in turn, frequency-major: at any moment the 148 CTAs then work on
// ~10 neighbouring frequencies, whose operands (1.4 MB each) stay in L2.  A contiguous split of the whole tile range
// - one frequency per CTA - kept every frequency's operands live at once: 138 MB against 126 MB of L2, 1.63 GB of
// DRAM reads per null (ncu, profiles/r02_ncu_full_raw.csv).  Within a chunk the channel tile is fixed, so the
// epilogue's per-tile registers (observed coherence, counters) are loaded / flushed once per chunk.
constexpr int kShChunk = 4;
template <class Fn>
__device__ __forceinline__ void sh_for_each_tile(const ShiftParams& p, Fn fn) {
    const int n_gc = (p.n_groups + kShChunk - 1) / kShChunk;
    const long long n_chunks = (long long)p.F * p.MT * p.NT * n_gc;
    for (long long ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
        ShTile c;
        const int gc = (int)(ci % n_gc);
        long long r = ci / n_gc;
        c.nt = (int)(r % p.NT);
        r /= p.NT;
        c.mt = (int)(r % p.MT);
        c.f = p.f0 + (int)(r / p.MT);
        const int g_end = min(p.n_groups, (gc + 1) * kShChunk);
        for (c.g = gc * kShChunk; c.g < g_end; ++c.g) fn(c);
    }
}
// multiplicities of the four shifts of a group (0 past the last position); all zero = nothing to do
__device__ __forceinline__ bool sh_mults(const ShiftParams& p, int g, uint32_t (&m)[kShQ]) {
    uint32_t any = 0;
#pragma unroll
    for (int q = 0; q < kShQ; ++q) {
        const int sh = g * kShQ + q;
        m[q] = sh < p.n_pos ? p.shift_mult[sh] : 0u;
        any |= m[q];
    }
    return any != 0;
}

struct __align__(8) ShBarriers {
    uint64_t full[kShStages];
    uint64_t empty[kShStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__global__ void __launch_bounds__(kShThreads, 1)
csd_shift4_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mAlo,
                  const __grid_constant__ CUtensorMap mBdbl, const __grid_constant__ CUtensorMap mBodd,
                  const __grid_constant__ CUtensorMap mBdblLo, const __grid_constant__ CUtensorMap mBoddLo,
                  const ShiftParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* sS = base;                                                         // [kShStages][96 KB]
    float* stage_tile = reinterpret_cast<float*>(sS + kShStages * kShStageBytes);     // [128][65]
    float* scale = stage_tile + kTileM * kShPitch;                                    // [128] 1/sqrt(Pxx), 1/sqrt(Pyy)
    ShBarriers* bars = reinterpret_cast<ShBarriers*>(scale + 128);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kShStages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->tmem_full[a], 1);
            mbar_init(&bars->tmem_empty[a], 4);
        }
        fence_barrier_init();
        tma_prefetch_desc(&mA);
        tma_prefetch_desc(&mAlo);
        tma_prefetch_desc(&mBdbl);
        tma_prefetch_desc(&mBodd);
        tma_prefetch_desc(&mBdblLo);
        tma_prefetch_desc(&mBoddLo);
    }
    if (warp == 2) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer: lane 0 owns the barriers, lanes 0-9 issue one box each =====================
        // lanes 0 / 1: A_hi / A_lo; lanes 2-5: B_hi view of shift q = lane - 2; lanes 6-9: B_lo view of shift q = lane - 6
        int stage = 0;
        uint32_t phase = 0;
        sh_for_each_tile(p, [&](const ShTile& c) {
            uint32_t mult[kShQ];
            if (!sh_mults(p, c.g, mult)) return;
            // a TMA box must start 16-byte aligned: offsets = 2 (mod 4) floats read the copy of the B rows that is
            // pre-shifted by one complex element
            int off = 0;
            bool odd = false;
            const int qb = lane >= 2 ? (lane - 2) & 3 : 0;
            if (lane >= 2 && lane < 2 + 2 * kShQ) {
                const int sh = c.g * kShQ + qb;
                off = sh < p.n_pos ? p.shift_off[sh] : 0;
                odd = (off & 2) != 0;
                off -= odd ? 2 : 0;
            }
            const bool lo_plane = lane >= 2 + kShQ;
            const CUtensorMap* bmap = lo_plane ? (odd ? &mBoddLo : &mBdblLo) : (odd ? &mBodd : &mBdbl);
            const int arow = (c.f * p.MT + c.mt) * kTileM;
            const int brow = (c.f * p.NT + c.nt) * kTileN;
            for (int kb = 0; kb < p.KB; ++kb) {
                unsigned char* st = sS + stage * kShStageBytes;
                if (lane == 0) {
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars->full[stage], kShStageBytes);
                }
                __syncwarp();
                if (lane < 2)
                    tma_load_2d(st + lane * kShABytes, lane ? &mAlo : &mA, &bars->full[stage], kb * kKBlock, arow);
                else if (lane < 2 + 2 * kShQ)
                    tma_load_2d(st + (lo_plane ? kShBLoOff : kShBOff) + qb * kShBBytes, bmap, &bars->full[stage],
                                off + kb * kKBlock, brow);
                if (++stage == kShStages) { stage = 0; phase ^= 1; }
            }
        });
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(kTileM, kShQ * kTileN);
            int stage = 0;
            uint32_t phase = 0, it = 0;
            sh_for_each_tile(p, [&](const ShTile& c) {
                uint32_t mult[kShQ];
                if (!sh_mults(p, c.g, mult)) return;
                const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
                mbar_wait(&bars->tmem_empty[acc], accphase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * (kShQ * kTileN);
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint32_t s0 = smem_u32(sS + stage * kShStageBytes);
#pragma unroll
                    for (int k = 0; k < kKBlock / 8; ++k) {
                        const uint64_t dah = make_smem_desc_k_sw128(s0 + k * 32);
                        const uint64_t dal = make_smem_desc_k_sw128(s0 + kShABytes + k * 32);
                        const uint64_t dbh = make_smem_desc_k_sw128(s0 + kShBOff + k * 32);
                        const uint64_t dbl = make_smem_desc_k_sw128(s0 + kShBLoOff + k * 32);
                        umma_tf32(d, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);      // small terms first
                        umma_tf32(d, dah, dbl, idesc, 1u);
                        umma_tf32(d, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&bars->empty[stage]);
                    if (++stage == kShStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars->tmem_full[acc]);
                ++it;
            });
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int q4 = warp - 4;                 // TMEM lane quadrant of this warp
        const int te = threadIdx.x - 128;        // 0..127
        // thread te owns the pairs idx = te + 128 n (il = idx >> 6, jl = idx & 63), n = 0..31, of every channel tile
        float obs[32];
        uint32_t cnt[32];
#pragma unroll
        for (int n = 0; n < 32; ++n) cnt[n] = 0;
        long long key = -1;
        int kf = 0, kmt = 0, knt = 0;
        uint32_t it = 0;
        auto flush = [&]() {
#pragma unroll
            for (int n = 0; n < 32; ++n) {
                const int idx = te + 128 * n;
                if (cnt[n])         // out-of-range pairs never count (their threshold is unreachable)
                    atomicAdd(&p.exceed[((long long)kf * p.Ne + kmt * 64 + (idx >> 6)) * p.Nm + knt * 64 + (idx & 63)], cnt[n]);
                cnt[n] = 0;
            }
        };
        sh_for_each_tile(p, [&](const ShTile& c) {
            uint32_t mult[kShQ];
            if (!sh_mults(p, c.g, mult)) return;
            const long long k2 = ((long long)c.f * p.MT + c.mt) * p.NT + c.nt;
            if (k2 != key) {
                if (key >= 0) flush();
                key = k2; kf = c.f; kmt = c.mt; knt = c.nt;
                // per channel tile: normalisation factors and the observed coherence, once for all its shifts
                // (out-of-range pairs get an unreachable threshold, so they never count and never raise the maximum)
                asm volatile("bar.sync 1, 128;" ::: "memory");   // the previous tile's readers of scale[] are done
                {
                    const bool is_x = te < 64;
                    const int ch = (is_x ? c.mt : c.nt) * 64 + (te & 63);
                    float pw = 0.f;
                    if (ch < (is_x ? p.Ne : p.Nm))
                        pw = __ldg((is_x ? p.pxx : p.pyy) + (long long)c.f * (is_x ? p.Ne : p.Nm) + ch);
                    scale[te] = pw > 0.f ? rsqrtf(pw) : 0.f;
                }
#pragma unroll
                for (int n = 0; n < 32; ++n) {
                    const int idx = te + 128 * n, i = c.mt * 64 + (idx >> 6), j = c.nt * 64 + (idx & 63);
                    obs[n] = (i < p.Ne && j < p.Nm) ? __ldg(p.coh_obs + ((long long)c.f * p.Ne + i) * p.Nm + j) : 2.0f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
            mbar_wait(&bars->tmem_full[acc], accphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * (kShQ * kTileN) + (static_cast<uint32_t>(q4 * 32) << 16);
#pragma unroll 1
            for (int q = 0; q < kShQ; ++q) {
                if (mult[q] == 0) continue;                       // CTA-uniform
                uint32_t r0[32], r1[32];
                tmem_ld_32x32(taddr + q * kTileN, r0);
                tmem_ld_32x32(taddr + q * kTileN + 32, r1);
                tmem_ld_wait();
                float* row = stage_tile + (q4 * 32 + lane) * kShPitch;
#pragma unroll
                for (int cidx = 0; cidx < 32; ++cidx) {
                    row[cidx] = __uint_as_float(r0[cidx]);
                    row[32 + cidx] = __uint_as_float(r1[cidx]);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                float vmax = 0.f;
                const uint32_t mq = mult[q];
#pragma unroll
                for (int n = 0; n < 32; ++n) {
                    const int idx = te + 128 * n;
                    const int il = idx >> 6, jl = idx & 63;
                    const float sc = scale[il] * scale[64 + jl];          // 0 for padding / silent channels
                    const float a = stage_tile[il * kShPitch + jl] * sc;
                    const float b = stage_tile[(64 + il) * kShPitch + jl] * sc;
                    const float cval = fminf(a * a + b * b, 1.0f);
                    cnt[n] += cval >= obs[n] ? mq : 0u;
                    vmax = fmaxf(vmax, cval);
                }
                const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(vmax));
                if (lane == 0) atomicMax(&p.max_u[c.g * kShQ + q], m);
                asm volatile("bar.sync 1, 128;" ::: "memory");     // stage_tile is rewritten by the next slice
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);   // accumulator may be overwritten
            ++it;
        });
        if (key >= 0) flush();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// launch on the operands of a filled pooled-CSD workspace (B_dbl / B_odd already built)
int launch_shift4(const CsdLayout& y, unsigned char* ws, int f_begin, int f_end, int Ne, int Nm, int n_pos,
                  const int32_t* shift_off, const uint32_t* shift_mult, const float* coh_obs, uint32_t* exceed,
                  uint32_t* max_u, cudaStream_t st) {
    CUtensorMap mA, mAlo, mBdbl, mBodd, mBdblLo, mBoddLo;
    int rc;
    const int64_t arows = (int64_t)y.F * y.MT * kTileM, brows = (int64_t)y.F * y.NT * kTileN;
    if ((rc = make_operand_map(&mA, reinterpret_cast<float*>(ws + y.off_ahi), y.KP, arows, kTileM))) return rc;
    if ((rc = make_operand_map(&mAlo, reinterpret_cast<float*>(ws + y.off_alo), y.KP, arows, kTileM))) return rc;
    if ((rc = make_operand_map(&mBdbl, reinterpret_cast<float*>(ws + y.off_bdbl), y.LB, brows, kTileN))) return rc;
    if ((rc = make_operand_map(&mBodd, reinterpret_cast<float*>(ws + y.off_bodd), y.LB, brows, kTileN))) return rc;
    if ((rc = make_operand_map(&mBdblLo, reinterpret_cast<float*>(ws + y.off_bdbl_lo), y.LB, brows, kTileN))) return rc;
    if ((rc = make_operand_map(&mBoddLo, reinterpret_cast<float*>(ws + y.off_bodd_lo), y.LB, brows, kTileN))) return rc;
    ShiftParams p{};
    p.F = f_end - f_begin; p.f0 = f_begin; p.MT = y.MT; p.NT = y.NT; p.Ne = Ne; p.Nm = Nm; p.KB = y.KP / kKBlock;
    p.n_pos = n_pos; p.n_groups = (n_pos + kShQ - 1) / kShQ;
    p.shift_off = shift_off; p.shift_mult = shift_mult;
    p.pxx = reinterpret_cast<const float*>(ws + y.off_pxx);
    p.pyy = reinterpret_cast<const float*>(ws + y.off_pyy);
    p.coh_obs = coh_obs; p.exceed = exceed; p.max_u = max_u;
    p.total_tiles = (long long)p.F * y.MT * y.NT * p.n_groups;
    const size_t smem = 1024 + (size_t)kShStages * kShStageBytes + sizeof(float) * kTileM * kShPitch + 128 * 4 +
                        sizeof(ShBarriers) + 16;
    rc = ensure_smem_attr(reinterpret_cast<const void*>(csd_shift4_kernel), smem);
    if (rc) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long grid = p.total_tiles < sms ? p.total_tiles : sms;
    csd_shift4_kernel<<<(unsigned)grid, kShThreads, smem, st>>>(mA, mAlo, mBdbl, mBodd, mBdblLo, mBoddLo, p);
    CMC_CHECK_LAUNCH("csd_shift4_kernel");
    return CMC_OK;
}

}  // namespace cmc

// K2 / K3: pooled cross-spectral density on the 5th-gen tensor cores (tcgen05.mma kind::tf32,
// TMA-staged K-major operands, FP32 accumulators in TMEM), fused with the auto-spectra
// normalisation into magnitude-squared coherence, and the shift-surrogate null that re-runs the
// same contraction on the cached operands with a rotated EMG segment index.
//
// Replaces signal_features.py:750-770 for averages over L segments / windows x tapers.
//
// Formulation.  The complex contraction is one real GEMM per frequency with K = (l, re/im) contiguous
// ("K-major" = complex64 memory order along the segment axis):
//     A rows 0..63   = X             (re, im interleaved along K)
//     A rows 64..127 = i * X         (-im, re)
//     B rows         = Y
//     D[i][j] = Re S_ij,  D[64 + i][j] = Im S_ij            (M = 128, N = 64, K = 2L)
// and the epilogue normalises with the auto-spectra: C = |S / sqrt(Pxx) / sqrt(Pyy)|^2 (TF32 is a floating
// format, so unscaled operands lose nothing).  One fused pack kernel transposes the spectra into the K-major
// operand rows, splits them into TF32 hi/lo planes and accumulates Pxx / Pyy on the way (one read of the spectra).
// The observed pass is error-compensated 3xTF32 (hi*hi + hi*lo + lo*hi with hi = tf32(x), lo = tf32(x - hi)):
// every pipeline stage carries the four tiles A_hi, A_lo, B_hi, B_lo of one k-block, so each operand byte is
// fetched once; |dC| ~ 2e-6 and the pass stays HBM bound.
// Shift surrogates (csd_shift.cu) read the B operand at K offset 2 * shift * group from a doubled row
// [Y | Y | 0...] - a TMA coordinate, no data movement - with the same three TF32 terms.
//
// Pipeline per CTA (persistent over a contiguous tile range): warp 0 = TMA producer, warp 1 = MMA issuer (one
// thread), warp 2 = TMEM allocator, warps 4-7 = epilogue.  mbarrier ring of smem stages (full/empty), 2 TMEM
// accumulators (tmem_full/tmem_empty) so the epilogue of tile t overlaps the MMAs of tile t + 1.
#include "common.cuh"
#include "tc_common.cuh"
#include "csd_layout.cuh"

namespace cmc {

using namespace tc;

constexpr int kStages = 3;                      // {A_hi, A_lo, B_hi, B_lo} = 48 KB each (a 4th stage measured no gain)
constexpr int kABytes = kTileM * kKBlock * 4;   // 16 KB
constexpr int kBBytes = kTileN * kKBlock * 4;   // 8 KB
constexpr int kStageBytes = 2 * (kABytes + kBBytes);
constexpr int kStagePitch = kTileN + 1;         // padded staging row (floats)
constexpr int kGemmThreads = 256;

struct CsdParams {
    int F, MT, NT, Ne, Nm, KB;
    float* coh;                   // out [F][Ne][Nm]
    float2* sxy;                  // optional out
    const float* pxx;             // [F][Ne] auto-spectra
    const float* pyy;             // [F][Nm]
    long long total_tiles;        // F * MT * NT
};

struct TileCoord {
    int f, mt, nt;
};
__device__ __forceinline__ TileCoord decode_tile(long long t, const CsdParams& p) {
    TileCoord c;
    c.nt = (int)(t % p.NT);
    const long long r = t / p.NT;
    c.mt = (int)(r % p.MT);
    c.f = (int)(r / p.MT);
    return c;
}

struct __align__(8) GemmBarriers {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
csd_gemm_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                const __grid_constant__ CUtensorMap mBhi, const __grid_constant__ CUtensorMap mBlo, const CsdParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* sS = base;                                   // [kStages][kStageBytes]
    float* stage_tile = reinterpret_cast<float*>(sS + kStages * kStageBytes);      // [128][65]
    GemmBarriers* bars = reinterpret_cast<GemmBarriers*>(stage_tile + kTileM * kStagePitch);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = p.total_tiles * blockIdx.x / gridDim.x;
    const long long t1 = p.total_tiles * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->tmem_full[a], 1);
            mbar_init(&bars->tmem_empty[a], 4);
        }
        fence_barrier_init();
        tma_prefetch_desc(&mAhi);
        tma_prefetch_desc(&mBhi);
    }
    if (warp == 2) {
        tmem_alloc(&bars->tmem_base, 2 * kTileN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = t0; t < t1; ++t) {
                const TileCoord c = decode_tile(t, p);
                const int arow = (c.f * p.MT + c.mt) * kTileM;
                const int brow = (c.f * p.NT + c.nt) * kTileN;
                for (int kb = 0; kb < p.KB; ++kb) {
                    unsigned char* st = sS + stage * kStageBytes;
                    mbar_wait(&bars->empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&bars->full[stage], kStageBytes);
                    tma_load_2d(st, &mAhi, &bars->full[stage], kb * kKBlock, arow);
                    tma_load_2d(st + kABytes, &mAlo, &bars->full[stage], kb * kKBlock, arow);
                    tma_load_2d(st + 2 * kABytes, &mBhi, &bars->full[stage], kb * kKBlock, brow);
                    tma_load_2d(st + 2 * kABytes + kBBytes, &mBlo, &bars->full[stage], kb * kKBlock, brow);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (single thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(kTileM, kTileN);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (long long t = t0; t < t1; ++t) {
                const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
                mbar_wait(&bars->tmem_empty[acc], accphase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * kTileN;
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint32_t s0 = smem_u32(sS + stage * kStageBytes);
                    const uint32_t ahi = s0, alo = s0 + kABytes, bhi = s0 + 2 * kABytes, blo = bhi + kBBytes;
#pragma unroll
                    for (int k = 0; k < kKBlock / 8; ++k) {
                        const uint64_t dah = make_smem_desc_k_sw128(ahi + k * 32), dal = make_smem_desc_k_sw128(alo + k * 32);
                        const uint64_t dbh = make_smem_desc_k_sw128(bhi + k * 32), dbl = make_smem_desc_k_sw128(blo + k * 32);
                        umma_tf32(d, dal, dbh, idesc, (kb | k) != 0 ? 1u : 0u);      // small terms first
                        umma_tf32(d, dah, dbl, idesc, 1u);
                        umma_tf32(d, dah, dbh, idesc, 1u);
                    }
                    umma_commit(&bars->empty[stage]);      // frees the smem slot when these MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bars->tmem_full[acc]);        // accumulator ready for the epilogue
                ++it;
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int q = warp - 4;                  // TMEM lane quadrant of this warp
        const int te = threadIdx.x - 128;        // 0..127
        uint32_t it = 0;
        for (long long t = t0; t < t1; ++t) {
            const TileCoord c = decode_tile(t, p);
            const uint32_t acc = it & 1, accphase = (it >> 1) & 1;
            mbar_wait(&bars->tmem_full[acc], accphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * kTileN + (static_cast<uint32_t>(q * 32) << 16);
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(taddr, r0);
            tmem_ld_32x32(taddr + 32, r1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);   // accumulator may be overwritten
            float* row = stage_tile + (q * 32 + lane) * kStagePitch;
#pragma unroll
            for (int cidx = 0; cidx < 32; ++cidx) {
                row[cidx] = __uint_as_float(r0[cidx]);
                row[32 + cidx] = __uint_as_float(r1[cidx]);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 4
            for (int n = 0; n < 32; ++n) {
                const int idx = te + 128 * n;
                const int il = idx >> 6, jl = idx & 63;
                const int i = c.mt * 64 + il, j = c.nt * 64 + jl;
                if (i < p.Ne && j < p.Nm) {
                    const float re = stage_tile[il * kStagePitch + jl];
                    const float im = stage_tile[(64 + il) * kStagePitch + jl];
                    const float px = __ldg(p.pxx + c.f * p.Ne + i), py = __ldg(p.pyy + c.f * p.Nm + j);
                    // |S|^2 / (Pxx Pyy) as |S / sqrt(Pxx) / sqrt(Pyy)|^2: no overflow, zero-power channels give 0
                    const float sc = (px > 0.f && py > 0.f) ? rsqrtf(px) * rsqrtf(py) : 0.f;
                    const float a = re * sc, b = im * sc;
                    const long long o = ((long long)c.f * p.Ne + i) * p.Nm + j;
                    p.coh[o] = fminf(a * a + b * b, 1.0f);
                    if (p.sxy) p.sxy[o] = make_float2(re, im);
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            ++it;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 2 * kTileN);
}

// ------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------
// Fused pack: transposes S[l][f][c] into K-major TF32 hi/lo operand rows (unscaled) and accumulates the
// auto-spectra P[f][c] = sum_l |S|^2 in a fixed order (deterministic) - one read of the spectra.
// grid (F, 2*MT + 2*NT): blockIdx.y < 2*MT -> 32 EEG channels (A rows: X and i*X), else 32 EMG channels (B rows).
// block (32, 8).  Columns k >= 2L are written as zeros.
__global__ void __launch_bounds__(256)
pack_fused_kernel(const float2* __restrict__ X, int64_t ldx, int Ne, const float2* __restrict__ Y, int64_t ldy, int Nm,
                  int L, int F, int MT, int NT, int KP, float* __restrict__ Ahi, float* __restrict__ Alo,
                  float* __restrict__ Bhi, float* __restrict__ Blo, float* __restrict__ pxx, float* __restrict__ pyy,
                  float* __restrict__ sxx_out, float* __restrict__ syy_out) {
    __shared__ float2 tile[32][33];       // [l][channel]
    __shared__ float part[8][33];
    const int f = blockIdx.x, tx = threadIdx.x, ty = threadIdx.y;
    const bool is_a = (int)blockIdx.y < 2 * MT;
    const int ct = is_a ? blockIdx.y : blockIdx.y - 2 * MT;      // 32-channel tile of this operand
    const float2* S = is_a ? X : Y;
    const int64_t ld = is_a ? ldx : ldy;
    const int C = is_a ? Ne : Nm;
    const int ch_ld = ct * 32 + tx;                               // channel loaded by this thread
    float pw = 0.f;
    for (int l0 = 0; l0 < KP / 2; l0 += 32) {
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int l = l0 + ty + 8 * n;
            float2 v = make_float2(0.f, 0.f);
            if (l < L && ch_ld < C) v = __ldg(S + ((int64_t)l * F + f) * ld + ch_ld);
            pw += v.x * v.x + v.y * v.y;
            tile[ty + 8 * n][tx] = v;
        }
        __syncthreads();
        const int k = 2 * (l0 + tx);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            if (k >= KP) break;                                   // KP / 2 need not be a multiple of 32
            const int cl = ty + 8 * n;                            // channel within the tile
            const int ch = ct * 32 + cl;
            const float2 v = tile[tx][cl];
            const float2 h = make_float2(to_tf32(v.x), to_tf32(v.y));
            const float2 lo = make_float2(to_tf32(v.x - h.x), to_tf32(v.y - h.y));
            if (is_a) {
                const int64_t row = ((int64_t)f * MT + (ch >> 6)) * kTileM + (ch & 63);
                const int64_t o = row * KP + k;
                *reinterpret_cast<float2*>(Ahi + o) = h;
                *reinterpret_cast<float2*>(Alo + o) = lo;
                const int64_t o2 = o + (int64_t)64 * KP;          // rows 64..127: i * X = (-im, re)
                *reinterpret_cast<float2*>(Ahi + o2) = make_float2(-h.y, h.x);
                *reinterpret_cast<float2*>(Alo + o2) = make_float2(-lo.y, lo.x);
            } else {
                const int64_t row = ((int64_t)f * NT + (ch >> 6)) * kTileN + (ch & 63);
                const int64_t o = row * KP + k;
                *reinterpret_cast<float2*>(Bhi + o) = h;
                *reinterpret_cast<float2*>(Blo + o) = lo;
            }
        }
        __syncthreads();
    }
    part[ty][tx] = pw;
    __syncthreads();
    if (ty == 0 && ch_ld < C) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += part[w][tx];
        if (is_a) {
            pxx[(int64_t)f * Ne + ch_ld] = t;
            if (sxx_out) sxx_out[(int64_t)f * Ne + ch_ld] = t;
        } else {
            pyy[(int64_t)f * Nm + ch_ld] = t;
            if (syy_out) syy_out[(int64_t)f * Nm + ch_ld] = t;
        }
    }
}

// Shift-surrogate views of the B operand: Bdbl[k] = B[k mod 2L] for k < 4L (else 0), Bodd[k] = Bdbl[k + 2], for
// both TF32 planes (blockIdx.z = 0: B_hi, 1: B_lo).  One thread per complex column; grid (rows, ceil(LB / 2 / 256), 2).
__global__ void __launch_bounds__(256)
shift_operand_kernel(const float* __restrict__ Bhi, const float* __restrict__ Blo, int L, int KP, int LB,
                     float* __restrict__ Bdbl, float* __restrict__ Bodd, float* __restrict__ BdblLo,
                     float* __restrict__ BoddLo) {
    const int lp = blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t row = blockIdx.x;                       // rows on the x axis: F * NT * 64 may exceed 65,535
    if (lp >= LB / 2) return;
    const bool lo = blockIdx.z != 0;
    const float2* src = reinterpret_cast<const float2*>((lo ? Blo : Bhi) + row * KP);
    const float2 z = make_float2(0.f, 0.f);
    const float2 a = lp < 2 * L ? src[lp % L] : z;
    const float2 b = lp + 1 < 2 * L ? src[(lp + 1) % L] : z;
    reinterpret_cast<float2*>((lo ? BdblLo : Bdbl) + row * LB)[lp] = a;
    reinterpret_cast<float2*>((lo ? BoddLo : Bodd) + row * LB)[lp] = b;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static size_t gemm_smem_bytes() {
    return 1024 + (size_t)kStages * kStageBytes + sizeof(float) * kTileM * kStagePitch + sizeof(GemmBarriers) + 16;
}

static int launch_gemm(const CsdLayout& y, unsigned char* ws, CsdParams p, cudaStream_t st) {
    CUtensorMap mAhi, mAlo, mBhi, mBlo;
    int rc;
    const int64_t arows = (int64_t)y.F * y.MT * kTileM, brows = (int64_t)y.F * y.NT * kTileN;
    if ((rc = make_operand_map(&mAhi, reinterpret_cast<float*>(ws + y.off_ahi), y.KP, arows, kTileM))) return rc;
    if ((rc = make_operand_map(&mAlo, reinterpret_cast<float*>(ws + y.off_alo), y.KP, arows, kTileM))) return rc;
    if ((rc = make_operand_map(&mBhi, reinterpret_cast<float*>(ws + y.off_bhi), y.KP, brows, kTileN))) return rc;
    if ((rc = make_operand_map(&mBlo, reinterpret_cast<float*>(ws + y.off_blo), y.KP, brows, kTileN))) return rc;
    const size_t smem = gemm_smem_bytes();
    rc = ensure_smem_attr(reinterpret_cast<const void*>(csd_gemm_kernel), smem);
    if (rc) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long grid = p.total_tiles < sms ? p.total_tiles : sms;
    csd_gemm_kernel<<<(unsigned)grid, kGemmThreads, smem, st>>>(mAhi, mAlo, mBhi, mBlo, p);
    CMC_CHECK_LAUNCH("csd_gemm_kernel");
    return CMC_OK;
}

// shift-surrogate bookkeeping
__global__ void shift_hist_kernel(const int32_t* __restrict__ shifts, int64_t n, int n_pos, int group,
                                  uint32_t* __restrict__ mult, int32_t* __restrict__ off) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_pos) off[i] = 2 * (int)i * group;
    if (i < n) {
        int s = shifts[i] % n_pos;
        if (s < 0) s += n_pos;
        atomicAdd(&mult[s], 1u);
    }
}
__global__ void shift_gather_kernel(const int32_t* __restrict__ shifts, int64_t n, int n_pos,
                                    const uint32_t* __restrict__ max_u, float* __restrict__ max_stat) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) {
        int s = shifts[i] % n_pos;
        if (s < 0) s += n_pos;
        max_stat[i] = __uint_as_float(max_u[s]);
    }
}

int launch_shift4(const CsdLayout& y, unsigned char* ws, int f_begin, int f_end, int Ne, int Nm, int n_pos,
                  const int32_t* shift_off, const uint32_t* shift_mult, const float* coh_obs, uint32_t* exceed,
                  uint32_t* max_u, cudaStream_t st);
bool csd_direct_ok(const float* X, const float* Y, int64_t ldx, int64_t ldy);
int csd_msc_direct(const float* X, const float* Y, int L, int F, int Ne, int Nm, int64_t ldx, int64_t ldy, float* coh,
                   float* sxx, float* syy, float* sxy, unsigned char* w, const CsdLayout& y, cudaStream_t st);

static int csd_common_check(const float* X, const float* Y, int L, int F, int Ne, int Nm, int64_t ldx, int64_t ldy,
                            const void* ws, const CsdLayout& y) {
    CMC_REQUIRE(X && Y && ws, "cmc_csd_*: null pointer");
    CMC_REQUIRE(L >= 1 && F >= 1 && Ne >= 1 && Nm >= 1 && ldx >= Ne && ldy >= Nm, "cmc_csd_*: bad shape");
    CMC_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "cmc_csd_*: workspace must be 256-byte aligned");
    CMC_REQUIRE((int64_t)F * y.MT * kTileM < (1ll << 31) && (int64_t)F * y.NT * kTileN < (1ll << 31) && F <= 65535,
                "cmc_csd_*: operand too large");
    return CMC_OK;
}

// pack pass alone: operand planes + auto-spectra into the workspace
static int launch_pack(const float* X, const float* Y, int L, int F, int Ne, int Nm, int64_t ldx, int64_t ldy,
                       float* sxx, float* syy, unsigned char* w, const CsdLayout& y, cudaStream_t st) {
    pack_fused_kernel<<<dim3(F, 2 * y.MT + 2 * y.NT), dim3(32, 8), 0, st>>>(
        reinterpret_cast<const float2*>(X), ldx, Ne, reinterpret_cast<const float2*>(Y), ldy, Nm, L, F, y.MT, y.NT, y.KP,
        reinterpret_cast<float*>(w + y.off_ahi), reinterpret_cast<float*>(w + y.off_alo),
        reinterpret_cast<float*>(w + y.off_bhi), reinterpret_cast<float*>(w + y.off_blo),
        reinterpret_cast<float*>(w + y.off_pxx), reinterpret_cast<float*>(w + y.off_pyy), sxx, syy);
    CMC_CHECK_LAUNCH("pack_fused_kernel");
    return CMC_OK;
}

int64_t phase_workspace_bytes(int L, int F, int Ne, int Nm, int64_t n_surr);
int phase_surrogate_null(const void* ws, int L, int F, int Ne, int Nm, uint64_t seed, int64_t s_begin, int64_t s_end,
                         int f_begin, int f_end, const float* coh_obs, uint32_t* exceed, float* max_stat, void* ws2,
                         int64_t ws2_bytes, cudaStream_t st);

int phase_surrogate_hist(const void* ws, int L, int F, int Ne, int Nm, uint64_t seed, int64_t s_begin, int64_t s_end,
                         int f_begin, int f_end, int n_bins, const float* bin_lo, const float* bin_scale,
                         uint32_t* hist, uint32_t* below, void* ws2, int64_t ws2_bytes, bool reuse, cudaStream_t st);

}  // namespace cmc

extern "C" int64_t cmc_csd_workspace_bytes(int L, int F, int Ne, int Nm) {
    if (L < 1 || F < 1 || Ne < 1 || Nm < 1) return CMC_EINVAL;
    return cmc::csd_layout(L, F, Ne, Nm).total;
}

extern "C" int cmc_csd_msc(const float* X, const float* Y, int L, int F, int Ne, int Nm, int64_t ldx, int64_t ldy,
                           float* coh, float* sxx, float* syy, float* sxy, void* ws, int64_t ws_bytes,
                           void* stream) {
    using namespace cmc;
    CMC_REQUIRE(X && Y && coh && ws, "cmc_csd_msc: null pointer");
    CMC_REQUIRE(L >= 1 && F >= 1 && Ne >= 1 && Nm >= 1 && ldx >= Ne && ldy >= Nm, "cmc_csd_msc: bad shape");
    const CsdLayout y = csd_layout(L, F, Ne, Nm);
    if (ws_bytes < y.total) {
        set_error("cmc_csd_msc: workspace %lld < %lld bytes", (long long)ws_bytes, (long long)y.total);
        return CMC_EWORKSPACE;
    }
    CMC_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "cmc_csd_msc: workspace must be 256-byte aligned");
    CMC_REQUIRE((int64_t)F * y.MT * kTileM < (1ll << 31) && (int64_t)F * y.NT * kTileN < (1ll << 31) && F <= 65535,
                "cmc_csd_msc: operand too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char* w = static_cast<unsigned char*>(ws);
    float* pxx = reinterpret_cast<float*>(w + y.off_pxx);
    float* pyy = reinterpret_cast<float*>(w + y.off_pyy);
    pack_fused_kernel<<<dim3(F, 2 * y.MT + 2 * y.NT), dim3(32, 8), 0, st>>>(
        reinterpret_cast<const float2*>(X), ldx, Ne, reinterpret_cast<const float2*>(Y), ldy, Nm, L, F, y.MT, y.NT, y.KP,
        reinterpret_cast<float*>(w + y.off_ahi), reinterpret_cast<float*>(w + y.off_alo),
        reinterpret_cast<float*>(w + y.off_bhi), reinterpret_cast<float*>(w + y.off_blo), pxx, pyy, sxx, syy);
    CMC_CHECK_LAUNCH("pack_fused_kernel");
    CsdParams p{};
    p.F = F; p.MT = y.MT; p.NT = y.NT; p.Ne = Ne; p.Nm = Nm; p.KB = y.KP / kKBlock;
    p.coh = coh; p.sxy = reinterpret_cast<float2*>(sxy); p.pxx = pxx; p.pyy = pyy;
    p.total_tiles = (long long)F * y.MT * y.NT;
    return launch_gemm(y, w, p, st);
}

extern "C" int64_t cmc_csd_workspace_bytes_min(int F, int Ne, int Nm) {
    if (F < 1 || Ne < 1 || Nm < 1) return CMC_EINVAL;
    return cmc::csd_layout(1, F, Ne, Nm).off_ahi;                   // auto-spectra only
}

extern "C" int cmc_csd_coherence(const float* X, const float* Y, int L, int F, int Ne, int Nm, int64_t ldx, int64_t ldy,
                                 float* coh, float* sxx, float* syy, float* sxy, void* ws, int64_t ws_bytes,
                                 void* stream) {
    using namespace cmc;
    const CsdLayout y = csd_layout(L, F, Ne, Nm);
    int rc = csd_common_check(X, Y, L, F, Ne, Nm, ldx, ldy, ws, y);
    if (rc) return rc;
    CMC_REQUIRE(coh, "cmc_csd_coherence: null pointer");
    static const bool no_direct = getenv("CMC_CSD_NO_DIRECT") != nullptr;
    if (no_direct || !csd_direct_ok(X, Y, ldx, ldy))     // odd channel pitch: TMA cannot address the rows
        return cmc_csd_msc(X, Y, L, F, Ne, Nm, ldx, ldy, coh, sxx, syy, sxy, ws, ws_bytes, stream);
    if (ws_bytes < y.off_ahi) {
        set_error("cmc_csd_coherence: workspace %lld < %lld bytes", (long long)ws_bytes, (long long)y.off_ahi);
        return CMC_EWORKSPACE;
    }
    return csd_msc_direct(X, Y, L, F, Ne, Nm, ldx, ldy, coh, sxx, syy, sxy, static_cast<unsigned char*>(ws), y,
                          static_cast<cudaStream_t>(stream));
}

extern "C" int cmc_csd_operands(const float* X, const float* Y, int L, int F, int Ne, int Nm, int64_t ldx, int64_t ldy,
                                void* ws, int64_t ws_bytes, void* stream) {
    using namespace cmc;
    const CsdLayout y = csd_layout(L, F, Ne, Nm);
    int rc = csd_common_check(X, Y, L, F, Ne, Nm, ldx, ldy, ws, y);
    if (rc) return rc;
    if (ws_bytes < y.total) {
        set_error("cmc_csd_operands: workspace %lld < %lld bytes", (long long)ws_bytes, (long long)y.total);
        return CMC_EWORKSPACE;
    }
    return launch_pack(X, Y, L, F, Ne, Nm, ldx, ldy, nullptr, nullptr, static_cast<unsigned char*>(ws), y,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int64_t cmc_surrogate_workspace_bytes(int L, int F, int Ne, int Nm, int mode, int64_t n_surr) {
    if (L < 1 || F < 1 || Ne < 1 || Nm < 1 || n_surr < 0) return CMC_EINVAL;
    if (mode == CMC_SURR_SHIFT) return (int64_t)L * 12 + 256;
    if (mode == CMC_SURR_PHASE) return cmc::phase_workspace_bytes(L, F, Ne, Nm, n_surr);
    return CMC_EINVAL;
}

extern "C" int cmc_surrogate_null(void* ws, int L, int F, int Ne, int Nm, int mode, int group,
                                  const int32_t* shifts, uint64_t seed, int64_t s_begin, int64_t s_end,
                                  const float* coh_obs, uint32_t* exceed, float* max_stat, void* ws2,
                                  int64_t ws2_bytes, void* stream) {
    return cmc_surrogate_null_range(ws, L, F, Ne, Nm, mode, group, shifts, seed, s_begin, s_end, 0, F, coh_obs, exceed,
                                    max_stat, ws2, ws2_bytes, stream);
}

extern "C" int cmc_surrogate_null_range(void* ws, int L, int F, int Ne, int Nm, int mode, int group,
                                        const int32_t* shifts, uint64_t seed, int64_t s_begin, int64_t s_end,
                                        int f_begin, int f_end, const float* coh_obs, uint32_t* exceed,
                                        float* max_stat, void* ws2, int64_t ws2_bytes, void* stream) {
    using namespace cmc;
    CMC_REQUIRE(ws && coh_obs && exceed && max_stat && ws2, "cmc_surrogate_null: null pointer");
    CMC_REQUIRE(s_end >= s_begin, "cmc_surrogate_null: bad surrogate range");
    CMC_REQUIRE(0 <= f_begin && f_begin <= f_end && f_end <= F, "cmc_surrogate_null: bad frequency range [%d, %d)",
                f_begin, f_end);
    const int64_t n = s_end - s_begin;
    if (n == 0) return CMC_OK;
    if (f_begin == f_end) {                      // nothing to compare against: every max statistic is 0
        return check_cuda(cudaMemsetAsync(max_stat, 0, (size_t)n * 4, static_cast<cudaStream_t>(stream)),
                          "memset(max_stat)");
    }
    if (mode == CMC_SURR_PHASE)
        return phase_surrogate_null(ws, L, F, Ne, Nm, seed, s_begin, s_end, f_begin, f_end, coh_obs, exceed, max_stat,
                                    ws2, ws2_bytes, static_cast<cudaStream_t>(stream));
    CMC_REQUIRE(mode == CMC_SURR_SHIFT, "cmc_surrogate_null: unknown mode %d", mode);
    CMC_REQUIRE(shifts, "cmc_surrogate_null: shift mode needs a shift table");
    CMC_REQUIRE(group >= 1 && L % group == 0, "cmc_surrogate_null: group must divide L");
    const int n_pos = L / group;
    if (ws2_bytes < cmc_surrogate_workspace_bytes(L, F, Ne, Nm, mode, n)) {
        set_error("cmc_surrogate_null: workspace too small");
        return CMC_EWORKSPACE;
    }
    const CsdLayout y = csd_layout(L, F, Ne, Nm);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char* w = static_cast<unsigned char*>(ws);
    uint32_t* mult = static_cast<uint32_t*>(ws2);
    uint32_t* max_u = mult + L;
    int32_t* off = reinterpret_cast<int32_t*>(max_u + L);
    int rc = check_cuda(cudaMemsetAsync(ws2, 0, (size_t)L * 8, st), "memset(shift tables)");
    if (rc) return rc;
    const int64_t nthreads = n > n_pos ? n : n_pos;
    shift_hist_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(shifts, n, n_pos, group, mult, off);
    CMC_CHECK_LAUNCH("shift_hist_kernel");
    // doubled / advanced views of the B operand (rebuilt per call: ~10 us against milliseconds of GEMM)
    shift_operand_kernel<<<dim3((unsigned)((int64_t)F * y.NT * kTileN), (y.LB / 2 + 255) / 256, 2), 256, 0, st>>>(
        reinterpret_cast<const float*>(w + y.off_bhi), reinterpret_cast<const float*>(w + y.off_blo), L, y.KP, y.LB,
        reinterpret_cast<float*>(w + y.off_bdbl), reinterpret_cast<float*>(w + y.off_bodd),
        reinterpret_cast<float*>(w + y.off_bdbl_lo), reinterpret_cast<float*>(w + y.off_bodd_lo));
    CMC_CHECK_LAUNCH("shift_operand_kernel");
    // four shifts share one staged A tile, three TF32 terms per product (csd_shift.cu)
    rc = launch_shift4(y, w, f_begin, f_end, Ne, Nm, n_pos, off, mult, coh_obs, exceed, max_u, st);
    if (rc) return rc;
    shift_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(shifts, n, n_pos, max_u, max_stat);
    CMC_CHECK_LAUNCH("shift_gather_kernel");
    return CMC_OK;
}

extern "C" int cmc_surrogate_null_hist(void* ws, int L, int F, int Ne, int Nm, int mode, uint64_t seed,
                                       int64_t s_begin, int64_t s_end, int f_begin, int f_end, int n_bins,
                                       const float* bin_lo, const float* bin_scale, uint32_t* hist, uint32_t* below,
                                       void* ws2, int64_t ws2_bytes, int reuse_operands, void* stream) {
    using namespace cmc;
    CMC_REQUIRE(ws && hist && ws2, "cmc_surrogate_null_hist: null pointer");
    CMC_REQUIRE(s_end >= s_begin, "cmc_surrogate_null_hist: bad surrogate range");
    CMC_REQUIRE(0 <= f_begin && f_begin <= f_end && f_end <= F, "cmc_surrogate_null_hist: bad frequency range [%d, %d)",
                f_begin, f_end);
    if (mode != CMC_SURR_PHASE) {
        set_error("cmc_surrogate_null_hist: histograms are built for phase surrogates only (a shift null has at most "
                  "L - 1 distinct values per pair)");
        return CMC_EUNSUPPORTED;
    }
    if (s_end == s_begin || f_begin == f_end) return CMC_OK;
    return phase_surrogate_hist(ws, L, F, Ne, Nm, seed, s_begin, s_end, f_begin, f_end, n_bins, bin_lo, bin_scale, hist,
                                below, ws2, ws2_bytes, reuse_operands != 0, static_cast<cudaStream_t>(stream));
}

// K4: cluster-based permutation test.  One CTA per permutation: sign-flip t-map (fp64, numpy
// operation order) -> threshold -> connected components over the CSR adjacency restricted to
// supra-threshold nodes (lock-free union-find in shared memory, smaller index wins so every
// cluster is rooted at its smallest flat index) -> order-free int64 fixed-point cluster
// masses -> signed mass of largest magnitude.
//
// Replaces the body of mne.stats.permutation_cluster_1samp_test as called by the reference at
// src/pipeline/cbpa.py:1027-1042; algorithm restated in oracle/cbpa.py.
#include "common.cuh"

namespace cmc {

constexpr int kCbpaThreads = 512;
constexpr int kCbpaMaxTests = 16384;

#ifdef CMC_CBPA_PROFILE
// instrumented build only: per-phase clock64() totals of thread 0 of every CTA
__device__ unsigned long long g_cbpa_cycles[8];
#define CBPA_TICK(i) do { if (tid == 0) { const long long now_ = clock64(); \
    atomicAdd(&g_cbpa_cycles[i], (unsigned long long)(now_ - tick_)); tick_ = now_; } } while (0)
#else
#define CBPA_TICK(i) do { } while (0)
#endif

__device__ __forceinline__ int uf_find(int* parent, int x) {
    int p = parent[x];
    while (p != x) {
        const int g = parent[p];
        if (g != p) parent[x] = g;  // path halving; racing writers only ever store an ancestor
        x = p;
        p = g;
    }
    return x;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    while (a != b) {
        if (a < b) { int t = a; a = b; b = t; }   // hook the larger root under the smaller
        const int old = atomicCAS(&parent[a], a, b);
        if (old == a) break;
        a = uf_find(parent, old);
        b = uf_find(parent, b);
    }
}

// The data are re-laid once per call as XT[test / 32][subject][test % 32] (tests padded to a multiple of 32):
// the subject column of 32 consecutive tests is then NS coalesced 256-byte rows at compile-time offsets from one
// base address - no per-load address arithmetic in the t-map loop.
__global__ void __launch_bounds__(256)
cbpa_tile_kernel(const double* __restrict__ X, int n_subj, int n_tests, double* __restrict__ XT) {
    const int64_t n_pad = ((int64_t)n_tests + 31) & ~31LL;
    const int64_t total = n_pad * n_subj;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int lane = (int)(i & 31);
        const int64_t rest = i >> 5;
        const int s = (int)(rest % n_subj);
        const int64_t v = (rest / n_subj) * 32 + lane;
        XT[i] = v < n_tests ? X[(int64_t)s * n_tests + v] : 0.0;
    }
}

__device__ __forceinline__ const double* tile_column(const double* __restrict__ XT, int n_subj, int v) {
    return XT + (int64_t)(v >> 5) * n_subj * 32 + (v & 31);
}

// t = mean / sqrt(var(ddof=1) / n) with numpy's evaluation order (oracle/cbpa.py:ttest_1samp_no_p):
// sequential sums over subjects, no fused multiply-add.  Generic subject count.
__device__ __forceinline__ double t_stat(const double* __restrict__ XT, const double* sg, int n_subj, int v) {
    const double* col = tile_column(XT, n_subj, v);
    double sum = __dmul_rn(col[0], sg[0]);
    for (int s = 1; s < n_subj; ++s) sum = __dadd_rn(sum, __dmul_rn(col[s * 32], sg[s]));
    const double n = (double)n_subj;
    const double mean = __ddiv_rn(sum, n);
    double d = __dsub_rn(__dmul_rn(col[0], sg[0]), mean);
    double ss = __dmul_rn(d, d);
    for (int s = 1; s < n_subj; ++s) {
        d = __dsub_rn(__dmul_rn(col[s * 32], sg[s]), mean);
        ss = __dadd_rn(ss, __dmul_rn(d, d));
    }
    const double var = __ddiv_rn(ss, (double)(n_subj - 1));
    return __ddiv_rn(mean, __dsqrt_rn(__ddiv_rn(var, n)));
}

// a / B for a compile-time integer B without the generic division sequence: q = RN(a * RN(1/B)) is within one
// ulp of a / B, and one exact-remainder correction r = a - B q (an FMA), q' = RN(q + r * RN(1/B)) then yields
// the correctly rounded quotient (Markstein's theorem), i.e. exactly what __ddiv_rn returns - provided a sits
// in a safe exponent band (no zeros, subnormals, huge values, Inf/NaN), which div_safe() tests.
template <int B>
__device__ __forceinline__ double div_fast(double a) {
    constexpr double y = 1.0 / (double)B;
    const double q = __dmul_rn(a, y);
    const double r = __fma_rn(-(double)B, q, a);
    return __fma_rn(r, y, q);
}
__device__ __forceinline__ bool div_safe(double a) {
    return (((unsigned)__double2hiint(a) & 0x7fffffffu) - (100u << 20)) < (1800u << 20);
}

// Out-of-line form of t_stat with sign-bit masks, for the rare columns t_stat_regs hands over.
__device__ __noinline__ double t_stat_flip(const double* __restrict__ XT, const unsigned* flip, int n_subj, int v) {
    const double* col = tile_column(XT, n_subj, v);
    auto x = [&](int s) {
        const double a = col[s * 32];
        return __hiloint2double(__double2hiint(a) ^ (int)flip[s], __double2loint(a));
    };
    double sum = x(0);
    for (int s = 1; s < n_subj; ++s) sum = __dadd_rn(sum, x(s));
    const double n = (double)n_subj;
    const double mean = __ddiv_rn(sum, n);
    double d = __dsub_rn(x(0), mean);
    double ss = __dmul_rn(d, d);
    for (int s = 1; s < n_subj; ++s) {
        d = __dsub_rn(x(s), mean);
        ss = __dadd_rn(ss, __dmul_rn(d, d));
    }
    const double var = __ddiv_rn(ss, (double)(n_subj - 1));
    return __ddiv_rn(mean, __dsqrt_rn(__ddiv_rn(var, n)));
}

// Same arithmetic as t_stat with the subject column held in registers and the subject count known at compile
// time: the NS loads are independent (one L2 latency), the second pass re-uses them, and the sign flip is an
// exact XOR of the sign bit (x * (+-1) == x with the sign bit flipped, NaNs included).  The three divisions by
// constants take the two-FMA form; one range test covers all three and falls back to the generic divisions.
// Evaluation order is unchanged, so t stays bit-identical to the oracle.  (profiles/r01c_cbpa.md)
// Returns mean and var / n (t = mean / sqrt(var / n)); false = outside the safe band, use t_stat_flip.
template <int NS>
__device__ __forceinline__ bool t_moments_regs(const double* __restrict__ XT, const unsigned* flip, int v,
                                               double& mean_out, double& vn_out) {
    const double* col = tile_column(XT, NS, v);
    double xs[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) xs[s] = col[s * 32];
#pragma unroll
    for (int s = 0; s < NS; ++s)
        xs[s] = __hiloint2double(__double2hiint(xs[s]) ^ (int)flip[s], __double2loint(xs[s]));
    double sum = xs[0];
#pragma unroll
    for (int s = 1; s < NS; ++s) sum = __dadd_rn(sum, xs[s]);
    const double mean = div_fast<NS>(sum);
    double d = __dsub_rn(xs[0], mean);
    double ss = __dmul_rn(d, d);
#pragma unroll
    for (int s = 1; s < NS; ++s) {
        d = __dsub_rn(xs[s], mean);
        ss = __dadd_rn(ss, __dmul_rn(d, d));
    }
    const double var = div_fast<NS - 1>(ss);
    const double vn = div_fast<NS>(var);
    mean_out = mean;
    vn_out = vn;
    return div_safe(sum) && div_safe(ss) && div_safe(var);
}

__device__ __forceinline__ signed char supra_sign(double t, double thr, int tail) {
    if (tail == 0) return (t > thr) ? 1 : ((t < -thr) ? -1 : 0);
    if (tail > 0) return (t > thr) ? 1 : 0;
    return (t < thr) ? -1 : 0;
}

__device__ __forceinline__ long long t_to_fixed(double t) {
    t = fmin(fmax(t, -CMC_T_CLAMP), CMC_T_CLAMP);
    return __double2ll_rn(t * (double)(1 << CMC_FIX_SHIFT));
}

// Permutation CTAs (OBSERVED == false) keep a compact list of the supra-threshold nodes (a few per cent of the
// map under H0): the t-map pass appends to it, and the hook / mass / max passes then walk the list instead of
// the whole map - eight lanes per listed node over its CSR row, next row prefetched - so their cost follows the
// number of supra-threshold nodes and every lane has work.  If the list overflows its capacity the CTA falls
// back to the whole-map passes for that permutation; results are identical either way (unions and the integer
// masses are order-free).  (profiles/r01c_cbpa.md)
template <bool OBSERVED, int NS>
__global__ void __launch_bounds__(kCbpaThreads)
cbpa_kernel(const double* __restrict__ XT, int n_subj, int n_tests, const int8_t* __restrict__ signs,
            int64_t n_perm, double thr, int tail, const int32_t* __restrict__ indptr,
            const int32_t* __restrict__ indices, long long* __restrict__ h0,
            double* __restrict__ t_obs, int32_t* __restrict__ root_out, long long* __restrict__ mass_out,
            int list_cap, unsigned* __restrict__ claim) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    long long* mass = reinterpret_cast<long long*>(smem_raw);                 // [n_tests]
    int* parent = reinterpret_cast<int*>(mass + n_tests);                      // [n_tests]
    double* sg = reinterpret_cast<double*>(parent + ((n_tests + 1) & ~1));      // [n_subj]
    unsigned* flip = reinterpret_cast<unsigned*>(sg + n_subj);                 // [n_subj] sign-bit masks
    signed char* sgn = reinterpret_cast<signed char*>(flip + ((n_subj + 1) & ~1));   // [n_tests]
    unsigned short* list = reinterpret_cast<unsigned short*>(sgn + ((n_tests + 1) & ~1));   // [list_cap]
    __shared__ long long red_abs[kCbpaThreads / 32];
    __shared__ long long red_val[kCbpaThreads / 32];
    __shared__ int n_supra;
    __shared__ long long next_p;
    const int tid = threadIdx.x;

#ifdef CMC_CBPA_PROFILE
    long long tick_ = clock64();
#endif
    // the first permutation of a CTA is static, every further one is claimed from a device counter (zeroed by the
    // host before the launch): CTAs that drew cheap permutations (few supra-threshold nodes) take more of them
    for (int64_t p = blockIdx.x; p < n_perm;) {
        __syncthreads();
        CBPA_TICK(0);
        if (tid == 0) next_p = claim ? (long long)gridDim.x + atomicAdd(claim, 1u) : (long long)p + gridDim.x;
        for (int s = tid; s < n_subj; s += kCbpaThreads) {
            const int sv = OBSERVED ? 1 : (int)signs[p * n_subj + s];
            sg[s] = (double)sv;
            flip[s] = sv < 0 ? 0x80000000u : 0u;
        }
        if (tid == 0) n_supra = 0;
        __syncthreads();
        // ---- t-map, threshold, fixed-point image (parent / mass are only ever read at supra-threshold nodes) ----
        for (int v = tid; v < n_tests; v += kCbpaThreads) {
            double t = 0.0;
            signed char s = 0;
            if (NS > 0) {
                constexpr int NSX = NS > 0 ? NS : 2;
                double mean, vn;
                const bool ok = t_moments_regs<NSX>(XT, flip, v, mean, vn);
                t = ok ? __ddiv_rn(mean, __dsqrt_rn(vn)) : t_stat_flip(XT, flip, n_subj, v);
                s = supra_sign(t, thr, tail);
            } else {
                t = t_stat(XT, sg, n_subj, v);
                s = supra_sign(t, thr, tail);
            }
            sgn[v] = s;
            if (s) {
                parent[v] = v;
                mass[v] = t_to_fixed(t);
                if (!OBSERVED) {
                    const int k = atomicAdd(&n_supra, 1);
                    if (k < list_cap) list[k] = (unsigned short)v;
                }
            }
            if (OBSERVED) t_obs[v] = t;
        }
        __syncthreads();
        CBPA_TICK(1);
        const int ns = OBSERVED ? 0 : n_supra;
        const bool compact = !OBSERVED && ns <= list_cap;     // CTA-uniform
        long long best_abs = -1, best_val = 0;
        if (compact) {
            // ---- hook: eight lanes per listed node, one CSR row each; the next row is fetched ahead ----
            const int sub = tid & 7, grp = tid >> 3;
            constexpr int kGroups = kCbpaThreads >> 3;
            int i = grp;
            int v = 0, e0 = 0, e1 = 0;
            if (i < ns) { v = list[i]; e0 = indptr[v]; e1 = indptr[v + 1]; }
            while (i < ns) {
                const int inext = i + kGroups;
                int vn = 0, e0n = 0, e1n = 0;
                if (inext < ns) { vn = list[inext]; e0n = indptr[vn]; e1n = indptr[vn + 1]; }
                const signed char s = sgn[v];
                for (int e = e0 + sub; e < e1; e += 8) {
                    const int u = indices[e];
                    if (u < v && sgn[u] == s) uf_union(parent, u, v);
                }
                i = inext; v = vn; e0 = e0n; e1 = e1n;
            }
            __syncthreads();
            CBPA_TICK(2);
            // ---- cluster mass at the root ----
            for (int k = tid; k < ns; k += kCbpaThreads) {
                const int w = list[k];
                const int r = uf_find(parent, w);
                if (r != w) atomicAdd(reinterpret_cast<unsigned long long*>(&mass[r]),
                                      static_cast<unsigned long long>(mass[w]));
            }
            __syncthreads();
            CBPA_TICK(3);
            for (int k = tid; k < ns; k += kCbpaThreads) {
                const int w = list[k];
                if (parent[w] == w) {
                    const long long m = mass[w];
                    const long long a = m < 0 ? -m : m;
                    if (a > best_abs || (a == best_abs && m > best_val)) { best_abs = a; best_val = m; }
                }
            }
        } else {
            // ---- hook every supra-threshold edge once (u < v), same sign only ----
            for (int v = tid; v < n_tests; v += kCbpaThreads) {
                const signed char s = sgn[v];
                if (!s) continue;
                const int e1 = indptr[v + 1];
                for (int e = indptr[v]; e < e1; ++e) {
                    const int u = indices[e];
                    if (u < v && sgn[u] == s) uf_union(parent, u, v);
                }
            }
            __syncthreads();
            CBPA_TICK(2);
            // ---- cluster mass at the root (smallest index of the component) ----
            for (int v = tid; v < n_tests; v += kCbpaThreads) {
                if (!sgn[v]) { if (OBSERVED) root_out[v] = -1; continue; }
                const int r = uf_find(parent, v);
                if (r != v) atomicAdd(reinterpret_cast<unsigned long long*>(&mass[r]),
                                      static_cast<unsigned long long>(mass[v]));
                if (OBSERVED) root_out[v] = r;
            }
            __syncthreads();
            CBPA_TICK(3);
            for (int v = tid; v < n_tests; v += kCbpaThreads) {
                if (sgn[v] && parent[v] == v) {
                    const long long m = mass[v];
                    const long long a = m < 0 ? -m : m;
                    if (a > best_abs || (a == best_abs && m > best_val)) { best_abs = a; best_val = m; }
                }
                if (OBSERVED) mass_out[v] = (sgn[v] && parent[v] == v) ? mass[v] : 0;
            }
        }
        // ---- signed mass of largest magnitude; a positive cluster wins an exact tie ----
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const long long oa = __shfl_xor_sync(0xffffffffu, best_abs, off);
            const long long ov = __shfl_xor_sync(0xffffffffu, best_val, off);
            if (oa > best_abs || (oa == best_abs && ov > best_val)) { best_abs = oa; best_val = ov; }
        }
        if ((tid & 31) == 0) { red_abs[tid >> 5] = best_abs; red_val[tid >> 5] = best_val; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kCbpaThreads / 32; ++w)
                if (red_abs[w] > best_abs || (red_abs[w] == best_abs && red_val[w] > best_val)) {
                    best_abs = red_abs[w];
                    best_val = red_val[w];
                }
            h0[p] = best_abs < 0 ? 0 : best_val;
        }
        CBPA_TICK(4);
        p = next_p;                 // written before this iteration's first barrier, read after its last one
    }
}

// Canonical cluster numbering of the observed map: t > thr clusters first, each group ordered by
// root (= smallest flat index).  Single CTA; n_tests is small.
__global__ void __launch_bounds__(1024)
cbpa_label_kernel(const int32_t* __restrict__ root, const long long* __restrict__ mass_root, int n_tests,
                  int32_t* __restrict__ labels, long long* __restrict__ mass_fixed,
                  double* __restrict__ mass_f64, int32_t* __restrict__ n_clusters, int32_t* __restrict__ rank) {
    __shared__ int warp_cnt[32];
    __shared__ int running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int pass = 0; pass < 2; ++pass) {          // 0: positive roots, 1: negative roots
        for (int base = 0; base < n_tests; base += 1024) {
            const int v = base + tid;
            bool is_root = false;
            if (v < n_tests && root[v] == v) is_root = pass == 0 ? (mass_root[v] >= 0) : (mass_root[v] < 0);
            const unsigned bal = __ballot_sync(0xffffffffu, is_root);
            if (lane == 0) warp_cnt[warp] = __popc(bal);
            __syncthreads();
            int off = running;
            for (int w = 0; w < warp; ++w) off += warp_cnt[w];
            if (is_root) {
                const int k = off + __popc(bal & ((1u << lane) - 1));
                rank[v] = k;
                mass_fixed[k] = mass_root[v];
                mass_f64[k] = (double)mass_root[v] / (double)(1 << CMC_FIX_SHIFT);
            }
            __syncthreads();
            if (tid == 0) {
                int tot = 0;
                for (int w = 0; w < 32; ++w) tot += warp_cnt[w];
                running += tot;
            }
            __syncthreads();
        }
        __syncthreads();
    }
    if (tid == 0) *n_clusters = running;
    for (int v = tid; v < n_tests; v += 1024) labels[v] = root[v] >= 0 ? rank[root[v]] + 1 : 0;
}

static size_t cbpa_smem_bytes(int n_subj, int n_tests, int list_cap) {
    return sizeof(long long) * n_tests + sizeof(int) * ((n_tests + 1) & ~1) + sizeof(double) * n_subj +
           sizeof(unsigned) * ((n_subj + 1) & ~1) + ((n_tests + 1) & ~1) + sizeof(unsigned short) * list_cap + 16;
}

// Capacity of the supra-threshold list: the whole map when two CTAs still fit on one SM, else what is left of
// the 227 KB a CTA may use (the kernel falls back to whole-map passes when a permutation overflows it).
static int cbpa_list_cap(int n_subj, int n_tests) {
    const size_t base = cbpa_smem_bytes(n_subj, n_tests, 0);
    const size_t two_per_sm = 112 * 1024, one_per_sm = 226 * 1024;
    size_t room = 0;
    if (base + 2 * (size_t)n_tests <= two_per_sm) room = 2 * (size_t)n_tests;
    else if (base < two_per_sm && (two_per_sm - base) / 2 >= (size_t)n_tests / 4) room = two_per_sm - base;
    else if (base < one_per_sm) room = one_per_sm - base;
    const size_t cap = room / 2;
    return (int)(cap < (size_t)n_tests ? cap : (size_t)n_tests);
}

static int cbpa_check(int n_subj, int n_tests, const int32_t* indptr, const int32_t* indices, int tail, double thr) {
    CMC_REQUIRE(indptr && indices, "cmc_cbpa: null pointer");
    CMC_REQUIRE(n_subj >= 2, "cmc_cbpa: need at least 2 subjects");
    CMC_REQUIRE(n_tests >= 1 && n_tests <= kCbpaMaxTests, "cmc_cbpa: n_tests=%d outside [1, %d]", n_tests,
                kCbpaMaxTests);
    CMC_REQUIRE(tail >= -1 && tail <= 1, "cmc_cbpa: tail must be -1, 0 or 1");
    CMC_REQUIRE(!(tail == 0 && thr < 0), "cmc_cbpa: two-tailed test needs a non-negative threshold");
    return CMC_OK;
}

}  // namespace cmc

#ifdef CMC_CBPA_PROFILE
extern "C" CMC_API int cmc_dbg_cbpa_cycles(unsigned long long* out, int reset) {
    cudaMemcpyFromSymbol(out, cmc::g_cbpa_cycles, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {}; cudaMemcpyToSymbol(cmc::g_cbpa_cycles, z, sizeof(z)); }
    return 0;
}
#endif

namespace cmc {
static int64_t cbpa_labels_bytes(int n_tests) {
    return ((int64_t)n_tests * 16 + 256 + 255) & ~255LL;           // root int32 + rank int32 + mass_root int64 per test
}
// Re-tiles X into the workspace (after the labelling scratch) and returns the tiled copy.  X == nullptr: the
// workspace already holds the tiled copy of an earlier call on the same data (cmc_cbpa_observed followed by
// cmc_cbpa_permute on one stream), only the pointer is returned.
static int cbpa_tile(const double* X, int n_subj, int n_tests, void* ws, int64_t ws_bytes, cudaStream_t st,
                     const double** XT, const char* who) {
    if (!ws || ws_bytes < cmc_cbpa_workspace_bytes(n_subj, n_tests)) {
        set_error("%s: workspace %lld < %lld bytes", who, (long long)ws_bytes,
                  (long long)cmc_cbpa_workspace_bytes(n_subj, n_tests));
        return CMC_EWORKSPACE;
    }
    CMC_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "%s: workspace must be 16-byte aligned", who);
    double* xt = reinterpret_cast<double*>(static_cast<unsigned char*>(ws) + cbpa_labels_bytes(n_tests));
    if (X) {
        const int64_t total = (((int64_t)n_tests + 31) & ~31LL) * n_subj;
        const int64_t blocks = (total + 255) / 256;
        cbpa_tile_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, st>>>(X, n_subj, n_tests, xt);
        CMC_CHECK_LAUNCH("cbpa_tile_kernel");
    }
    *XT = xt;
    return CMC_OK;
}
}  // namespace cmc

extern "C" int64_t cmc_cbpa_workspace_bytes(int n_subj, int n_tests) {
    // labelling scratch + the tiled copy of X (tests padded to a multiple of 32)
    return cmc::cbpa_labels_bytes(n_tests) + (((int64_t)n_tests + 31) & ~31LL) * n_subj * 8;
}

extern "C" int cmc_cbpa_permute(const double* X, int n_subj, int n_tests, const int8_t* signs,
                                int64_t p_begin, int64_t p_end, double thr, int tail,
                                const int32_t* indptr, const int32_t* indices, int64_t* h0_fixed,
                                void* ws, int64_t ws_bytes, void* stream) {
    using namespace cmc;
    int rc = cbpa_check(n_subj, n_tests, indptr, indices, tail, thr);
    if (rc) return rc;
    CMC_REQUIRE(p_end >= p_begin, "cmc_cbpa_permute: bad permutation range");
    if (p_end == p_begin) return CMC_OK;
    CMC_REQUIRE(signs && h0_fixed, "cmc_cbpa_permute: null pointer");
    const int64_t n_perm = p_end - p_begin;
    if (n_perm == 0) return CMC_OK;
    const double* XT = nullptr;
    rc = cbpa_tile(X, n_subj, n_tests, ws, ws_bytes, static_cast<cudaStream_t>(stream), &XT, "cmc_cbpa_permute");
    if (rc) return rc;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int list_cap = cbpa_list_cap(n_subj, n_tests);
    const size_t smem = cbpa_smem_bytes(n_subj, n_tests, list_cap);
    // exact subject count as a template parameter for the usual group sizes, generic loop otherwise
    using KernT = void (*)(const double*, int, int, const int8_t*, int64_t, double, int, const int32_t*,
                           const int32_t*, long long*, double*, int32_t*, long long*, int, unsigned*);
    KernT kern = cbpa_kernel<false, 0>;
    switch (n_subj) {
#define CMC_CBPA_CASE(n) case n: kern = cbpa_kernel<false, n>; break;
        CMC_CBPA_CASE(2) CMC_CBPA_CASE(3) CMC_CBPA_CASE(4) CMC_CBPA_CASE(5) CMC_CBPA_CASE(6) CMC_CBPA_CASE(7)
        CMC_CBPA_CASE(8) CMC_CBPA_CASE(9) CMC_CBPA_CASE(10) CMC_CBPA_CASE(11) CMC_CBPA_CASE(12) CMC_CBPA_CASE(13)
        CMC_CBPA_CASE(14) CMC_CBPA_CASE(15) CMC_CBPA_CASE(16) CMC_CBPA_CASE(17) CMC_CBPA_CASE(18) CMC_CBPA_CASE(19)
        CMC_CBPA_CASE(20) CMC_CBPA_CASE(21) CMC_CBPA_CASE(22) CMC_CBPA_CASE(23) CMC_CBPA_CASE(24) CMC_CBPA_CASE(25)
        CMC_CBPA_CASE(26) CMC_CBPA_CASE(27) CMC_CBPA_CASE(28) CMC_CBPA_CASE(29) CMC_CBPA_CASE(30) CMC_CBPA_CASE(31)
        CMC_CBPA_CASE(32)
#undef CMC_CBPA_CASE
        default: break;
    }
    rc = ensure_smem_attr(reinterpret_cast<const void*>(kern), smem);
    if (rc) return rc;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCbpaThreads, smem);
    if (per_sm < 1) per_sm = 1;
    const int64_t grid = n_perm < (int64_t)sms * per_sm ? n_perm : (int64_t)sms * per_sm;
    // permutation claim counter: the first word of the labelling scratch (unused by this call)
    unsigned* claim = static_cast<unsigned*>(ws);
    static const bool static_split = getenv("CMC_CBPA_STATIC_SPLIT") != nullptr;
    if (static_split) claim = nullptr;
    else if ((rc = check_cuda(cudaMemsetAsync(claim, 0, 4, static_cast<cudaStream_t>(stream)), "memset(claim)"))) return rc;
    kern<<<(unsigned)grid, kCbpaThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        XT, n_subj, n_tests, signs + p_begin * n_subj, n_perm, thr, tail, indptr, indices,
        reinterpret_cast<long long*>(h0_fixed), nullptr, nullptr, nullptr, list_cap, claim);
    CMC_CHECK_LAUNCH("cbpa_kernel<perm>");
    return CMC_OK;
}

extern "C" int cmc_cbpa_observed(const double* X, int n_subj, int n_tests, double thr, int tail,
                                 const int32_t* indptr, const int32_t* indices, double* t_obs,
                                 int32_t* labels, int64_t* mass_fixed, double* mass_f64,
                                 int32_t* n_clusters, void* ws, int64_t ws_bytes, void* stream) {
    using namespace cmc;
    int rc = cbpa_check(n_subj, n_tests, indptr, indices, tail, thr);
    if (rc) return rc;
    CMC_REQUIRE(X && t_obs && labels && mass_fixed && mass_f64 && n_clusters, "cmc_cbpa_observed: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double* XT = nullptr;
    rc = cbpa_tile(X, n_subj, n_tests, ws, ws_bytes, st, &XT, "cmc_cbpa_observed");
    if (rc) return rc;
    long long* mass_root = reinterpret_cast<long long*>(ws);                    // 8-byte items first
    long long* h0_tmp = mass_root + n_tests;
    int32_t* root = reinterpret_cast<int32_t*>(h0_tmp + 1);
    int32_t* rank = root + n_tests;
    const size_t smem = cbpa_smem_bytes(n_subj, n_tests, 0);
    rc = ensure_smem_attr(reinterpret_cast<const void*>(cbpa_kernel<true, 0>), smem);
    if (rc) return rc;
    cbpa_kernel<true, 0><<<1, kCbpaThreads, smem, st>>>(XT, n_subj, n_tests, nullptr, 1, thr, tail, indptr,
                                                     indices, h0_tmp, t_obs, root, mass_root, 0, nullptr);
    CMC_CHECK_LAUNCH("cbpa_kernel<observed>");
    cbpa_label_kernel<<<1, 1024, 0, st>>>(root, mass_root, n_tests, labels,
                                           reinterpret_cast<long long*>(mass_fixed), mass_f64, n_clusters, rank);
    CMC_CHECK_LAUNCH("cbpa_label_kernel");
    return CMC_OK;
}

// Register-level DFT butterflies, twiddle application and Stockham plans shared by the K1 kernels.
#pragma once
#include "common.cuh"

namespace cmc {

// ---------------------------------------------------------------- small DFTs in registers
template <int R> __device__ __forceinline__ void dft(float2 (&v)[R]);

template <> __device__ __forceinline__ void dft<2>(float2 (&v)[2]) {
    float2 a = v[0];
    v[0] = cadd(a, v[1]);
    v[1] = csub(a, v[1]);
}
template <> __device__ __forceinline__ void dft<4>(float2 (&v)[4]) {
    float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    float2 t2 = cadd(v[1], v[3]), d = csub(v[1], v[3]);
    float2 t3 = make_float2(d.y, -d.x);  // -i * d
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
}
template <> __device__ __forceinline__ void dft<8>(float2 (&v)[8]) {
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    dft<4>(e);
    dft<4>(o);
    const float h = 0.70710678118654752f;
    o[1] = make_float2((o[1].x + o[1].y) * h, (o[1].y - o[1].x) * h);
    o[2] = make_float2(o[2].y, -o[2].x);
    o[3] = make_float2((o[3].y - o[3].x) * h, -(o[3].x + o[3].y) * h);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = cadd(e[k], o[k]);
        v[k + 4] = csub(e[k], o[k]);
    }
}
template <> __device__ __forceinline__ void dft<16>(float2 (&v)[16]) {
    float2 e[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        e[k] = v[2 * k];
        o[k] = v[2 * k + 1];
    }
    dft<8>(e);
    dft<8>(o);
    // W16^k = (cos(pi k / 8), -sin(pi k / 8))
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f;
    const float h = 0.70710678118654752f;
    o[1] = cmul(o[1], make_float2(c1, -s1));
    o[2] = make_float2((o[2].x + o[2].y) * h, (o[2].y - o[2].x) * h);
    o[3] = cmul(o[3], make_float2(s1, -c1));
    o[4] = make_float2(o[4].y, -o[4].x);
    o[5] = cmul(o[5], make_float2(-s1, -c1));
    o[6] = make_float2((o[6].y - o[6].x) * h, -(o[6].x + o[6].y) * h);
    o[7] = cmul(o[7], make_float2(-c1, -s1));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = cadd(e[k], o[k]);
        v[k + 8] = csub(e[k], o[k]);
    }
}

// twiddle powers w^1..w^(R-1) from a few table entries (<= 2 chained multiplies each)
template <int R>
__device__ __forceinline__ void apply_twiddles(float2 (&v)[R], const float2* __restrict__ tw, int base) {
    // tw index of w^r is r * base
    float2 w1 = __ldg(tw + base);
    v[1] = cmul(v[1], w1);
    if (R >= 4) {
        float2 w2 = __ldg(tw + 2 * base);
        v[2] = cmul(v[2], w2);
        float2 w3 = cmul(w1, w2);
        v[3] = cmul(v[3], w3);
        if (R >= 8) {
            float2 w4 = __ldg(tw + 4 * base);
            v[4] = cmul(v[4], w4);
            v[5] = cmul(v[5], cmul(w4, w1));
            v[6] = cmul(v[6], cmul(w4, w2));
            float2 w7 = cmul(w4, w3);
            v[7] = cmul(v[7], w7);
            if (R >= 16) {
                float2 w8 = __ldg(tw + 8 * base);
                v[8] = cmul(v[8], w8);
                v[9] = cmul(v[9], cmul(w8, w1));
                v[10] = cmul(v[10], cmul(w8, w2));
                v[11] = cmul(v[11], cmul(w8, w3));
                v[12] = cmul(v[12], cmul(w8, w4));
                v[13] = cmul(v[13], cmul(w8, cmul(w4, w1)));
                v[14] = cmul(v[14], cmul(w8, cmul(w4, w2)));
                v[15] = cmul(v[15], cmul(w8, w7));
            }
        }
    }
}

template <int M> struct Plan;
template <> struct Plan<64>   { static constexpr int R0 = 8,  R1 = 8,  R2 = 1;  };
template <> struct Plan<128>  { static constexpr int R0 = 16, R1 = 8,  R2 = 1;  };
template <> struct Plan<256>  { static constexpr int R0 = 16, R1 = 16, R2 = 1;  };
template <> struct Plan<512>  { static constexpr int R0 = 8,  R1 = 8,  R2 = 8;  };
template <> struct Plan<1024> { static constexpr int R0 = 16, R1 = 8,  R2 = 8;  };
template <> struct Plan<2048> { static constexpr int R0 = 16, R1 = 16, R2 = 8;  };
template <> struct Plan<4096> { static constexpr int R0 = 16, R1 = 16, R2 = 16; };


// Read-only table of float2 entries: global memory through the read-only path, or a copy a persistent kernel staged
// in shared memory (`s` = shared-window byte address of entry 0; loads then cost a shared-memory round trip instead
// of an L1 / L2 one).
template <bool SMEM>
struct Tab {
    const float2* g;
    uint32_t s;
    __device__ __forceinline__ float2 operator()(int i) const {
        if (SMEM) {
            float2 v;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(s + 8u * (uint32_t)i));
            return v;
        }
        return __ldg(g + i);
    }
};

// the same twiddles applied to two channels held by one thread (table loads and powers shared)
template <int R, class TW>
__device__ __forceinline__ void apply_twiddles2(float2 (&v)[R], float2 (&u)[R], const TW& tw, int base) {
    float2 w[R];
    w[1] = tw(base);
    if (R >= 4) {
        w[2] = tw(2 * base);
        w[3] = cmul(w[1], w[2]);
    }
    if (R >= 8) {
        w[4] = tw(4 * base);
        w[5] = cmul(w[4], w[1]);
        w[6] = cmul(w[4], w[2]);
        w[7] = cmul(w[4], w[3]);
    }
    if (R >= 16) {
        w[8] = tw(8 * base);
#pragma unroll
        for (int r = 1; r < 8; ++r) w[8 + r] = cmul(w[8], w[r]);
    }
#pragma unroll
    for (int r = 1; r < R; ++r) {
        v[r] = cmul(v[r], w[r]);
        u[r] = cmul(u[r], w[r]);
    }
}

}  // namespace cmc

// Workspace layout of the pooled-CSD operands and TMA tensor-map helpers shared by K2 / K3.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace cmc {

constexpr int kKBlock = 32;                     // floats per k-block = 128 bytes = one swizzle row
constexpr int kTileM = 128;
constexpr int kTileN = 64;

// Workspace of one pooled-CSD problem.  Operands are K-major (K = (segment, re/im) contiguous), UNSCALED TF32
// splits of the spectra; the normalisation by the auto-spectra happens in the GEMM epilogues.
//   A_hi / A_lo [F][MT*128][KP]  rows r < 64: X of channel mt*64 + r, rows 64 + r: i * X
//   B_hi / B_lo [F][NT*64][KP]   rows: Y
//   B_dbl / B_odd [F][NT*64][LB] shift-surrogate views of B_hi, built on demand: [Y | Y | 0..] and the same advanced
//                                by one complex element (TMA boxes must start 16-byte aligned);
//   B_dbl_lo / B_odd_lo          the same views of B_lo (second and third term of the 3xTF32 shift contraction)
struct CsdLayout {
    int L, F, Ne, Nm, MT, NT, KP, LB;
    int64_t a_elems, b_elems, bs_elems;          // floats per plane
    int64_t off_pxx, off_pyy, off_ahi, off_alo, off_bhi, off_blo, off_bdbl, off_bodd, off_bdbl_lo, off_bodd_lo, total;
};

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

inline CsdLayout csd_layout(int L, int F, int Ne, int Nm) {
    CsdLayout y;
    y.L = L; y.F = F; y.Ne = Ne; y.Nm = Nm;
    y.MT = (Ne + 63) / 64;
    y.NT = (Nm + 63) / 64;
    y.KP = (int)align_up(2 * (int64_t)L, kKBlock);
    y.LB = (int)align_up(2 * (int64_t)L + y.KP, kKBlock);
    y.a_elems = (int64_t)F * y.MT * kTileM * y.KP;
    y.b_elems = (int64_t)F * y.NT * kTileN * y.KP;
    y.bs_elems = (int64_t)F * y.NT * kTileN * y.LB;
    int64_t o = 0;
    y.off_pxx = o; o = align_up(o + (int64_t)F * Ne * 4, 1024);
    y.off_pyy = o; o = align_up(o + (int64_t)F * Nm * 4, 1024);
    y.off_ahi = o; o = align_up(o + y.a_elems * 4, 1024);
    y.off_alo = o; o = align_up(o + y.a_elems * 4, 1024);
    y.off_bhi = o; o = align_up(o + y.b_elems * 4, 1024);
    y.off_blo = o; o = align_up(o + y.b_elems * 4, 1024);
    y.off_bdbl = o; o = align_up(o + y.bs_elems * 4, 1024);
    y.off_bodd = o; o = align_up(o + y.bs_elems * 4, 1024);
    y.off_bdbl_lo = o; o = align_up(o + y.bs_elems * 4, 1024);
    y.off_bodd_lo = o; o = align_up(o + y.bs_elems * 4, 1024);
    y.total = o;
    return y;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int get_encode_fn(EncodeTiledFn* fn) {
    static EncodeTiledFn cached = nullptr;
    if (!cached) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        int rc = check_cuda(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q),
                            "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)");
        if (rc) return rc;
        if (q != cudaDriverEntryPointSuccess || !ptr) {
            set_error("cuTensorMapEncodeTiled not available in this driver");
            return CMC_ECUDA;
        }
        cached = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    *fn = cached;
    return CMC_OK;
}

// 2-D K-major operand map: dim0 = row_len floats (contiguous), dim1 = rows; box = 32 floats x box_rows
inline int make_operand_map(CUtensorMap* m, const float* base, int64_t row_len, int64_t rows, int box_rows) {
    EncodeTiledFn enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    cuuint64_t dims[2] = {(cuuint64_t)row_len, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_len * 4};
    cuuint32_t box[2] = {(cuuint32_t)kKBlock, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (row_len=%lld rows=%lld)", (int)r,
                  (long long)row_len, (long long)rows);
        return CMC_ECUDA;
    }
    return CMC_OK;
}


// generic 2-D K-major map: rows of `row_len` elements of `elem_bytes`, box = 128 bytes x box_rows
inline int make_kmajor_map(CUtensorMap* m, const void* base, CUtensorMapDataType dt, int elem_bytes, int64_t row_len,
                           int64_t rows, int box_rows) {
    EncodeTiledFn enc;
    int rc = get_encode_fn(&enc);
    if (rc) return rc;
    cuuint64_t dims[2] = {(cuuint64_t)row_len, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_len * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (row_len=%lld rows=%lld)", (int)r,
                  (long long)row_len, (long long)rows);
        return CMC_ECUDA;
    }
    return CMC_OK;
}

}  // namespace cmc

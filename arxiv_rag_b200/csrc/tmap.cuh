// Host-side TMA tensor-map construction. The driver symbol is fetched through the
// runtime (cudaGetDriverEntryPoint) so the library links against cudart only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace arb {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) !=
                cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// 2-D row-major bf16 matrix [rows, cols] with leading dimension ld (elements); box =
// [box_rows, 64 cols] with the 128-byte swizzle the UMMA descriptors expect. Out-of-bounds
// elements read as zero, so ragged M/N/K edges need no special casing in the main loop.
inline bool make_tmap_bf16_k64(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                               uint64_t ld, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Batched 16-bit matrix [batch, rows, cols] (row stride ld, batch stride rows*ld elements), box =
// [1, box_rows, 64 cols], 128-byte swizzle. Rows >= `rows` of a batch entry read as zero, so a
// sequence never sees its neighbour's tokens.
inline bool make_tmap_bf16_batched_k64(CUtensorMap* map, const void* base, uint64_t batch,
                                       uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[3] = {cols, rows, batch};
    cuuint64_t strides[2] = {ld * 2, rows * ld * 2};
    cuuint32_t box[3] = {64, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Output / residual tiles: row-major [rows, cols] with `elt_bytes`-wide elements; box =
// [128 rows, 128 bytes of columns], 128-byte swizzle (the epilogue writes its staging tile with
// the same XOR pattern). Out-of-bounds rows/columns are clipped on store and zero on load.
inline bool make_tmap_rows128(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                              uint64_t ld, int elt_bytes) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * static_cast<uint64_t>(elt_bytes)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elt_bytes), 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                     2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace arb

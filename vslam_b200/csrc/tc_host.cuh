// tc_host.cuh — host side of the tensor path: TMA descriptor (CUtensorMap) construction.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace vb {
namespace tc {

// cuTensorMapEncodeTiled is a driver-API symbol. It is resolved through the runtime at first use instead of being
// linked, so libvslam_b200.so loads (and its exports can be listed) on a machine without libcuda.so.1.
typedef CUresult (*tensor_map_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                         const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline tensor_map_encode_fn tensor_map_encode() {
    static tensor_map_encode_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tensor_map_encode_fn>(p);
    }
    return fn;
}

// Row-major [rows][inner] matrix of `elem_bytes`-wide elements, box = box_inner x box_rows elements with the
// 128-byte swizzle (box_inner * elem_bytes must be 128). Out-of-range rows read as zero.
static inline int make_map_2d(CUtensorMap *m, CUtensorMapDataType dtype, uint32_t elem_bytes, const void *base, uint64_t inner,
                              uint64_t rows, uint32_t box_inner, uint32_t box_rows) {
    const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)(inner * elem_bytes)};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    tensor_map_encode_fn enc = tensor_map_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return VB_ERR_CUDA;
    }
    const CUresult r = enc(m, dtype, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled -> CUresult %d", (int)r);
        return VB_ERR_CUDA;
    }
    return VB_OK;
}

}  // namespace tc
}  // namespace vb

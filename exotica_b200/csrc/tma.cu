// tma.cu -- host side of the TMA output path: tensor-map encoding through the driver entry point.
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include "kernels.cuh"

namespace exo
{
namespace
{
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried)
    {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}
}  // namespace

static_assert(sizeof(TmaDesc) == sizeof(CUtensorMap), "TmaDesc must be a CUtensorMap");

bool tma_encode_2d(TmaDesc& out, const void* base, uint64_t inner, uint64_t rows, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (pitch_bytes & 15u) || box_inner == 0 || box_rows == 0 || box_inner > 256 || box_rows > 256) return false;
    if (inner == 0 || rows == 0 || inner > (1ull << 31) || rows > (1ull << 31)) return false;
    CUtensorMap m;
    const cuuint64_t dims[2] = {inner, rows};
    const cuuint64_t strides[1] = {pitch_bytes};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t es[2] = {1, 1};
    const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    memcpy(&out, &m, sizeof(m));
    return true;
}

bool tma_encode_3d(TmaDesc& out, const void* base, uint64_t inner, uint64_t d1, uint64_t d2, uint64_t pitch1_bytes, uint64_t pitch2_bytes, uint32_t box_inner,
                   uint32_t b1, uint32_t b2)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (pitch1_bytes & 15u) || (pitch2_bytes & 15u)) return false;
    if (box_inner == 0 || b1 == 0 || b2 == 0 || box_inner > 256 || b1 > 256 || b2 > 256) return false;
    if (inner == 0 || d1 == 0 || d2 == 0 || inner > (1ull << 31) || d1 > (1ull << 31) || d2 > (1ull << 31)) return false;
    CUtensorMap m;
    const cuuint64_t dims[3] = {inner, d1, d2};
    const cuuint64_t strides[2] = {pitch1_bytes, pitch2_bytes};
    const cuuint32_t box[3] = {box_inner, b1, b2};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    memcpy(&out, &m, sizeof(m));
    return true;
}

}  // namespace exo

// Host side of the TMA path: cuTensorMapEncodeTiled is fetched through the runtime's driver entry point query, so
// libexb.so needs no link-time dependency on libcuda.
#include "exb_tma.h"

#include <mutex>

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_once;
}  // namespace

int exb_tma_encode(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const char** err) {
  std::call_once(g_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = (EncodeTiledFn)fn;
  });
  if (!g_encode) {
    *err = "cuTensorMapEncodeTiled is not available from this driver";
    return -1;
  }
  if (rank < 2 || rank > 5) {
    *err = "tensor map rank must be 2..5";
    return -1;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i] * (i == 0 ? 2 : 1);  // complex -> 2 floats
    bx[i] = box[i] * (i == 0 ? 2 : 1);
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    *err = "cuTensorMapEncodeTiled failed (alignment / extent constraints)";
    return -1;
  }
  return 0;
}

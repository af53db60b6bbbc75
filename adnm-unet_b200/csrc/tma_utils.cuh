// TMA (cp.async.bulk.tensor) building blocks shared by the tcgen05 GEMM (tcgemm.cuh) and the Haar WTConv2d halo tiles
// (wtconv.cu): tensor-map encoding on the host, bulk tensor loads completing on an mbarrier on the device.
#pragma once
#include <cuda.h>      // CUtensorMap (types only: cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint)

#include "adn_common.cuh"
#include "sm100_utils.cuh"

namespace adn {
namespace tma {
using namespace adn::sm100;

// one TMA box: coordinates (c0 innermost, c1, c2 = batch) of a rank-3 tensor map -> shared memory, bytes counted on `bar`
__device__ __forceinline__ void tma_load_3d(uint32_t sdst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(sdst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

}  // namespace tma
}  // namespace adn

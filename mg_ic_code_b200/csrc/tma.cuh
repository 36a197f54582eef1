// tma.cuh -- what the plane-streaming kernels (gsrb_fused.cu, restrict_tma.cu) share: mbarrier / cp.async.bulk.tensor wrappers
// (device), tensor-map encoding through the driver entry point and the z-chunk planner (host).
#ifndef MGIC_TMA_CUH
#define MGIC_TMA_CUH

#include <cuda.h>

#include <cstdint>

#include "mgic_internal.h"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded spin: a lost TMA completion traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }


// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

inline int make_tmap(CUtensorMap *m, const double *base, int nx, int ny, int nplanes, int boxx, int boxy) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { mgic_set_error("cuTensorMapEncodeTiled is not available from the driver"); return MGIC_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nplanes};
  cuuint64_t strides[2] = {(cuuint64_t)nx * 8, (cuuint64_t)nx * ny * 8};
  cuuint32_t box[3] = {(cuuint32_t)boxx, (cuuint32_t)boxy, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { mgic_set_error("cuTensorMapEncodeTiled failed (%d) for %dx%dx%d box %dx%d", (int)r, nx, ny, nplanes, boxx, boxy); return MGIC_ERR_CUDA; }
  return MGIC_OK;
}

struct Plan { int nch, zchunk; };

// number of z chunks: fill whole waves of resident CTAs while keeping the two redundant planes per chunk cheap
inline Plan plan_chunks(int tiles, int nz, int resident) {
  Plan best = {1, nz};
  double bestScore = -1.0;
  for (int nch = 1; nch <= nz; nch++) {
    const int zc = (nz + nch - 1) / nch;
    if (zc < 8 && nch > 1) break;
    const int nchEff = (nz + zc - 1) / zc;
    const long long total = (long long)tiles * nchEff;
    const long long waves = (total + resident - 1) / resident;
    const double fill = (double)total / (double)(waves * resident);
    const double score = fill * zc / (zc + 2.0);
    if (score > bestScore + 1e-9) { bestScore = score; best = {nchEff, zc}; }
  }
  return best;
}


}  // namespace

#endif

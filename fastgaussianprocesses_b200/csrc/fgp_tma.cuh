// TMA building blocks shared by the transform and posterior kernels: mbarrier completion tracking, 1-D bulk copies
// (cp.async.bulk, SASS UBLKCP) and 2-D tiled tensor loads (cp.async.bulk.tensor, SASS UTMALDG) into shared memory.
#pragma once
#include <cuda.h>  // CUtensorMap and its enums only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint (no -lcuda)
#include "fgp_common.cuh"

namespace fgp {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// spin on the phase parity; traps (a launch error, not a hang) if the copies never land
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const uint32_t addr = smem_u32(bar);
  for (unsigned spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (spins > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 2-D tiled TMA load (cp.async.bulk.tensor, SASS UTMALDG): box (c0.., c1..) of the tensor described by `tmap` -> dense shared memory
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst_smem)),
               "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

// row-major (rows, cols) float64 matrix at `base` as a tiled tensor map with boxes of (box_rows, box_cols)
static inline int make_tmap_2d(const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows, CUtensorMap* out) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      cudaGetLastError();
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return FGP_ECUDA;
    }
    encode = (encode_fn)fn;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * sizeof(double)};  // bytes between rows (dimension 0 is contiguous)
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code %d (cols=%llu rows=%llu box=%ux%u)", (int)r, (unsigned long long)cols, (unsigned long long)rows, box_cols, box_rows);
    return FGP_ECUDA;
  }
  return FGP_OK;
}


}  // namespace fgp

// TMA (cp.async.bulk.tensor) store helpers for sm_100a: a 3-D tensor map over a [planes, H, W] fp32 tensor and the
// shared -> global tile store issued by one elected thread.  The driver entry point is fetched through the runtime
// (cudaGetDriverEntryPoint), so libfovea_b200.so does not link libcuda.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace fovea {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// base [planes][H][W] fp32 (W % 4 == 0, 16-byte aligned); box = box_w x box_h pixels of one plane, no swizzle:
// the shared-memory tile is plain row-major [box_h][box_w].  Stores outside the tensor are clipped by the hardware.
inline int make_plane_store_map(CUtensorMap* map, float* base, long long planes, int H, int W, int box_w, int box_h) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return FOVEA_ERR_CUDA;
  }
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(planes)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(W) * 4, static_cast<cuuint64_t>(H) * W * 4};
  const cuuint32_t box[3] = {static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) for a [%lld,%d,%d] tensor", static_cast<int>(r), planes, H, W);
    return FOVEA_ERR_CUDA;
  }
  return FOVEA_OK;
}

// the same map for LOADS of box_w x box_h pixel tiles; elements outside the tensor arrive as zeros (F.grid_sample's
// padding_mode='zeros' for free)
inline int make_plane_load_map(CUtensorMap* map, const float* base, long long planes, int H, int W, int box_w, int box_h) {
  return make_plane_store_map(map, const_cast<float*>(base), planes, H, W, box_w, box_h);
}

__device__ __forceinline__ unsigned smem_addr(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}

// one thread: store the [box_h][box_w] tile at `smem` to plane z, rows y.., columns x.. of the mapped tensor
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* map, const void* smem, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
               :: "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_addr(smem))
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() {  // at most N committed groups may still be READING shared memory
  asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
// ---- loads: global -> shared tile, completion signalled on an mbarrier (transaction bytes)
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      :: "r"(smem_addr(bar)), "r"(parity) : "memory");
}
// one thread: load the [box_h][box_w] tile whose first element is (x, y) of plane z into `smem` (x, y may be negative or
// reach past the tensor: those elements are zero-filled).  x * 4 bytes must be a multiple of 16: a misaligned inner
// coordinate raises "illegal instruction" (measured with tools/probes/tma_load_test.cu: x = -1 faults, x = -4 does not)
__device__ __forceinline__ void tma_load_tile(const CUtensorMap* map, void* smem, unsigned long long* bar, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :: "r"(smem_addr(smem)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(smem_addr(bar)), "r"(x), "r"(y), "r"(z)
               : "memory");
}

// generic-proxy writes to shared memory (st.shared) become visible to the async proxy (TMA)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace fovea

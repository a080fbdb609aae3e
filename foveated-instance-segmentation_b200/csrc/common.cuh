// Shared helpers for libfovea_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "fovea_b200.h"

namespace fovea {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return FOVEA_ERR_CUDA;
  }
  return FOVEA_OK;
}

#define FOVEA_REQUIRE(cond, ...)   \
  do {                             \
    if (!(cond)) {                 \
      fovea::set_error(__VA_ARGS__); \
      return FOVEA_ERR_ARG;        \
    }                              \
  } while (0)

#define FOVEA_CUDA(call)                                                     \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) {                                                \
      fovea::set_error("%s failed: %s", #call, cudaGetErrorString(e__));     \
      return FOVEA_ERR_CUDA;                                                 \
    }                                                                        \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Source index of a padded coordinate t (already shifted by -R, i.e. in source-pixel units), -1 = zero tap.
__device__ __forceinline__ int pad_map(int t, int n, int mode) {
  if (mode == FOVEA_PAD_NONE) return t;  // caller indexes the padded map directly
  if (t >= 0 && t < n) return t;
  if (mode == FOVEA_PAD_REPLICATION) return t < 0 ? 0 : n - 1;
  if (mode == FOVEA_PAD_REFLECT) {  // F.pad 'reflect' (no edge repeat); pad < n is checked on the host
    return t < 0 ? -t : 2 * (n - 1) - t;
  }
  return -1;  // zero
}

// aten area_pixel_compute_source_index (align_corners=False, linear): max(scale*(dst+0.5)-0.5, 0)
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0,
                                             float& l1) {
  float real = fmaxf(scale * (static_cast<float>(dst) + 0.5f) - 0.5f, 0.f);
  i0 = static_cast<int>(real);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = fminf(fmaxf(real - static_cast<float>(i0), 0.f), 1.f);
  l0 = 1.f - l1;
}

}  // namespace fovea

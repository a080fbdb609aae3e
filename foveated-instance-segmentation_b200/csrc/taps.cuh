// Bilinear tap set of aten's grid_sampler_2d (bilinear, zeros padding, align_corners=False), bit-compatible
// with torch's CPU kernel: ix = fma(x+1, W/2, -0.5), weights from the distances to the opposite corner.
#pragma once
#include "common.cuh"

namespace fovea {

struct Taps {
  int x0, y0;
  float nw, ne, sw, se;
  bool ok_nw, ok_ne, ok_sw, ok_se;
  float ix, iy;
};

__device__ __forceinline__ Taps make_taps(float gx, float gy, int H, int W) {
  Taps t;
  t.ix = fmaf(gx + 1.f, 0.5f * static_cast<float>(W), -0.5f);
  t.iy = fmaf(gy + 1.f, 0.5f * static_cast<float>(H), -0.5f);
  const float fx = floorf(t.ix), fy = floorf(t.iy);
  const float tx = t.ix - fx, ty = t.iy - fy;
  const float ex = 1.f - tx, ey = 1.f - ty;
  t.nw = ey * ex;
  t.ne = ey * tx;
  t.sw = ty * ex;
  t.se = ty * tx;
  // NaN / huge coordinates: the casts saturate and every tap fails the bounds test
  t.x0 = static_cast<int>(fx);
  t.y0 = static_cast<int>(fy);
  const bool xl = t.x0 >= 0 && t.x0 < W, xr = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  const bool yt = t.y0 >= 0 && t.y0 < H, yb = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  const bool finite = (t.ix == t.ix) && (t.iy == t.iy);
  t.ok_nw = finite && xl && yt;
  t.ok_ne = finite && xr && yt;
  t.ok_sw = finite && xl && yb;
  t.ok_se = finite && xr && yb;
  return t;
}

}  // namespace fovea

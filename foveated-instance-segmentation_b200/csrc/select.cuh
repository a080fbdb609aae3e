// Which filled pixels are interpolation sites: getPixelsForInterp / getPixelsForInterp_NB of fillMissingValues_tensor
// (models/models.py:169-242), evaluated on the A7 winner map (`winner < 0` = the reference's NaN pattern).
#pragma once
#include "common.cuh"

namespace fovea {

struct SelectParams {
  int h, w, H, W, cap;
  int scaled;           // 1: dilation runs on the nearest-downscaled mask (max(C,H,W) > 512)
  int hs, ws;           // downscaled size
  float dn_y, dn_x;     // H/hs, W/ws  (nearest downscale source scale)
  float up_y, up_x;     // hs/H, ws/W  (nearest upscale source scale)
};

// models/models.py:183-187 (and :222-226): dr = max(C,H,W)/512 in Python floats (double), sizes through int()
inline int make_select_params(SelectParams& p, int h, int w, int H, int W, int nchan, int cap, const char* who) {
  p.h = h; p.w = w; p.H = H; p.W = W; p.cap = cap;
  const int mx = nchan > H ? (nchan > W ? nchan : W) : (H > W ? H : W);
  p.scaled = mx > 512;
  p.hs = H; p.ws = W; p.dn_y = p.dn_x = p.up_y = p.up_x = 1.f;
  if (p.scaled) {
    const double dr = static_cast<double>(mx) / 512.0;
    p.hs = static_cast<int>(static_cast<double>(H) / dr);
    p.ws = static_cast<int>(static_cast<double>(W) / dr);
    FOVEA_REQUIRE(p.hs > 0 && p.ws > 0, "%s: downscaled mask is empty (%dx%d)", who, p.hs, p.ws);
    p.dn_y = static_cast<float>(H) / static_cast<float>(p.hs);
    p.dn_x = static_cast<float>(W) / static_cast<float>(p.ws);
    p.up_y = static_cast<float>(p.hs) / static_cast<float>(H);
    p.up_x = static_cast<float>(p.ws) / static_cast<float>(W);
  }
  return FOVEA_OK;
}

// "Is pixel (y,x) unfilled?" -- answered from the dense A7 winner map ...
struct DenseWinners {
  const int32_t* win;
  int W;
  __device__ __forceinline__ bool unfilled(int y, int x) const { return win[static_cast<size_t>(y) * W + x] < 0; }
};
// ... or from the frame's node targets themselves, sorted in shared memory as (row << 16 | column) << 32 | node + 1
// (entries with a zero low word are the forced corners' place holders, not nodes): a binary search, no dense map.
struct SortedTargets {
  const unsigned long long* keys;
  int n;
  __device__ __forceinline__ bool unfilled(int y, int x) const {
    const unsigned long long want = (static_cast<unsigned long long>((y << 16) | x) << 32) | 1ull;
    int lo = 0, hi = n;   // first entry >= want
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (keys[mid] < want) lo = mid + 1; else hi = mid;
    }
    return lo >= n || static_cast<unsigned>(keys[lo] >> 32) != static_cast<unsigned>((y << 16) | x);
  }
};

template <class Winners>
__device__ __forceinline__ bool invalid_at(const Winners& win, int H, int W, int y, int x) {
  return y >= 0 && y < H && x >= 0 && x < W && win.unfilled(y, x);
}

// dilated(y,x): does the structuring element around (y,x) touch an invalid pixel?
//   kVerticalOnly = false: the 3x3 cross of getPixelsForInterp ('tri', F.conv2d over [1,C,H,W], models.py:180-197);
//   kVerticalOnly = true : getPixelsForInterp_NB ('nearest'/'BI', :213-239) hands the [C,H,W] array to cv2.dilate, which
//                          reads it as rows=C, cols=H, channels=W -- the cross then spans the CLASS and ROW axes; the
//                          NaN pattern is identical in every class, so what is left is (y-1, y, y+1) in the same column.
template <bool kVerticalOnly, class Winners>
__device__ bool dilation_covers(const Winners& win, const SelectParams& p, int y, int x) {
  if (!p.scaled) {
    bool c = invalid_at(win, p.H, p.W, y - 1, x) || invalid_at(win, p.H, p.W, y + 1, x) || invalid_at(win, p.H, p.W, y, x);
    if (!kVerticalOnly) c = c || invalid_at(win, p.H, p.W, y, x - 1) || invalid_at(win, p.H, p.W, y, x + 1);
    return c;
  }
  // nearest upscale: dilated[y][x] = dilated_s[ys][xs]
  const int ys = min(static_cast<int>(floorf(static_cast<float>(y) * p.up_y)), p.hs - 1);
  const int xs = min(static_cast<int>(floorf(static_cast<float>(x) * p.up_x)), p.ws - 1);
  const int dy[5] = {0, -1, 1, 0, 0}, dx[5] = {0, 0, 0, -1, 1};
#pragma unroll
  for (int k = 0; k < (kVerticalOnly ? 3 : 5); ++k) {
    const int yy = ys + dy[k], xx = xs + dx[k];
    if (yy < 0 || yy >= p.hs || xx < 0 || xx >= p.ws) continue;  // zero padding / BORDER_CONSTANT 0
    // nearest downscale: scaled[yy][xx] = invalid[min(floor(yy*H/hs), H-1)][...]
    const int sy = min(static_cast<int>(floorf(static_cast<float>(yy) * p.dn_y)), p.H - 1);
    const int sx = min(static_cast<int>(floorf(static_cast<float>(xx) * p.dn_x)), p.W - 1);
    if (win.unfilled(sy, sx)) return true;
  }
  return false;
}

}  // namespace fovea

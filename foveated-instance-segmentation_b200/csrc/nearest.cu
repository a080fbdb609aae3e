// Stage 3, rev_deform_interp = 'nearest' (the mode config/deform.yaml:17 ships): every unfilled pixel takes the value of
// the NEAREST interpolation site.
//
// Reference: fillMissingValues_tensor(..., 'nearest'), models/models.py:213-250, 259-272: sites = getPixelsForInterp_NB
// (filled pixels with an unfilled pixel directly above / below, at <= 512 px directly, else on the nearest-downscaled
// mask; no forced corners), then scipy.interpolate.NearestNDInterpolator over the 3-D (class,row,col) voxels on the HOST
// (113 s per 1024^2 frame).  The NaN pattern is the same in every class, so the nearest voxel is
// always in the query's own class plane: the operation is a 2-D nearest-site (Voronoi) labelling, done here exactly with
// integer distances in two passes -- no triangulation:
//   1. columns: g[y][x] = signed row offset from (y,x) to the nearest site of column x           (one thread per column)
//   2. rows   : pixel (y,x) scans columns x-k, x+k for k = 0,1,2,... and keeps min k^2 + g[y][x+-k]^2; it stops as soon
//               as k^2 >= best -- O(distance to the nearest site) steps, every load coalesced      (one thread per pixel)
// The result is the same per-pixel source map `loc` the 'tri' mode produces (every entry a direct table row), so
// fovea_inverse_fill streams the scores unchanged.  Equidistant sites: the reference's KD-tree returns whichever it
// meets first; here the smaller |dx| wins, then the left one, then (within a column) the upper one.
#include "common.cuh"
#include "select.cuh"

namespace fovea {

constexpr short kNoSite = 32767;

__global__ void __launch_bounds__(128)
nearest_columns_kernel(const int32_t* __restrict__ winner, short* __restrict__ g, SelectParams p) {
  const int b = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= p.W) return;
  const int32_t* win = winner + static_cast<size_t>(b) * p.H * p.W;
  short* gb = g + static_cast<size_t>(b) * p.H * p.W;
  // Both sweeps read 8 rows ahead (independent loads) before the sequential part, so the column walk is not one
  // exposed memory latency per pixel.
  constexpr int U = 8;
  // sweep down: offset to the nearest site at or above
  int last = -1;
  for (int y0 = 0; y0 < p.H; y0 += U) {
    int wv[U];
#pragma unroll
    for (int k = 0; k < U; ++k) wv[k] = (y0 + k < p.H) ? __ldg(win + static_cast<size_t>(y0 + k) * p.W + x) : -1;
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int y = y0 + k;
      if (y >= p.H) break;
      if (wv[k] >= 0 && dilation_covers<true>(win, p, y, x)) last = y;
      gb[static_cast<size_t>(y) * p.W + x] = last >= 0 ? static_cast<short>(last - y) : kNoSite;
    }
  }
  // sweep up: a strictly nearer site below replaces it (ties keep the upper one)
  int next = -1;
  for (int y0 = p.H - 1; y0 >= 0; y0 -= U) {
    short uv[U];
#pragma unroll
    for (int k = 0; k < U; ++k) uv[k] = (y0 - k >= 0) ? gb[static_cast<size_t>(y0 - k) * p.W + x] : kNoSite;
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int y = y0 - k;
      if (y < 0) break;
      const short up = uv[k];
      if (up == 0) { next = y; continue; }
      if (next >= 0 && (up == kNoSite || next - y < -up)) gb[static_cast<size_t>(y) * p.W + x] = static_cast<short>(next - y);
    }
  }
}

__global__ void __launch_bounds__(256)
nearest_rows_kernel(const int32_t* __restrict__ winner, const short* __restrict__ g, uint16_t* __restrict__ loc,
                    int hw, int H, int W) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const size_t row = (static_cast<size_t>(b) * H + y) * W;
  const int n = winner[row + x];
  if (n >= 0) {  // a filled pixel keeps its own value
    loc[row + x] = static_cast<uint16_t>(0x8000 | n);
    return;
  }
  const short* gr = g + row;
  // squared distances fit 32 bits: k, |dy| < 32767  =>  k^2 + dy^2 < 2^31
  unsigned best = 0xffffffffu;
  int bx = -1, bdy = 0;
  for (int k = 0; k < W; ++k) {
    const unsigned k2 = static_cast<unsigned>(k) * static_cast<unsigned>(k);
    if (k2 >= best) break;
    const int xl = x - k, xr = x + k;
    if (xl < 0 && xr >= W) break;
    const int dl = xl >= 0 ? gr[xl] : kNoSite;
    const int dr = (k > 0 && xr < W) ? gr[xr] : kNoSite;
    if (dl != kNoSite) {
      const unsigned d = k2 + static_cast<unsigned>(dl * dl);
      if (d < best) { best = d; bx = xl; bdy = dl; }
    }
    if (dr != kNoSite) {
      const unsigned d = k2 + static_cast<unsigned>(dr * dr);
      if (d < best) { best = d; bx = xr; bdy = dr; }
    }
  }
  int out = hw;  // no site anywhere: the NaN row of the value table
  if (bx >= 0) out = winner[(static_cast<size_t>(b) * H + (y + bdy)) * W + bx];
  loc[row + x] = static_cast<uint16_t>(0x8000 | out);   // 16-bit source map: bit 15 = direct table row
}

}  // namespace fovea

using namespace fovea;

extern "C" int64_t fovea_nearest_workspace_bytes(int B, int H, int W) {
  return static_cast<int64_t>(B) * H * W * static_cast<int64_t>(sizeof(short));
}

extern "C" int fovea_nearest_locate(const int32_t* winner, int B, int h, int w, int H, int W, int nchan,
                                    void* workspace, uint16_t* loc, fovea_stream_t stream) {
  FOVEA_REQUIRE(winner && workspace && loc, "fovea_nearest_locate: null pointer");
  FOVEA_REQUIRE(B > 0 && h > 0 && w > 0 && H > 1 && W > 1 && nchan > 0, "fovea_nearest_locate: bad sizes");
  FOVEA_REQUIRE(H < 32767 && W < 32767 && B <= 65535 && H <= 65535, "fovea_nearest_locate: canvas or batch too large");
  FOVEA_REQUIRE(h * w < 32767, "fovea_nearest_locate: table rows must fit 15 bits (h*w=%d)", h * w);
  SelectParams p;
  if (int rc = make_select_params(p, h, w, H, W, nchan, 0, "fovea_nearest_locate")) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  short* g = static_cast<short*>(workspace);
  nearest_columns_kernel<<<dim3(ceil_div(W, 128), B), 128, 0, s>>>(winner, g, p);
  if (int rc = check_launch("fovea_nearest_locate (columns)")) return rc;
  nearest_rows_kernel<<<dim3(ceil_div(W, 256), H, B), 256, 0, s>>>(winner, g, loc, h * w, H, W);
  return check_launch("fovea_nearest_locate (rows)");
}

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
// meets first; here the leftmost column wins, then (within a column) the upper site.
#include <stdlib.h>

#include "common.cuh"
#include "select.cuh"

namespace fovea {

constexpr short kNoSite = 32767;

// kAllSites: every filled pixel is a site (DynamicFocus deformed_unsampler: an exact Euclidean distance transform of the
// scattered labels, nn_B0_deformed_sampler.py:139-149) instead of the reference's dilation rule.
template <bool kAllSites>
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
      if (wv[k] >= 0 && (kAllSites || dilation_covers<true>(win, p, y, x))) last = y;
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

// Row pass, exact and O(W log W) per row regardless of how sparse the sites are: for one image row the cost
// c(x, x') = (x - x')^2 + g[y][x']^2 is a Monge array, so its leftmost row minima opt(x) are non-decreasing in x.
// One warp resolves one row by divide and conquer over x: position 0 first, then the odd multiples of the stride
// s = P/2, P/4, ..., 1; position x only scans candidates between the optima of its already-resolved neighbours
// x - s and x + s.  While a level has fewer positions than lanes the warp scans each range cooperatively (strided
// candidates + a shuffle arg-min); afterwards every lane resolves its own positions.  The per-level candidate count is
// <= W + (number of positions), ~11 W evaluations per row in total.
constexpr int kDcMaxWarps = 8;

__device__ __forceinline__ void argmin_merge(unsigned& d, int& i, unsigned d2, int i2) {
  if (d2 < d || (d2 == d && i2 < i)) { d = d2; i = i2; }  // leftmost among equal distances
}

__global__ void __launch_bounds__(kDcMaxWarps * 32)
nearest_rows_dc_kernel(const int32_t* __restrict__ winner, const short* __restrict__ g, uint16_t* __restrict__ loc, int hw,
                       int H, int W, int P /* smallest power of two >= W */) {
  extern __shared__ unsigned short dc_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, y = blockIdx.x * (blockDim.x >> 5) + warp;
  if (y >= H) return;  // warp-uniform
  unsigned short* f = dc_smem + static_cast<size_t>(warp) * 2 * W;   // [W] |g| (0xffff = no site in the column)
  unsigned short* opt = f + W;                                        // [W] leftmost nearest column
  const size_t row = (static_cast<size_t>(b) * H + y) * W;
  bool any = false;
  for (int x = lane; x < W; x += 32) {
    const int dy = g[row + x];
    f[x] = dy == kNoSite ? 0xffffu : static_cast<unsigned short>(dy < 0 ? -dy : dy);
    any |= dy != kNoSite;
  }
  any = __any_sync(0xffffffffu, any);
  __syncwarp();
  if (!any) {  // no site in this image: the NaN row of the value table (filled pixels keep their own value)
    for (int x = lane; x < W; x += 32) {
      const int n = winner[row + x];
      loc[row + x] = static_cast<uint16_t>(0x8000 | (n >= 0 ? n : hw));
    }
    return;
  }
  auto cost = [&](int x, int c) {
    const unsigned fc = f[c];
    const int dx = x - c;
    return fc == 0xffffu ? 0xffffffffu : fc * fc + static_cast<unsigned>(dx * dx);   // < 2^31, see nearest_rows_kernel
  };
  // cooperative resolution of one position: all lanes scan [lo, hi] strided, then arg-min across the warp
  auto resolve_coop = [&](int x, int lo, int hi) {
    unsigned d = 0xffffffffu;
    int i = lo;
    for (int c = lo + lane; c <= hi; c += 32) argmin_merge(d, i, cost(x, c), c);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned d2 = __shfl_xor_sync(0xffffffffu, d, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
      argmin_merge(d, i, d2, i2);
    }
    if (lane == 0) opt[x] = static_cast<unsigned short>(i);
  };
  resolve_coop(0, 0, W - 1);
  __syncwarp();
  for (int s = P >> 1; s >= 1; s >>= 1) {
    const int npos = P / (2 * s);  // positions s, 3s, 5s, ... (those < W)
    if (npos < 32) {
      for (int k = 0; k < npos; ++k) {
        const int x = (2 * k + 1) * s;
        if (x >= W) break;
        const int lo = opt[x - s], hi = x + s < W ? opt[x + s] : W - 1;
        resolve_coop(x, lo, hi);
      }
    } else {
      for (int k = lane; k < npos; k += 32) {
        const int x = (2 * k + 1) * s;
        if (x >= W) break;
        const int lo = opt[x - s], hi = x + s < W ? opt[x + s] : W - 1;
        unsigned d = 0xffffffffu;
        int i = lo;
        for (int c = lo; c <= hi; ++c) {
          const unsigned dc = cost(x, c);
          if (dc < d) { d = dc; i = c; }
        }
        opt[x] = static_cast<unsigned short>(i);
      }
    }
    __syncwarp();
  }
  for (int x = lane; x < W; x += 32) {
    int n = winner[row + x];
    if (n < 0) {
      const int c = opt[x];
      n = winner[(static_cast<size_t>(b) * H + (y + g[row + c])) * W + c];
    }
    loc[row + x] = static_cast<uint16_t>(0x8000 | n);
  }
}

}  // namespace fovea

using namespace fovea;

extern "C" int64_t fovea_nearest_workspace_bytes(int B, int H, int W) {
  return static_cast<int64_t>(B) * H * W * static_cast<int64_t>(sizeof(short));
}

static int nearest_locate_impl(const int32_t* winner, int B, int h, int w, int H, int W, int nchan, bool all_sites,
                               void* workspace, uint16_t* loc, fovea_stream_t stream, const char* who) {
  FOVEA_REQUIRE(winner && workspace && loc, "%s: null pointer", who);
  FOVEA_REQUIRE(B > 0 && h > 0 && w > 0 && H > 1 && W > 1 && nchan > 0, "%s: bad sizes", who);
  FOVEA_REQUIRE(H < 32767 && W < 32767 && B <= 65535 && H <= 65535, "%s: canvas or batch too large", who);
  FOVEA_REQUIRE(h * w < 32767, "%s: table rows must fit 15 bits (h*w=%d)", who, h * w);
  SelectParams p;
  if (int rc = make_select_params(p, h, w, H, W, nchan, 0, who)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  short* g = static_cast<short*>(workspace);
  if (all_sites)
    nearest_columns_kernel<true><<<dim3(ceil_div(W, 128), B), 128, 0, s>>>(winner, g, p);
  else
    nearest_columns_kernel<false><<<dim3(ceil_div(W, 128), B), 128, 0, s>>>(winner, g, p);
  if (int rc = check_launch("fovea_nearest_locate (columns)")) return rc;
  // rows: the divide-and-conquer envelope keeps a row in shared memory; rows too long for that (W > ~14000) fall back to
  // the outward scan (exact too, but O(distance to the nearest site) per pixel)
  // (4 bytes of shared memory per pixel of a row and warp; long rows run with fewer warps per CTA so that several CTAs
  // still fit an SM)
  const int nw = W > 2048 ? 4 : kDcMaxWarps;
  const size_t smem = static_cast<size_t>(nw) * 2 * W * sizeof(unsigned short);
  static const bool force_scan = [] { const char* e = getenv("FOVEA_NEAREST_SCAN"); return e && e[0] == '1'; }();
  if (smem <= 227 * 1024 && !force_scan) {
    int P = 1;
    while (P < W) P <<= 1;
    FOVEA_CUDA(cudaFuncSetAttribute(nearest_rows_dc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    nearest_rows_dc_kernel<<<dim3(ceil_div(H, nw), B), nw * 32, smem, s>>>(winner, g, loc, h * w, H, W, P);
  } else {
    nearest_rows_kernel<<<dim3(ceil_div(W, 256), H, B), 256, 0, s>>>(winner, g, loc, h * w, H, W);
  }
  return check_launch("fovea_nearest_locate (rows)");
}

extern "C" int fovea_nearest_locate(const int32_t* winner, int B, int h, int w, int H, int W, int nchan,
                                    void* workspace, uint16_t* loc, fovea_stream_t stream) {
  return nearest_locate_impl(winner, B, h, w, H, W, nchan, false, workspace, loc, stream, "fovea_nearest_locate");
}

extern "C" int fovea_nearest_locate_all(const int32_t* winner, int B, int h, int w, int H, int W, void* workspace,
                                        uint16_t* loc, fovea_stream_t stream) {
  return nearest_locate_impl(winner, B, h, w, H, W, 1, true, workspace, loc, stream, "fovea_nearest_locate_all");
}

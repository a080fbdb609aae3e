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
      if (wv[k] >= 0 && (kAllSites || dilation_covers<true>(DenseWinners{win, p.W}, p, y, x))) last = y;
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


// ---------------------------------------------------------------------------------------------------------------------
// Tiled exact nearest-site labelling (chosen for sparse sites, see nearest_locate_impl): at most h*w (~6.4 k) sites per frame, so almost every 32 x 32 pixel
// tile can only be claimed by a handful of them.
//   1. nearest_sites_kernel   : one CTA per tile scans its pixels of the winner map, applies the site rule and appends the
//                               tile's sites (packed position + node) to the frame's site list: one contiguous BUCKET per
//                               tile (one atomicAdd per CTA reserves the range).
//   2. nearest_tiles_kernel   : one CTA per tile.  d0 = the distance from the tile centre c to SOME site s0 (the nearest one
//                               found in a growing square of buckets around the tile).  With r = the tile's half diagonal,
//                               a site that is nearest to ANY pixel p of the tile lies within d0 + 2r of c
//                               (|s-p| <= |s0-p| <= d0 + r and |p-c| <= r), so
//                               only the buckets meeting that disc are read; their sites inside the disc are compacted into
//                               shared memory and every pixel takes the exact integer arg-min over them (ties: leftmost
//                               column, then the upper site -- the rule of the scan kernels above).
// Tiles without an unfilled pixel (the fovea core) skip the search entirely.  All distances are integers; the floating-point
// disc test only has to be conservative (it is, by a one-pixel margin).
constexpr int kNtTile = 32;           // tile edge in pixels
constexpr int kNtThreads = 256;       // 4 pixels per thread
constexpr int kNtCand = 3072;         // candidate slots in shared memory (a bucket holds at most 1024 sites)

struct TileGeom { int H, W, nbx, nby, hw, cap; };

template <bool kAllSites>
__global__ void __launch_bounds__(kNtThreads)
nearest_sites_kernel(const int32_t* __restrict__ winner, SelectParams p, TileGeom g, int2* __restrict__ sites,
                     int2* __restrict__ buckets, int* __restrict__ counters) {
  __shared__ int s_warp[kNtThreads / 32];
  __shared__ int s_base;
  const int b = blockIdx.z;
  const int32_t* win = winner + static_cast<size_t>(b) * g.H * g.W;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int y = blockIdx.y * kNtTile + (tid >> 3), x0 = blockIdx.x * kNtTile + (tid & 7) * 4;
  int node[4];
  unsigned mask = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x0 + k;
    node[k] = (y < g.H && x < g.W) ? __ldg(win + static_cast<size_t>(y) * g.W + x) : -1;
    if (node[k] >= 0 && (kAllSites || dilation_covers<true>(DenseWinners{win, p.W}, p, y, x))) mask |= 1u << k;
  }
  const int cnt = __popc(mask);
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kNtThreads / 32; ++w) {
    if (w < warp) before += s_warp[w];
    total += s_warp[w];
  }
  if (tid == 0) {
    s_base = total ? atomicAdd(counters + b, total) : 0;
    // a winner map from fovea_grid_inv_scatter / fovea_scatter_nodes holds every node at most once, so a frame has at most
    // cap sites; a foreign map that breaks this contract loses the excess sites instead of writing past the list
    const int room = max(0, g.cap - s_base);
    buckets[(static_cast<size_t>(b) * g.nby + blockIdx.y) * g.nbx + blockIdx.x] = make_int2(s_base, min(total, room));
  }
  __syncthreads();
  int pos = s_base + before + incl - cnt;
  int2* out = sites + static_cast<size_t>(b) * g.cap;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if ((mask >> k) & 1u) {
      if (pos < g.cap) out[pos] = make_int2((y << 16) | (x0 + k), node[k]);
      ++pos;
    }
}

__global__ void __launch_bounds__(kNtThreads)
nearest_tiles_kernel(const int32_t* __restrict__ winner, const int2* __restrict__ sites, const int2* __restrict__ buckets,
                     const int* __restrict__ counters, TileGeom g, uint16_t* __restrict__ loc) {
  __shared__ int2 cand[kNtCand];
  __shared__ int s_count, s_scan[kNtThreads / 32];
  __shared__ unsigned s_d0[kNtThreads / 32];
  const int b = blockIdx.z, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = blockIdx.y, tx = blockIdx.x;
  const int y = ty * kNtTile + (tid >> 3), x0 = tx * kNtTile + (tid & 7) * 4;
  const size_t fb = static_cast<size_t>(b) * g.H * g.W;
  int node[4];
  bool need = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool in = y < g.H && x0 + k < g.W;
    node[k] = in ? __ldg(winner + fb + static_cast<size_t>(y) * g.W + x0 + k) : 0;
    need |= in && node[k] < 0;
  }
  unsigned best[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
  unsigned bkey[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};   // (x << 16 | y) of the best site: leftmost, then upper
  int bnode[4] = {g.hw, g.hw, g.hw, g.hw};                                    // no site at all: the NaN row
  if (__syncthreads_or(need) && counters[b] > 0) {   // (a frame without any site: the NaN row everywhere)
    const int ntiles = g.nbx * g.nby;
    const int2* bk = buckets + static_cast<size_t>(b) * ntiles;
    const int2* sb = sites + static_cast<size_t>(b) * g.cap;
    const float cy = ty * kNtTile + 0.5f * (kNtTile - 1), cx = tx * kNtTile + 0.5f * (kNtTile - 1);
    // d0: an upper bound of the distance from the tile centre to its nearest site -- the nearest site found in a growing
    // square of buckets around the tile (radius 1, 3, 7, ... buckets) -- as the bit pattern of a non-negative float
    unsigned d0bits = 0x7f800000u;
    for (int k = 1;; k = 2 * k + 1) {
      const int qy0 = max(0, ty - k), qy1 = min(g.nby - 1, ty + k), qx0 = max(0, tx - k), qx1 = min(g.nbx - 1, tx + k);
      const int qw = qx1 - qx0 + 1, nq = qw * (qy1 - qy0 + 1);
      unsigned mind = 0x7f800000u;
      for (int q = tid; q < nq; q += kNtThreads) {
        const int2 e = __ldg(bk + (qy0 + q / qw) * g.nbx + qx0 + q % qw);
        for (int i = 0; i < e.y; ++i) {
          const int sv = __ldg(&sb[e.x + i].x);
          const float dy = static_cast<float>(sv >> 16) - cy, dx = static_cast<float>(sv & 0xFFFF) - cx;
          mind = min(mind, __float_as_uint(sqrtf(dy * dy + dx * dx)));
        }
      }
      mind = __reduce_min_sync(0xffffffffu, mind);
      __syncthreads();  // s_d0 of the previous round has been read
      if (lane == 0) s_d0[warp] = mind;
      __syncthreads();
#pragma unroll
      for (int w = 0; w < kNtThreads / 32; ++w) d0bits = min(d0bits, s_d0[w]);
      if (d0bits != 0x7f800000u || nq == ntiles) break;   // found, or the square already covers the whole frame
    }
    if (d0bits != 0x7f800000u) {
      // disc radius in pixels: d0 + 2r (+1 margin); r = half diagonal of the tile = (T-1)/sqrt(2)
      const float R = __uint_as_float(d0bits) + 2.f * 0.70710678f * (kNtTile - 1) + 1.f;
      const float R2 = R * R;
      const int by0 = max(0, static_cast<int>(floorf((cy - R) / kNtTile))), by1 = min(g.nby - 1, static_cast<int>(floorf((cy + R) / kNtTile)));
      const int bx0 = max(0, static_cast<int>(floorf((cx - R) / kNtTile))), bx1 = min(g.nbx - 1, static_cast<int>(floorf((cx + R) / kNtTile)));
      const int bw = bx1 - bx0 + 1, nb = bw * (by1 - by0 + 1);
      // buckets are visited in chunks of 256 (one per thread); within a chunk, passes of at most kNtCand sites
      for (int c0 = 0; c0 < nb; c0 += kNtThreads) {
        int2 mine = make_int2(0, 0);
        const int q = c0 + tid;
        if (q < nb) {
          const int by = by0 + q / bw, bx = bx0 + q % bw;
          // distance from the centre to the bucket's pixel rectangle
          const float ry = fmaxf(fmaxf(by * kNtTile - cy, cy - (by * kNtTile + kNtTile - 1)), 0.f);
          const float rx = fmaxf(fmaxf(bx * kNtTile - cx, cx - (bx * kNtTile + kNtTile - 1)), 0.f);
          if (ry * ry + rx * rx <= R2) mine = __ldg(bk + by * g.nbx + bx);
        }
        // exclusive block scan of the bucket sizes (upper bound of the candidates they contribute)
        int incl = mine.y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        __syncthreads();  // s_scan / cand of the previous chunk are no longer read
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kNtThreads / 32; ++w) {
          if (w < warp) before += s_scan[w];
          total += s_scan[w];
        }
        const int off = before + incl - mine.y;
        constexpr int kWindow = kNtCand - 1024;  // a pass takes the buckets whose offset falls in one window: <= kNtCand sites
        for (int w0 = 0; w0 < total; w0 += kWindow) {
          if (tid == 0) s_count = 0;
          __syncthreads();
          // every lane walks its own bucket (buckets are small: a few sites on average)
          if (mine.y > 0 && off >= w0 && off < w0 + kWindow) {
            for (int i = 0; i < mine.y; ++i) {
              const int2 sv = __ldg(sb + mine.x + i);
              const float dy = static_cast<float>(sv.x >> 16) - cy, dx = static_cast<float>(sv.x & 0xFFFF) - cx;
              if (dy * dy + dx * dx <= R2) cand[atomicAdd(&s_count, 1)] = sv;
            }
          }
          __syncthreads();
          const int nc = s_count;
          if (need) {
            for (int i = 0; i < nc; ++i) {
              const int2 sv = cand[i];
              const int sy = sv.x >> 16, sx = sv.x & 0xFFFF;
              const int dy = sy - y;
              const unsigned dy2 = static_cast<unsigned>(dy * dy);
              const unsigned key = (static_cast<unsigned>(sx) << 16) | static_cast<unsigned>(sy);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int dx = sx - (x0 + k);
                const unsigned d = dy2 + static_cast<unsigned>(dx * dx);   // < 2^31: coordinates < 32767
                if (d < best[k] || (d == best[k] && key < bkey[k])) { best[k] = d; bkey[k] = key; bnode[k] = sv.y; }
              }
            }
          }
          __syncthreads();
        }
      }
    }
  }
  if (y < g.H) {
    uint16_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = static_cast<uint16_t>(0x8000 | (node[k] >= 0 ? node[k] : bnode[k]));
    uint16_t* dst = loc + fb + static_cast<size_t>(y) * g.W + x0;
    if ((g.W & 3) == 0 && x0 + 3 < g.W) {
      *reinterpret_cast<uint2*>(dst) = make_uint2(o[0] | (static_cast<unsigned>(o[1]) << 16), o[2] | (static_cast<unsigned>(o[3]) << 16));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (x0 + k < g.W) dst[k] = o[k];
    }
  }
}

}  // namespace fovea

using namespace fovea;

// Workspace layout.  Scan path: [B][H][W] short.  Tiled path: per-frame counters [B] int (256-byte aligned block), site
// lists [B][cap] int2, buckets [B][ntiles] int2; sized for cap = min(H*W, 32768) >= h*w sites.
static inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }
static inline int nt_cap(int H, int W) {
  const long long px = static_cast<long long>(H) * W;
  return static_cast<int>(px < 32768 ? px : 32768);
}
extern "C" int64_t fovea_nearest_workspace_bytes(int B, int H, int W) {
  const int64_t scan = static_cast<int64_t>(B) * H * W * static_cast<int64_t>(sizeof(short));
  const int64_t ntiles = static_cast<int64_t>(ceil_div(H, kNtTile)) * ceil_div(W, kNtTile);
  const int64_t tiled = align256(4ll * B) + align256(8ll * B * nt_cap(H, W)) + 8ll * B * ntiles;
  return scan > tiled ? scan : tiled;
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
  // Which exact algorithm: the tiled search wins when sites are sparse (few candidates per 32 x 32 tile), the column/row
  // scans when they are dense.  Measured on B200, 80 x 80 nodes: 1024^2 (spacing ~13 px) tiled 3.2 ms vs scans 2.0 ms per
  // 64 frames; 2048^2 (~26 px) tiled 1.3 ms vs scans 2.4 ms per 16 frames.  Threshold: expected spacing
  // sqrt(H*W / (h*w)) >= 20 pixels.  FOVEA_NEAREST_TILES=0|1 forces one of them (A/B runs).
  static const int force = [] { const char* e = getenv("FOVEA_NEAREST_TILES"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
  const bool tiled = force >= 0 ? force == 1 : static_cast<double>(H) * W >= 400.0 * static_cast<double>(h) * w;
  const int cap = nt_cap(H, W) < h * w ? nt_cap(H, W) : h * w;   // every node lands on at most one pixel
  if (tiled && ceil_div(H, kNtTile) <= 65535) {
    TileGeom tg{H, W, ceil_div(W, kNtTile), ceil_div(H, kNtTile), h * w, cap};
    const int64_t ntiles = static_cast<int64_t>(tg.nbx) * tg.nby;
    char* base = static_cast<char*>(workspace);
    int* counters = reinterpret_cast<int*>(base);
    int2* sites = reinterpret_cast<int2*>(base + align256(4ll * B));
    int2* buckets = reinterpret_cast<int2*>(reinterpret_cast<char*>(sites) + align256(8ll * B * cap));
    FOVEA_CUDA(cudaMemsetAsync(counters, 0, 4 * static_cast<size_t>(B), s));
    dim3 tgrid(tg.nbx, tg.nby, B);
    if (all_sites)
      nearest_sites_kernel<true><<<tgrid, kNtThreads, 0, s>>>(winner, p, tg, sites, buckets, counters);
    else
      nearest_sites_kernel<false><<<tgrid, kNtThreads, 0, s>>>(winner, p, tg, sites, buckets, counters);
    if (int rc = check_launch("fovea_nearest_locate (sites)")) return rc;
    nearest_tiles_kernel<<<tgrid, kNtThreads, 0, s>>>(winner, sites, buckets, counters, tg, loc);
    return check_launch("fovea_nearest_locate (tiles)");
  }
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

// Stage 3: inverse resampling of the low-resolution prediction back to the full-resolution canvas.
//
// Reference: models/models.py:640-655 (grid_inv scatter), :935-938 (F.grid_sample(pred, grid_inv) + NaN mask),
// :159-286 (fillMissingValues_tensor 'tri'), interp2d.py:37-91 (Interp2D), models_instance.py:940 (NaN -> 0),
// models/models.py:1044 (argmax).  The reference materialises grid_inv [B,H,W,2], the sampled scores, a NaN
// mask, a [3,H*W,C] gather and the final scores; here the full-resolution tensor is written exactly once.
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"
#include "mesh.cuh"
#include "fill.cuh"
#include "select.cuh"
#include "taps.cuh"
#include "tma.cuh"

namespace fovea {

// ------------------------------------------------------------------------------------------------ A7
// u = int(((gx+1)/2)*(W-1)), v = int(((gy+1)/2)*(H-1))  -- models/models.py:644-645, fp32 op for op.
__device__ __forceinline__ int target_coord(float g, int size) {
  const float f = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), static_cast<float>(size - 1));
  return f == f ? __float2int_rz(f) : -1;  // NaN (cvt.rzi gives 0): lands nowhere, like any out-of-range target
}

// The same scatter for targets that are already integer pixel coordinates (DynamicFocus int_rount_scale_grid +
// deformed_unsampler, nn_B0_deformed_sampler.py:83-137): coords [B,2,h,w] int64, channel 0 = row, 1 = column.
__global__ void scatter_nodes_kernel(const long long* __restrict__ coords, int32_t* __restrict__ winner, int B, int hw,
                                     int H, int W) {
  const int total = B * hw;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int b = idx / hw, node = idx - b * hw;
    const long long v = coords[static_cast<size_t>(b) * 2 * hw + node], u = coords[static_cast<size_t>(b) * 2 * hw + hw + node];
    if (u >= 0 && u < W && v >= 0 && v < H) atomicMax(winner + (static_cast<size_t>(b) * H + v) * W + u, node);
  }
}

__global__ void grid_inv_scatter_kernel(const float2* __restrict__ grid, int32_t* __restrict__ winner, int B, int hw,
                                        int H, int W) {
  const int total = B * hw;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int b = idx / hw, node = idx - b * hw;
    const float2 g = grid[idx];
    const int u = target_coord(g.x, W), v = target_coord(g.y, H);
    if (u < 0 || u >= W || v < 0 || v >= H) continue;  // NaN / out-of-range grids index nothing
    atomicMax(&winner[(static_cast<size_t>(b) * H + v) * W + u], node);
  }
}

__global__ void grid_inv_canvas_kernel(const int32_t* __restrict__ winner, float2* __restrict__ out, long long total,
                                       int h, int w) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = winner[idx];
    float2 o;
    if (n < 0) {
      o.x = o.y = CUDART_NAN_F;
    } else {
      const int i = n / w, j = n - i * w;
      // models/models.py:652-653: x_cor / w * 2 - 1
      o.x = __fadd_rn(__fmul_rn(__fdiv_rn(static_cast<float>(j), static_cast<float>(w)), 2.f), -1.f);
      o.y = __fadd_rn(__fmul_rn(__fdiv_rn(static_cast<float>(i), static_cast<float>(h)), 2.f), -1.f);
    }
    out[idx] = o;
  }
}

// ------------------------------------------------------------------------------------------------ A8
// table[b][node][c] = grid_sample(pred, grid_inv)[node]: the bilinear sample of pred at the coordinate the
// reference stores for node (i,j).  One CTA handles 64 consecutive nodes of one image: a thread gathers FOUR channels of
// its node (the lanes of a warp along the nodes: coalesced plane reads), parks them as one float4 in a shared-memory tile
// and the tile leaves as whole 16-byte pieces of the channel-contiguous rows.
constexpr int kTabNodes = 64;
constexpr int kTabThreads = 256;
constexpr int kTabQuads = 16;    // channel quads per pass (64 channels)
constexpr int kTabStride = 17;   // float4 per tile row: odd, so a quarter-warp's rows land in distinct bank groups

// kBox = false: the plain transpose table[b][node][c] = pred[b][c][node] (DynamicFocus deformed_unsampler scatters the
// low-resolution labels themselves, nn_B0_deformed_sampler.py:137)
template <bool kBox>
__global__ void __launch_bounds__(kTabThreads, 8)
box4_table_kernel(const float* __restrict__ pred, float* __restrict__ table, int C, int Cs, int h, int w) {
  __shared__ float4 tile[kTabNodes * kTabStride];
  const int hw = h * w;
  const int b = blockIdx.y;
  const int node0 = blockIdx.x * kTabNodes;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 8 warps: 2 node halves x 4 channel-quad groups
  const int nl = (warp & 1) * 32 + lane, qg = warp >> 1;
  const float* pb = pred + static_cast<size_t>(b) * C * hw;
  float* tb = table + static_cast<size_t>(b) * (hw + 2) * Cs;

  const int node = node0 + nl;
  Taps t;
  const bool valid = node < hw;
  // The four tap addresses of a node differ from channel to channel only by the plane offset: formed once, as 32-bit
  // indices into pred[b] (an out-of-range tap reads element 0 and is replaced by 0 -- the zeros padding of grid_sample,
  // exactly: the VALUE is dropped, not weighted by 0).
  unsigned o_nw = 0, o_ne = 0, o_sw = 0, o_se = 0;
  if (valid) {
    const int i = node / w, j = node - i * w;
    const float gx = __fadd_rn(__fmul_rn(__fdiv_rn(static_cast<float>(j), static_cast<float>(w)), 2.f), -1.f);
    const float gy = __fadd_rn(__fmul_rn(__fdiv_rn(static_cast<float>(i), static_cast<float>(h)), 2.f), -1.f);
    t = make_taps(gx, gy, h, w);
    const int base = t.y0 * w + t.x0;
    o_nw = t.ok_nw ? base : 0; o_ne = t.ok_ne ? base + 1 : 0; o_sw = t.ok_sw ? base + w : 0; o_se = t.ok_se ? base + w + 1 : 0;
  }
  auto sample = [&](int c) -> float {
    if (!valid || c >= C) return 0.f;
    const float* pc = pb + static_cast<size_t>(c) * hw;
    if (!kBox) return __ldg(pc + node);
    const float v_nw = __ldg(pc + o_nw), v_ne = __ldg(pc + o_ne), v_sw = __ldg(pc + o_sw), v_se = __ldg(pc + o_se);
    float acc = (t.ok_nw ? v_nw : 0.f) * t.nw;
    acc = fmaf(t.ok_ne ? v_ne : 0.f, t.ne, acc);
    acc = fmaf(t.ok_sw ? v_sw : 0.f, t.sw, acc);
    return fmaf(t.ok_se ? v_se : 0.f, t.se, acc);
  };
  for (int c0 = 0; c0 < Cs; c0 += 4 * kTabQuads) {
    const int nq = min(kTabQuads, (Cs - c0) / 4);   // (Cs is a multiple of 4)
    for (int q = qg; q < nq; q += 4) {
      const int c = c0 + 4 * q;
      tile[nl * kTabStride + q] = make_float4(sample(c), sample(c + 1), sample(c + 2), sample(c + 3));
    }
    __syncthreads();
    // rows leave 16 bytes per thread: 16 threads along a row's quads, 16 rows per pass
    const int q = threadIdx.x & 15;
    if (q < nq)
      for (int n = threadIdx.x >> 4; n < kTabNodes && node0 + n < hw; n += 16)
        reinterpret_cast<float4*>(tb + static_cast<size_t>(node0 + n) * Cs + c0)[q] = tile[n * kTabStride + q];
    __syncthreads();
  }
  if (blockIdx.x == 0)  // row hw: NaN (an image corner no node landed on); row hw+1: zeros (NaN after NaN -> 0)
    for (int c = threadIdx.x; c < Cs; c += kTabThreads) {
      tb[static_cast<size_t>(hw) * Cs + c] = CUDART_NAN_F;
      tb[static_cast<size_t>(hw + 1) * Cs + c] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------------ A9 points
constexpr int kSelThreads = 1024;
constexpr int kSelMax = 16384;  // keys of one frame (h*w + 4 <= 16384); shared memory is sized by the launch to the
                                // next power of two >= cap (64 KB for the 80x80 lattice, 128 KB for 64x128)

// kNB = false: the sites of 'tri' (getPixelsForInterp, models/models.py:169-211: 3x3-cross dilation, four forced corners)
// kNB = true : the sites of 'nearest' / 'BI' (getPixelsForInterp_NB, :213-242: cv2.dilate on the [C,H,W] array = pixels
//              directly above / below only; no forced corners)
template <bool kNB>
__global__ void __launch_bounds__(kSelThreads, 1)
select_points_kernel(const float2* __restrict__ grid, const int32_t* __restrict__ winner, int32_t* __restrict__ pts,
                     int32_t* __restrict__ src, int32_t* __restrict__ npts, SelectParams p) {
  extern __shared__ unsigned long long keys[];  // [kSelMax]  (row<<16|col) << 32 | table row
  __shared__ int count;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int hw = p.h * p.w;
  const int32_t* win = winner + static_cast<size_t>(b) * p.H * p.W;
  if (tid == 0) count = 0;
  __syncthreads();

  for (int node = tid; node < hw; node += kSelThreads) {
    const float2 g = grid[static_cast<size_t>(b) * hw + node];
    const int u = target_coord(g.x, p.W), v = target_coord(g.y, p.H);
    if (u < 0 || u >= p.W || v < 0 || v >= p.H) continue;
    if (win[static_cast<size_t>(v) * p.W + u] != node) continue;  // lost a collision
    const bool corner = (v == 0 || v == p.H - 1) && (u == 0 || u == p.W - 1);
    if (!kNB && corner) continue;  // corners are appended below, exactly once
    if (!dilation_covers<kNB>(DenseWinners{win, p.W}, p, v, u)) continue;
    const int slot = atomicAdd(&count, 1);
    keys[slot] = (static_cast<unsigned long long>((v << 16) | u) << 32) | static_cast<unsigned>(node);
  }
  if (!kNB && tid < 4) {  // models/models.py:202-209: the four corners are always interpolation points
    const int v = (tid & 2) ? p.H - 1 : 0, u = (tid & 1) ? p.W - 1 : 0;
    // (degenerate 1-pixel-wide canvases would duplicate corners; the host rejects H,W < 2)
    const int n = win[static_cast<size_t>(v) * p.W + u];
    const int slot = atomicAdd(&count, 1);
    keys[slot] = (static_cast<unsigned long long>((v << 16) | u) << 32) | static_cast<unsigned>(n < 0 ? hw : n);
  }
  __syncthreads();
  const int n = count;
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int i = n + tid; i < n2; i += kSelThreads) keys[i] = ~0ull;
  __syncthreads();
  // bitonic sort, ascending by pixel key = row-major order of torch.where (models/models.py:265)
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n2; i += kSelThreads) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], c = keys[l];
          const bool up = (i & k) == 0;
          if ((a > c) == up) { keys[i] = c; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n; i += kSelThreads) {
    pts[static_cast<size_t>(b) * p.cap + i] = static_cast<int32_t>(keys[i] >> 32);
    src[static_cast<size_t>(b) * p.cap + i] = static_cast<int32_t>(keys[i] & 0xFFFFFFFFull);
  }
  if (tid == 0) npts[b] = n;
}

// The same sites WITHOUT the dense winner map (4 bytes per canvas pixel to clear, scatter into and probe): the frame's
// <= 6 400 node targets are sorted in shared memory by (pixel, node); the last entry of a run of equal pixels is the
// node that wins the pixel (atomicMax of the A7 scatter), "is this pixel filled" is a binary search, and the sites come
// out already in row-major order.  Also written: targets[b][n] = (row << 16 | column) of node n if it won its pixel,
// else -1 -- what the raster needs to stamp the node pixels -- with four more entries for image corners nobody landed on.
__global__ void __launch_bounds__(kSelThreads, 1)
select_points_sparse_kernel(const float2* __restrict__ grid, int32_t* __restrict__ pts, int32_t* __restrict__ src,
                            int32_t* __restrict__ npts, int32_t* __restrict__ targets, SelectParams p) {
  extern __shared__ unsigned long long keys[];  // (row<<16|col) << 32 | node + 1   (0: a corner's place holder)
  __shared__ int s_warp[kSelThreads / 32];
  __shared__ int s_base;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hw = p.h * p.w;
  const int n = hw + 4;
  int32_t* tg = targets + static_cast<size_t>(b) * n;
  for (int node = tid; node < n; node += kSelThreads) {
    tg[node] = -1;
    unsigned long long key = ~0ull;   // targets outside the canvas sort last and win nothing
    if (node < hw) {
      const float2 g = grid[static_cast<size_t>(b) * hw + node];
      const int u = target_coord(g.x, p.W), v = target_coord(g.y, p.H);
      if (u >= 0 && u < p.W && v >= 0 && v < p.H)
        key = (static_cast<unsigned long long>((v << 16) | u) << 32) | static_cast<unsigned>(node + 1);
    } else {                          // models/models.py:202-209: the four corners are always interpolation points
      const int c = node - hw, v = (c & 2) ? p.H - 1 : 0, u = (c & 1) ? p.W - 1 : 0;
      key = static_cast<unsigned long long>((v << 16) | u) << 32;
    }
    keys[node] = key;
  }
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int i = n + tid; i < n2; i += kSelThreads) keys[i] = ~0ull;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int k = 2; k <= n2; k <<= 1) {   // bitonic sort, ascending
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n2; i += kSelThreads) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], c = keys[l];
          const bool up = (i & k) == 0;
          if ((a > c) == up) { keys[i] = c; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  const SortedTargets filled{keys, n};
  // 1024 entries at a time: winners of their pixel decide whether they are sites; a block-wide scan keeps the order
  for (int i0 = 0; i0 < n; i0 += kSelThreads) {
    const int i = i0 + tid;
    const unsigned long long key = i < n ? keys[i] : ~0ull;
    const unsigned pix = static_cast<unsigned>(key >> 32), tag = static_cast<unsigned>(key);
    const bool live = key != ~0ull;
    const bool wins = live && (i + 1 >= n || static_cast<unsigned>(keys[i + 1] >> 32) != pix);   // last of its run
    bool site = false;
    int source = hw;
    if (wins) {
      const int v = static_cast<int>(pix >> 16), u = static_cast<int>(pix & 0xFFFFu);
      const bool corner = (v == 0 || v == p.H - 1) && (u == 0 || u == p.W - 1);
      if (tag) { source = static_cast<int>(tag) - 1; tg[source] = static_cast<int32_t>(pix); }
      else tg[hw + ((v ? 2 : 0) | (u ? 1 : 0))] = static_cast<int32_t>(pix);   // an unfilled corner: "no value" is stamped there
      site = corner || dilation_covers<false>(filled, p, v, u);
    }
    const unsigned ball = __ballot_sync(0xffffffffu, site);
    if (lane == 0) s_warp[warp] = __popc(ball);
    __syncthreads();
    int before = s_base;
    for (int wq = 0; wq < warp; ++wq) before += s_warp[wq];
    const int slot = before + __popc(ball & ((1u << lane) - 1u));
    if (site) {
      pts[static_cast<size_t>(b) * p.cap + slot] = static_cast<int32_t>(pix);
      src[static_cast<size_t>(b) * p.cap + slot] = source;
    }
    __syncthreads();
    if (tid == kSelThreads - 1) s_base = slot + (site ? 1 : 0);
    __syncthreads();
  }
  if (tid == 0) npts[b] = s_base;
}

// ------------------------------------------------------------------------------------------------ hints
// One CTA per image stages the mesh records and the points in shared memory (<= 205 KB + 26 KB) so that the long
// first-level walks run at shared-memory latency.  Three levels: 16x16 probes walked from triangle 0, then
// 32x32-pixel cells walked from the nearest probe, then the 32x8 hint cells walked from their 32x32 parent.
constexpr int kHintThreads = 1024;

__global__ void __launch_bounds__(kHintThreads, 1)
locate_hints_kernel(const int32_t* __restrict__ pts, const int32_t* __restrict__ npts, const uint4* __restrict__ mesh,
                    const int32_t* __restrict__ ntri, int cap, int tcap, int H, int W, int32_t* __restrict__ hints,
                    int32_t* __restrict__ mid_ws) {
  extern __shared__ __align__(16) uint4 srec[];
  __shared__ int coarse[16 * 16];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int T = ntri[b];
  const int ch = ceil_div(H, FOVEA_HINT_CELL_H), cw = ceil_div(W, FOVEA_HINT_CELL_W);
  int32_t* hb = hints + static_cast<size_t>(b) * ch * cw;
  if (T <= 0) {
    for (int i = tid; i < ch * cw; i += kHintThreads) hb[i] = 0;
    return;
  }
  const uint4* mb = mesh + static_cast<size_t>(b) * tcap;
  for (int t = tid; t < T; t += kHintThreads) srec[t] = mb[t];
  int* spts = reinterpret_cast<int*>(srec + tcap);
  const int n = npts[b];
  for (int i = tid; i < n; i += kHintThreads) spts[i] = pts[static_cast<size_t>(b) * cap + i];
  __syncthreads();
  Mesh m{spts, srec, T};
  if (tid < 256) {
    const int py = tid / 16, px = tid % 16;
    const int qr = min(H - 1, (2 * py + 1) * H / 32), qc = min(W - 1, (2 * px + 1) * W / 32);
    const Located L = locate(m, qr, qc, 0);
    coarse[tid] = L.tri < 0 ? 0 : L.tri;
  }
  __syncthreads();
  // level 1: 32 x 32 pixel cells (4 hint cells tall) -> mid_ws [ceil(ch/4), cw]
  const int mh = ceil_div(ch, 4);
  int32_t* mid = mid_ws + static_cast<size_t>(b) * mh * cw;
  for (int i = tid; i < mh * cw; i += kHintThreads) {
    const int cy = i / cw, cx = i - cy * cw;
    const int qr = min(H - 1, cy * 32 + 16), qc = min(W - 1, cx * FOVEA_HINT_CELL_W + FOVEA_HINT_CELL_W / 2);
    const int py = min(15, qr * 16 / H), px = min(15, qc * 16 / W);
    const Located L = locate(m, qr, qc, coarse[py * 16 + px]);
    mid[i] = L.tri < 0 ? 0 : L.tri;
  }
  __syncthreads();
  for (int i = tid; i < ch * cw; i += kHintThreads) {
    const int cy = i / cw, cx = i - cy * cw;
    const int qr = min(H - 1, cy * FOVEA_HINT_CELL_H + FOVEA_HINT_CELL_H / 2);
    const int qc = min(W - 1, cx * FOVEA_HINT_CELL_W + FOVEA_HINT_CELL_W / 2);
    const Located L = locate(m, qr, qc, mid[(cy / 4) * cw + cx]);
    hb[i] = L.tri < 0 ? 0 : L.tri;
  }
}

// ------------------------------------------------------------------------------------------------ A9 location
// Stage 3 is split in two kernels.  locate_pixels (below) resolves, for every full-resolution pixel, which low-res
// node or which triangle of the mesh produces its value -- 2 B per pixel, a function of the sampling grid only.
// inverse_fill then streams the C channels of every pixel from that map with no walks and no divergent loops, so
// the store-bound kernel runs at a register count / occupancy chosen for streaming, and the location work can
// overlap the encoder on a side stream (it does not depend on `pred`).
//
//   loc[b,y,x] >= 0        : id of the triangle that owns pixel (y,x)          (interp2d.py:58 find_simplex)
//   loc[b,y,x] = -(n+1)    : the pixel received low-res node n directly        (models/models.py:650-651);   [signed,
//                            in-kernel form; stored as 16 bits, see encode_loc]
//                            n == h*w: no value (outside the triangulation)    -> the NaN row of the value table

__global__ void __launch_bounds__(256)
triangle_setup_kernel(const int32_t* __restrict__ pts, const int32_t* __restrict__ src, const uint4* __restrict__ mesh,
                      const int32_t* __restrict__ ntri, TriRec* __restrict__ recs, int cap, int tcap, int nan_row) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntri[b]) return;
  const int32_t* ptsb = pts + static_cast<size_t>(b) * cap;
  const int32_t* srcb = src + static_cast<size_t>(b) * cap;
  const uint4 q = __ldg(mesh + static_cast<size_t>(b) * tcap + t);
  const unsigned i0 = q.x & 0xFFFFu, i1 = q.x >> 16, i2 = q.y & 0xFFFFu;
  const unsigned n0 = q.z & 0xFFFFu, n1 = q.z >> 16, n2 = q.w & 0xFFFFu;
  const int p0 = __ldg(ptsb + i0), p1 = __ldg(ptsb + i1), p2 = __ldg(ptsb + i2);
  const int r[3] = {p0 >> 16, p1 >> 16, p2 >> 16}, c[3] = {p0 & 0xFFFF, p1 & 0xFFFF, p2 & 0xFFFF};
  const int area2 = (c[1] - c[0]) * (r[2] - r[0]) - (r[1] - r[0]) * (c[2] - c[0]);  // orient(p0,p1,p2)
  const int s = area2 < 0 ? -1 : 1;
  const unsigned nb[3] = {n0, n1, n2};
  const int s0 = __ldg(srcb + i0), s1 = __ldg(srcb + i1), s2 = __ldg(srcb + i2);
  const bool nan_me = s0 == nan_row || s1 == nan_row || s2 == nan_row;  // a vertex without a value (unfilled corner)
  int A[3], Bc[3], Cc[3];
  unsigned m = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {  // edge opposite vertex i runs from vertex (i+1)%3 to vertex (i+2)%3
    const int ar = r[(i + 1) % 3], ac = c[(i + 1) % 3], br = r[(i + 2) % 3], bc = c[(i + 2) % 3];
    const int dr = br - ar, dc = bc - ac;
    A[i] = s * dc;
    Bc[i] = -s * dr;
    Cc[i] = s * (dr * ac - dc * ar);
    // Who owns a pixel lying EXACTLY on this edge?  Hull edge: this triangle.  One side has a vertex without a value
    // and the other has not: the side that can produce a value (the reference's eps-tolerant walk accepts either;
    // coming from the filled region it meets the finite one, e.g. the row of the last nodes above an unfilled image
    // corner).  Otherwise the top-left rule of mesh.cuh (edge_owned at e == 0).
    bool tie_owned = true;
    if (nb[i] != kNoTri) {
      const uint4 qn = __ldg(mesh + static_cast<size_t>(b) * tcap + nb[i]);
      const bool nan_nb = __ldg(srcb + (qn.x & 0xFFFFu)) == nan_row || __ldg(srcb + (qn.x >> 16)) == nan_row ||
                          __ldg(srcb + (qn.y & 0xFFFFu)) == nan_row;
      tie_owned = nan_me != nan_nb ? !nan_me : (dr != 0 ? (s * dr > 0) : (s * dc > 0));
    }
    if (!tie_owned) m |= 1u << i;
  }
  const int area = area2 < 0 ? -area2 : area2;
  const double inv = area ? 1.0 / static_cast<double>(area) : 0.0;
  TriRec R;
  R.q0 = make_uint4(A[0], Bc[0], Cc[0], A[1]);
  R.q1 = make_uint4(Bc[1], Cc[1], A[2], Bc[2]);
  R.q2 = make_uint4(Cc[2], n0 | (n1 << 16), n2 | (m << 16), area);
  R.q3 = make_uint4(static_cast<unsigned>(s0) | (static_cast<unsigned>(s1) << 16), static_cast<unsigned>(s2),
                    static_cast<unsigned>(__double2loint(inv)),
                    static_cast<unsigned>(__double2hiint(inv)));
  recs[static_cast<size_t>(b) * tcap + t] = R;
}

// One triangle held in registers while a lane walks: its three edge functions evaluated at the lane's pixel.
struct WalkState {
  int t;
  int e0, e1, e2;      // edge functions at the current pixel
  int b0, b1, b2;      // d e_i / d col
  unsigned n01, n2m;   // n0 | n1 << 16,  n2 | m << 16
  int area;
};

__device__ __forceinline__ void walk_load(WalkState& S, const TriRec* __restrict__ recs, int t, int y, int x) {
  const uint4* r = reinterpret_cast<const uint4*>(recs + t);
  const uint4 q0 = __ldg(r), q1 = __ldg(r + 1), q2 = __ldg(r + 2);
  S.t = t;
  S.b0 = static_cast<int>(q0.y); S.b1 = static_cast<int>(q1.x); S.b2 = static_cast<int>(q1.w);
  S.e0 = static_cast<int>(q0.x) * y + S.b0 * x + static_cast<int>(q0.z);
  S.e1 = static_cast<int>(q0.w) * y + S.b1 * x + static_cast<int>(q1.y);
  S.e2 = static_cast<int>(q1.z) * y + S.b2 * x + static_cast<int>(q2.x);
  S.n01 = q2.y; S.n2m = q2.z;
  S.area = static_cast<int>(q2.w);
}

// -1: the pixel belongs to S.t; otherwise the neighbour across the first edge that rejects it (kNoTri = hull)
__device__ __forceinline__ int walk_test(const WalkState& S) {
  const unsigned m = S.n2m >> 16;
  if (S.e0 < static_cast<int>(m & 1u)) return static_cast<int>(S.n01 & 0xFFFFu);
  if (S.e1 < static_cast<int>((m >> 1) & 1u)) return static_cast<int>(S.n01 >> 16);
  if (S.e2 < static_cast<int>((m >> 2) & 1u)) return static_cast<int>(S.n2m & 0xFFFFu);
  return -1;
}

__device__ __noinline__ int walk_bruteforce(const TriRec* __restrict__ recs, int T, int y, int x) {
  WalkState S;
  for (int t = 0; t < T; ++t) {
    walk_load(S, recs, t, y, x);
    if (S.area != 0 && walk_test(S) < 0) return t;
  }
  return -1;
}

constexpr int kLocThreads = 256;
constexpr int kLocRun = FOVEA_HINT_CELL_W;           // pixels one thread walks along its row (= one hint cell, 32)
constexpr int kLocTileW = 8 * kLocRun, kLocTileH = 32;  // CTA tile: 2 warps across x 4 down; warp = 4 runs x 8 rows
constexpr int kLocStride = kLocRun + 1;              // shared-memory row stride (bank-conflict-free)
static_assert(kLocRun == 32, "one run = one bit mask = one coalesced 128-byte row of the staging tile");

// floor(num / den) for 0 <= num, 0 < den, saturated at `cap` (only quotients below the run length matter)
__device__ __forceinline__ int floor_div_capped(int num, int den, int cap) {
  const float qf = fminf(__fdividef(static_cast<float>(num), static_cast<float>(den)), static_cast<float>(cap));
  int q = static_cast<int>(qf);
  if (q >= cap) return cap;
  if (q * den > num) --q;             // float rounding is within one unit for q < cap <= 64
  else if ((q + 1) * den <= num) ++q;
  return q;
}

// Scan-line span walker.  Each lane owns one 32-pixel run of one image row.  Instead of testing pixel after pixel
// (which serialises the warp on whichever lane happens to cross an edge), a lane that stands in triangle t at column
// x computes in closed form how many further pixels of its row t owns (the edge functions are linear in x), emits
// the whole span into a shared-memory staging tile, and crosses into the next triangle: all lanes take their
// transitions in lock step, and the per-pixel work is one shared-memory store.  Pixels that received a node (A7
// winners) are masked out up front, so the densely filled fovea costs no walks at all.  The tile is then merged with
// the winners and written with coalesced 128-byte rows.
__global__ void __launch_bounds__(kLocThreads)
locate_pixels_kernel(const int32_t* __restrict__ winner, const TriRec* __restrict__ trirec,
                     const int32_t* __restrict__ ntri, const int32_t* __restrict__ hints, uint16_t* __restrict__ loc,
                     int hw, int H, int W, int tcap) {
  __shared__ int tile_all[kLocThreads / 32][32 * kLocStride];
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* tile = tile_all[warp];
  // lane l of this warp owns run (l & 3) of row (l >> 2) of the warp's 128 x 8 pixel tile (measured: 0.63 ms per 64
  // frames of 1024^2 against 0.82 ms for a 32 x 32 block per warp, whose row-per-lane staging reads coalesce worse)
  const int wx0 = blockIdx.x * kLocTileW + (warp & 1) * 4 * kLocRun;
  const int wy0 = blockIdx.y * kLocTileH + (warp >> 1) * 8;
  if (wx0 >= W || wy0 >= H) return;  // warp-uniform
  const size_t img = static_cast<size_t>(b) * H * W;
  const int none = -(hw + 1);

  // ---- which pixels of my run need a triangle?  (bit j of `need`: pixel x0 + j received no node)
  // ---- stage the winners: tile[l][j] = -(n+1) for a pixel that received node n (its final value); bit j of `need` =
  // pixel x0 + j received no node and needs a triangle.  (Fully unrolled: all 32 independent loads are in flight at once.)
  unsigned need = 0;
#pragma unroll
  for (int l = 0; l < 32; ++l) {
    const int yy = wy0 + (l >> 2), xx = wx0 + (l & 3) * kLocRun + lane;
    const int n = (yy < H && xx < W) ? __ldg(winner + img + static_cast<size_t>(yy) * W + xx) : 0;
    tile[l * kLocStride + lane] = -(n + 1);
    const unsigned m = __ballot_sync(0xffffffffu, n < 0);
    if (l == lane) need = m;
  }
  __syncwarp();

  const int x0 = wx0 + (lane & 3) * kLocRun;
  const int y = wy0 + (lane >> 2);
  const int T = ntri[b];
  int* mine = tile + lane * kLocStride;
  if (T <= 0)  // empty mesh: nothing owns anything
    for (unsigned m = need; m; m &= m - 1) mine[__ffs(m) - 1] = none;
  if (need != 0 && T > 0) {
    const TriRec* recs = trirec + static_cast<size_t>(b) * tcap;
    const int start = hints[(static_cast<size_t>(b) * ceil_div(H, FOVEA_HINT_CELL_H) + y / FOVEA_HINT_CELL_H) *
                                ceil_div(W, FOVEA_HINT_CELL_W) + x0 / FOVEA_HINT_CELL_W];
    WalkState S;
    bool have = false;
    int sx = 0;  // column (relative to x0) the edge functions of S are evaluated at
    while (need) {
      const int j = __ffs(need) - 1;  // next pixel that needs a triangle
      const int x = x0 + j;
      bool ok = false;
      if (have) {  // move the held triangle's edge functions to column x
        const int dj = j - sx;
        S.e0 += dj * S.b0; S.e1 += dj * S.b1; S.e2 += dj * S.b2;
        sx = j;
        ok = walk_test(S) < 0;
      }
      if (!ok) {
        int t = have ? S.t : ((start >= 0 && start < T) ? start : 0);
        for (int step = 0; step < T + 8; ++step) {
          walk_load(S, recs, t, y, x);
          if (S.area == 0) break;
          const int nxt = walk_test(S);
          if (nxt < 0) { ok = true; break; }
          if (static_cast<unsigned>(nxt) == kNoTri) break;
          t = nxt;
        }
        if (!ok) {  // degenerate triangle on the way (host meshes only) or outside the hull: exhaustive search
          const int tb = walk_bruteforce(recs, T, y, x);
          if (tb >= 0) { walk_load(S, recs, tb, y, x); ok = true; }
        }
        have = ok;
        sx = j;
      }
      if (!ok) {  // no triangle owns this pixel
        mine[j] = none;
        need &= need - 1;
        continue;
      }
      // how many pixels to the right does S.t still own?  e_i(x + k) = e_i + k B_i must stay >= m_i; only edges the
      // row is running towards (B_i < 0) can end the span
      const unsigned m = S.n2m >> 16;
      int more = kLocRun;
      if (S.b0 < 0) more = min(more, floor_div_capped(S.e0 - static_cast<int>(m & 1u), -S.b0, kLocRun));
      if (S.b1 < 0) more = min(more, floor_div_capped(S.e1 - static_cast<int>((m >> 1) & 1u), -S.b1, kLocRun));
      if (S.b2 < 0) more = min(more, floor_div_capped(S.e2 - static_cast<int>((m >> 2) & 1u), -S.b2, kLocRun));
      const int last = min(j + more, kLocRun - 1);                       // last owned column of this run
      const unsigned upto = last >= 31 ? 0xffffffffu : ((2u << last) - 1u);
      for (unsigned e = need & upto; e; e &= e - 1) mine[__ffs(e) - 1] = S.t;  // node pixels keep their own value
      need &= ~upto;                                                     // every pixel up to `last` is settled
    }
  }
  __syncwarp();

  // ---- write the tile: row l of the staging tile is 32 consecutive pixels = one 128-byte segment
#pragma unroll 8
  for (int l = 0; l < 32; ++l) {
    const int yy = wy0 + (l >> 2), xx = wx0 + (l & 3) * kLocRun + lane;
    if (yy < H && xx < W) loc[img + static_cast<size_t>(yy) * W + xx] = encode_loc(tile[l * kLocStride + lane]);
  }
}

// ------------------------------------------------------------------------------------------------ A8+A9+A10
#ifndef FOVEA_FILL_THREADS
#define FOVEA_FILL_THREADS 256
#endif
constexpr int kFillThreads = FOVEA_FILL_THREADS;
#ifndef FOVEA_FILL_TILES_Y
#define FOVEA_FILL_TILES_Y 1
#endif
constexpr int kFillTilesY = FOVEA_FILL_TILES_Y;  // vertically adjacent tiles streamed by one CTA
// Thread -> pixel mapping.  One thread = 4 consecutive pixels; one warp = WL lanes across x (32 / WL) rows; one CTA =
// WX warps across x (8 / WX) down.  What matters is how few row segments one warp-wide store touches -- measured on
// 64 frames of 1024^2 (ms per launch, fraction of the 6 542 GB/s peak):
//   warp 16 px x 8 rows (WL 4): 2.52 (0.83)   32 x 4 (WL 8): 2.50-2.64 depending on the CTA tile (0.79-0.84)
//   warp 64 px x 2 rows (WL 16): 2.28-2.29 (0.915-0.918) for CTA tiles 64x16, 128x8, 256x4    128 x 1 (WL 32): 2.29-2.36
// i.e. two 256-byte segments per store instruction instead of four 128-byte ones is worth 9 %.  WL = 16 with a 128 x 8
// CTA tile (WX = 2) is the default; canvases narrower than 128 pixels use WX = 1.
#ifndef FOVEA_FILL_WL
#define FOVEA_FILL_WL 16  // lanes of a warp across a row (x 4 pixels each); the warp covers 32 / WL rows
#endif
constexpr int kFillTileW = 128, kFillTileH = kFillThreads / 32;  // the default CTA tile (also the store-ceiling probe's)

// The table rows of the three vertices of one pixel, G channels each.
template <int G>
struct Rows {
  float a[G], b[G], c[G];
};

// r = p ? (rows at pa, pb, pc) : unchanged -- three predicated loads under ONE predicate (no branch, no register
// copies): the rows of a pixel are re-read only if they differ from its left neighbour's.  G = 4: 128-bit loads;
// G = 8: the 256-bit loads of sm_100 (LDG.E.256), which halve the number of L1 requests per output byte.
template <int G>
__device__ __forceinline__ void ldg3_if(Rows<G>& r, unsigned long long pa, unsigned long long pb, unsigned long long pc,
                                        unsigned p);

template <>
__device__ __forceinline__ void ldg3_if<4>(Rows<4>& r, unsigned long long pa, unsigned long long pb,
                                           unsigned long long pc, unsigned p) {
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %15, 0;\n\t"
      "@q ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%12];\n\t"
      "@q ld.global.nc.v4.f32 {%4,%5,%6,%7}, [%13];\n\t"
      "@q ld.global.nc.v4.f32 {%8,%9,%10,%11}, [%14];\n\t}"
      : "+f"(r.a[0]), "+f"(r.a[1]), "+f"(r.a[2]), "+f"(r.a[3]), "+f"(r.b[0]), "+f"(r.b[1]), "+f"(r.b[2]), "+f"(r.b[3]),
        "+f"(r.c[0]), "+f"(r.c[1]), "+f"(r.c[2]), "+f"(r.c[3])
      : "l"(pa), "l"(pb), "l"(pc), "r"(p));
}

template <>
__device__ __forceinline__ void ldg3_if<8>(Rows<8>& r, unsigned long long pa, unsigned long long pb,
                                           unsigned long long pc, unsigned p) {
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %27, 0;\n\t"
      "@q ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%24];\n\t"
      "@q ld.global.nc.v8.f32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%25];\n\t"
      "@q ld.global.nc.v8.f32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%26];\n\t}"
      : "+f"(r.a[0]), "+f"(r.a[1]), "+f"(r.a[2]), "+f"(r.a[3]), "+f"(r.a[4]), "+f"(r.a[5]), "+f"(r.a[6]), "+f"(r.a[7]),
        "+f"(r.b[0]), "+f"(r.b[1]), "+f"(r.b[2]), "+f"(r.b[3]), "+f"(r.b[4]), "+f"(r.b[5]), "+f"(r.b[6]), "+f"(r.b[7]),
        "+f"(r.c[0]), "+f"(r.c[1]), "+f"(r.c[2]), "+f"(r.c[3]), "+f"(r.c[4]), "+f"(r.c[5]), "+f"(r.c[6]), "+f"(r.c[7])
      : "l"(pa), "l"(pb), "l"(pc), "r"(p));
}

__device__ __forceinline__ uint2 fill_loc(const uint16_t* __restrict__ loc, const FillParams& p, int b, int x0, int y) {
  return __ldcs(reinterpret_cast<const uint2*>(loc + (static_cast<size_t>(b) * p.H + y) * p.W + x0));
}

// Each thread owns 4 consecutive pixels of one row: resolve their table rows + barycentric weights from `loc`
// (exact integer edge functions, stepped in registers while the triangle does not change), then stream all channels
// with 128-bit stores.  Table rows are re-loaded only where they differ from the previous pixel's.
template <bool kScores, bool kMask, int G>
__device__ __forceinline__ void fill_tile(const uint16_t* __restrict__ loc, const TriRec* __restrict__ trirec,
                                          const float* __restrict__ table, float* __restrict__ scores,
                                          long long* __restrict__ mask, const FillParams& p, int b, int x0, int y,
                                          const uint2 l2) {   // l2: the tile's 4 x 16 bits of `loc` (fill_loc)
  const int hw = p.h * p.w;
  const unsigned pixoff = static_cast<unsigned>(y) * p.W + x0;  // H*W < 2^32 is checked on the host
  const size_t plane = static_cast<size_t>(p.H) * p.W;
  const TriRec* recs = trirec + static_cast<size_t>(b) * p.tcap;

  const int lc[4] = {decode_loc(l2.x & 0xFFFFu), decode_loc(l2.x >> 16), decode_loc(l2.y & 0xFFFFu), decode_loc(l2.y >> 16)};
  unsigned nd0[4], nd1[4], nd2[4];  // table rows (node ids) of the three vertices of each pixel
  float w0[4], w1[4], w2[4];        // barycentric weights ((1,0,0) for a pixel that received a node)
  unsigned nanmask = 0;             // pixels whose value is NaN in every channel
  unsigned reload[4] = {1u, 0u, 0u, 0u};  // pixel k's table rows differ from pixel k-1's

  int cur = -1, sn0 = hw, sn1 = hw, sn2 = hw;
  int e0 = 0, e1 = 0, d0 = 0, d1 = 0;
  double inv_area = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int n0, n1, n2;
    float a0 = 1.f, a1 = 0.f, a2 = 0.f;
    if (lc[k] < 0) {
      n0 = n1 = n2 = -(lc[k] + 1);
    } else {
      if (lc[k] != cur) {  // one setup record = three independent 16-byte loads (no pts / src indirection)
        cur = lc[k];
        const uint4* r = reinterpret_cast<const uint4*>(recs + cur);
        const uint4 q0 = __ldg(r), q1 = __ldg(r + 1), q3 = __ldg(r + 3);
        d0 = static_cast<int>(q0.y);
        d1 = static_cast<int>(q1.x);
        e0 = static_cast<int>(q0.x) * y + d0 * (x0 + k) + static_cast<int>(q0.z);
        e1 = static_cast<int>(q0.w) * y + d1 * (x0 + k) + static_cast<int>(q1.y);
        sn0 = static_cast<int>(q3.x & 0xFFFFu); sn1 = static_cast<int>(q3.x >> 16); sn2 = static_cast<int>(q3.y);
        inv_area = __hiloint2double(static_cast<int>(q3.w), static_cast<int>(q3.z));
      }
      // interp2d.py:58-65 / qhull.pyx:1210-1264: c0, c1 in float64, c2 = 1 - c0 - c1, then cast to float32
      const double c0 = static_cast<double>(e0) * inv_area;
      const double c1 = static_cast<double>(e1) * inv_area;
      a0 = static_cast<float>(c0);
      a1 = static_cast<float>(c1);
      // (1 - c0 - c1 is exactly what the reference computes; when the pixel lies ON the edge opposite vertex 2 its true
      //  value is 0 and the double expression is rounding noise of either sign, ~1e-16 -- the reference's own noise comes
      //  from a different formula (an LU-inverted transform).  Clamped at zero so that every weight is >= 0, the property
      //  the pruned arg-max fill (mask_fill.cu) relies on; the scores move by <= 1e-16 * |value| at those pixels.)
      a2 = fmaxf(static_cast<float>(1.0 - c0 - c1), 0.f);
      n0 = sn0; n1 = sn1; n2 = sn2;
    }
    e0 += d0;  // advance to column x0 + k + 1 (harmless while no triangle is held: d == 0)
    e1 += d1;
    if (n0 == hw || n1 == hw || n2 == hw) {  // a NaN vertex poisons every channel (NaN*w, even for w == 0)
      nanmask |= 1u << k;
      n0 = n1 = n2 = p.zero_residual ? hw + 1 : hw;  // models_instance.py:940: residual NaN -> 0
      a0 = 1.f; a1 = 0.f; a2 = 0.f;
    }
    nd0[k] = static_cast<unsigned>(n0);
    nd1[k] = static_cast<unsigned>(n1);
    nd2[k] = static_cast<unsigned>(n2);
    w0[k] = a0; w1[k] = a1; w2[k] = a2;
    if (k > 0) reload[k] = (nd0[k] != nd0[k - 1]) | (nd1[k] != nd1[k - 1]) | (nd2[k] != nd2[k - 1]);
  }

  // The channel loop forms each row address as base + node * row_bytes (one IMAD.WIDE); the base advances by 4*G bytes
  // per G-channel group.
  const unsigned row_bytes = static_cast<unsigned>(p.Cs) * 4u;
  unsigned long long tbase = reinterpret_cast<unsigned long long>(table + static_cast<size_t>(b) * (hw + 2) * p.Cs);
  unsigned long long obase = reinterpret_cast<unsigned long long>(
      kScores ? scores + static_cast<size_t>(b) * p.C * plane + pixoff : nullptr);
  const unsigned long long ostep = static_cast<unsigned long long>(plane) * 4ull;
  // keep the hoisted bases in registers (ptxas otherwise rematerialises the 64-bit products inside the loop)
  asm volatile("" : "+l"(tbase), "+l"(obase));
  float best[4] = {0.f, 0.f, 0.f, 0.f};
  int besti[4] = {0, 0, 0, 0};


  for (int c = 0; c < p.Cs; c += G, tbase += 4ull * G) {
    float v[4][G];  // [pixel][channel]
    Rows<G> r;
#pragma unroll
    for (int e = 0; e < G; ++e) r.a[e] = r.b[e] = r.c[e] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ldg3_if<G>(r, tbase + static_cast<unsigned long long>(nd0[k]) * row_bytes,
                 tbase + static_cast<unsigned long long>(nd1[k]) * row_bytes,
                 tbase + static_cast<unsigned long long>(nd2[k]) * row_bytes, reload[k]);
      // interp2d.py:85-89: mul, then sum over the three vertices in order (separate roundings, no FMA)
#pragma unroll
      for (int e = 0; e < G; ++e)
        v[k][e] = __fadd_rn(__fadd_rn(__fmul_rn(r.a[e], w0[k]), __fmul_rn(r.b[e], w1[k])), __fmul_rn(r.c[e], w2[k]));
    }
#pragma unroll
    for (int e = 0; e < G; ++e) {
      if (c + e < p.C) {
        if (kScores) {
          __stcs(reinterpret_cast<float4*>(obase), make_float4(v[0][e], v[1][e], v[2][e], v[3][e]));
          obase += ostep;
        }
        if (kMask) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // torch.argmax: the first maximum wins.  (A NaN in SOME channels of a pixel --
            // only possible when `pred` itself holds NaN / Inf -- is skipped here, where torch would return its index:
            // measured, the NaN-propagating compare costs the mask variants 20 % (2.37 vs 1.98 ms) for a case no finite
            // prediction reaches; fovea_argmax_classes on materialised scores has torch's NaN rule.)
            if (c + e == 0 || v[k][e] > best[k]) { best[k] = v[k][e]; besti[k] = c + e; }
          }
        }
      }
    }
  }
  if (kMask) {
    // a pixel that is NaN in every channel: torch.argmax returns the first NaN -> class 0
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if ((nanmask >> k) & 1u) besti[k] = 0;
    if (p.mask_u8) {
      *reinterpret_cast<uchar4*>(reinterpret_cast<unsigned char*>(mask) + static_cast<size_t>(b) * plane + pixoff) =
          make_uchar4(besti[0], besti[1], besti[2], besti[3]);
    } else {
      longlong2* mp = reinterpret_cast<longlong2*>(mask + static_cast<size_t>(b) * plane + pixoff);
      mp[0] = make_longlong2(besti[0], besti[1]);
      mp[1] = make_longlong2(besti[2], besti[3]);
    }
  }
}

// One CTA streams kFillTilesY vertically adjacent 128 x 8 tiles: neighbouring tiles share most of their triangles, so
// the table rows fetched for one tile are L1 hits for the next (CTAs are handed to SMs round-robin, so ACROSS CTAs
// there is no such reuse).
template <bool kScores, bool kMask, int G, int WX>
__global__ void __launch_bounds__(kFillThreads, (G == 8 ? 2 : 4) * 256 / kFillThreads)
inverse_fill_kernel(const uint16_t* __restrict__ loc, const TriRec* __restrict__ trirec, const float* __restrict__ table,
                    float* __restrict__ scores, long long* __restrict__ mask, FillParams p) {
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int WL = FOVEA_FILL_WL, kWarpW = 4 * WL, kWarpH = 32 / WL;
  constexpr int kTileH = kWarpH * (kFillThreads / 32 / WX);
  const int x0 = blockIdx.x * (kWarpW * WX) + (warp % WX) * kWarpW + (lane % WL) * 4;
  if (x0 >= p.W) return;
  // the next tile's slice of `loc` (the one DRAM read at the head of a tile's dependent loads) is requested before this
  // tile's channel loop
  const int ybase = blockIdx.y * kFillTilesY * kTileH + (warp / WX) * kWarpH + (lane / WL);
  if (ybase >= p.H) return;
  uint2 l2 = fill_loc(loc, p, b, x0, ybase);
  for (int ty = 0; ty < kFillTilesY; ++ty) {
    const int y = ybase + ty * kTileH;
    const uint2 cur = l2;
    const bool more = ty + 1 < kFillTilesY && y + kTileH < p.H;
    if (more) l2 = fill_loc(loc, p, b, x0, y + kTileH);
    fill_tile<kScores, kMask, G>(loc, trirec, table, scores, mask, p, b, x0, y, cur);
    if (!more) return;
  }
}

// Diagnostic: the store pattern of inverse_fill with no computation (same tiling, same 128-bit streaming stores,
// one 4-pixel store per channel plane) -- the practical write-only ceiling the fill kernel is measured against.
__global__ void __launch_bounds__(kFillThreads)
store_ceiling_kernel(float* __restrict__ scores, const int4* __restrict__ side_read, int C, int H, int W) {
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int WL = FOVEA_FILL_WL, WX = 2;  // the default mapping of inverse_fill_kernel
  static_assert(4 * WL * WX == kFillTileW && (32 / WL) * (kFillThreads / 32 / WX) == kFillTileH,
                "the probe's launch grid assumes the default 128 x 8 CTA tile");
  const int x0 = blockIdx.x * kFillTileW + (warp % WX) * (4 * WL) + (lane % WL) * 4;
  const int y = blockIdx.y * kFillTileH + (warp / WX) * (32 / WL) + (lane / WL);
  if (x0 >= W || y >= H) return;
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t pix = static_cast<size_t>(y) * W + x0;
  float* o = scores + static_cast<size_t>(b) * C * plane + pix;
  float f = static_cast<float>(lane);
  if (side_read) {  // the 2-byte-per-pixel read stream the fill kernel carries (its `loc` map)
    const int2 l = __ldcs(reinterpret_cast<const int2*>(side_read) + (static_cast<size_t>(b) * plane + pix) / 4);
    f += static_cast<float>(l.x ^ l.y);
  }
  for (int c = 0; c < C; ++c, o += plane) __stcs(reinterpret_cast<float4*>(o), make_float4(f, f + 1.f, f + 2.f, f + 3.f));
}

// The same probe with the stores routed through shared memory and TMA (cp.async.bulk.tensor): every thread parks its
// 4 pixels of 4 channel planes with st.shared.v4, one elected thread issues one 128 x 8 tile store per plane.
// kStages tiles of 4 planes (16 KB each) are in flight per CTA.
constexpr int kTmaG = 4;
template <int kStages>
__global__ void __launch_bounds__(kFillThreads)
store_ceiling_tma_kernel(const __grid_constant__ CUtensorMap tmap, const int4* __restrict__ side_read, int C, int H,
                         int W) {
  extern __shared__ __align__(128) float tma_stage[];  // [kStages][kTmaG][kFillTileH][kFillTileW]
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int WL = FOVEA_FILL_WL, WX = 2;
  const int tx = (warp % WX) * (4 * WL) + (lane % WL) * 4, ty = (warp / WX) * (32 / WL) + (lane / WL);
  const int x0 = blockIdx.x * kFillTileW + tx, y = blockIdx.y * kFillTileH + ty;
  const bool live = x0 < W && y < H;
  const size_t plane = static_cast<size_t>(H) * W;
  float f = static_cast<float>(lane);
  if (side_read && live) {
    const int2 l = __ldcs(reinterpret_cast<const int2*>(side_read) + (static_cast<size_t>(b) * plane +
                                                                       static_cast<size_t>(y) * W + x0) / 4);
    f += static_cast<float>(l.x ^ l.y);
  }
  constexpr int kPlaneFloats = kFillTileH * kFillTileW;
  int stage = 0;
  for (int c = 0; c < C; c += kTmaG) {
    float* buf = tma_stage + stage * (kTmaG * kPlaneFloats);
#pragma unroll
    for (int e = 0; e < kTmaG; ++e)
      *reinterpret_cast<float4*>(buf + e * kPlaneFloats + ty * kFillTileW + tx) = make_float4(f, f + 1.f, f + 2.f, f + 3.f);
    fence_proxy_async();
    if (threadIdx.x == 0) tma_wait_read<kStages - 2>();  // the tile the NEXT iteration overwrites has been read
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int e = 0; e < kTmaG; ++e)
        if (c + e < C) tma_store_tile(&tmap, buf + e * kPlaneFloats, blockIdx.x * kFillTileW, blockIdx.y * kFillTileH, b * C + c + e);
      tma_commit();
    }
    stage = stage + 1 == kStages ? 0 : stage + 1;
  }
  if (threadIdx.x == 0) tma_wait_read<0>();
}

// mask[b,p] = lut[b][labels[b,p]]: widens the uint8 output of a reduced-channel fill (fovea.ops.inverse_mask_c1) to the
// reference's int64 class ids.  4 pixels per thread: one 32-bit load, two 128-bit streaming stores.
__global__ void __launch_bounds__(256)
relabel_mask_kernel(const uchar4* __restrict__ labels, const long long* __restrict__ lut, long long quads, int nl,
                    longlong2* __restrict__ mask) {
  const int b = blockIdx.y;
  const long long* l = lut + static_cast<size_t>(b) * nl;
  const uchar4* in = labels + static_cast<size_t>(b) * quads;
  longlong2* out = mask + static_cast<size_t>(b) * quads * 2;
  for (long long q = blockIdx.x * 256ll + threadIdx.x; q < quads; q += 256ll * gridDim.x) {
    const uchar4 v = __ldcs(in + q);
    __stcs(out + 2 * q, make_longlong2(__ldg(l + min(static_cast<int>(v.x), nl - 1)), __ldg(l + min(static_cast<int>(v.y), nl - 1))));
    __stcs(out + 2 * q + 1, make_longlong2(__ldg(l + min(static_cast<int>(v.z), nl - 1)), __ldg(l + min(static_cast<int>(v.w), nl - 1))));
  }
}

// torch.argmax(scores, dim=1) as a stand-alone streaming pass
__global__ void __launch_bounds__(256)
argmax_classes_kernel(const float* __restrict__ scores, long long* __restrict__ mask, int C, long long HW,
                      long long total) {
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = idx / HW, pix = idx - b * HW;
    const float* s = scores + b * C * HW + pix;
    float best = __ldcs(s);
    int bi = 0;
    bool bn = !(best == best);
    for (int c = 1; c < C; ++c) {
      const float v = __ldcs(s + static_cast<long long>(c) * HW);
      const bool isn = !(v == v);
      if (!bn && (isn || v > best)) { best = v; bi = c; bn = isn; }
    }
    mask[idx] = bi;
  }
}

}  // namespace fovea

using namespace fovea;

extern "C" int fovea_grid_inv_scatter(const float* grid, int B, int h, int w, int H, int W, int32_t* winner,
                                      fovea_stream_t stream) {
  FOVEA_REQUIRE(grid && winner && B > 0 && h > 0 && w > 0 && H > 1 && W > 1, "fovea_grid_inv_scatter: bad arguments");
  FOVEA_REQUIRE(H < 65536 && W < 65536, "fovea_grid_inv_scatter: canvas side must be < 65536");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FOVEA_CUDA(cudaMemsetAsync(winner, 0xFF, sizeof(int32_t) * static_cast<size_t>(B) * H * W, s));
  const int total = B * h * w;
  grid_inv_scatter_kernel<<<min(ceil_div(total, 256), kNumSMs * 8), 256, 0, s>>>(
      reinterpret_cast<const float2*>(grid), winner, B, h * w, H, W);
  return check_launch("fovea_grid_inv_scatter");
}

extern "C" int fovea_grid_inv_canvas(const int32_t* winner, int B, int h, int w, int H, int W, float* grid_inv,
                                     fovea_stream_t stream) {
  FOVEA_REQUIRE(winner && grid_inv && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "fovea_grid_inv_canvas: bad arguments");
  const long long total = static_cast<long long>(B) * H * W;
  const long long blocks = (total + 255) / 256;
  grid_inv_canvas_kernel<<<static_cast<int>(blocks < kNumSMs * 32 ? blocks : kNumSMs * 32), 256, 0,
                           static_cast<cudaStream_t>(stream)>>>(winner, reinterpret_cast<float2*>(grid_inv), total, h, w);
  return check_launch("fovea_grid_inv_canvas");
}

extern "C" int fovea_box4_table(const float* pred, int B, int C, int h, int w, int Cs, float* table,
                                fovea_stream_t stream) {
  FOVEA_REQUIRE(pred && table && B > 0 && C > 0 && h > 0 && w > 0, "fovea_box4_table: bad arguments");
  FOVEA_REQUIRE(Cs >= C && Cs % 4 == 0, "fovea_box4_table: Cs=%d must be a multiple of 4 and >= C=%d", Cs, C);
  dim3 grid(ceil_div(h * w, kTabNodes), B);
  box4_table_kernel<true><<<grid, kTabThreads, 0, static_cast<cudaStream_t>(stream)>>>(pred, table, C, Cs, h, w);
  return check_launch("fovea_box4_table");
}

extern "C" int fovea_node_table(const float* values, int B, int C, int h, int w, int Cs, float* table,
                                fovea_stream_t stream) {
  FOVEA_REQUIRE(values && table && B > 0 && C > 0 && h > 0 && w > 0, "fovea_node_table: bad arguments");
  FOVEA_REQUIRE(Cs >= C && Cs % 4 == 0, "fovea_node_table: Cs=%d must be a multiple of 4 and >= C=%d", Cs, C);
  dim3 grid(ceil_div(h * w, kTabNodes), B);
  box4_table_kernel<false><<<grid, kTabThreads, 0, static_cast<cudaStream_t>(stream)>>>(values, table, C, Cs, h, w);
  return check_launch("fovea_node_table");
}

extern "C" int fovea_scatter_nodes(const int64_t* coords, int B, int h, int w, int H, int W, int32_t* winner,
                                   fovea_stream_t stream) {
  FOVEA_REQUIRE(coords && winner && B > 0 && h > 0 && w > 0 && H > 0 && W > 0, "fovea_scatter_nodes: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FOVEA_CUDA(cudaMemsetAsync(winner, 0xFF, sizeof(int32_t) * static_cast<size_t>(B) * H * W, s));
  const int total = B * h * w;
  scatter_nodes_kernel<<<min(ceil_div(total, 256), kNumSMs * 8), 256, 0, s>>>(
      reinterpret_cast<const long long*>(coords), winner, B, h * w, H, W);
  return check_launch("fovea_scatter_nodes");
}

template <bool kNB>
static int launch_select(const float* grid, const int32_t* winner, int B, int h, int w, int H, int W, int nchan, int cap,
                         int32_t* pts, int32_t* src, int32_t* npts, fovea_stream_t stream, const char* who) {
  FOVEA_REQUIRE(grid && winner && pts && src && npts, "%s: null pointer", who);
  FOVEA_REQUIRE(B > 0 && h > 0 && w > 0 && H > 1 && W > 1 && nchan > 0, "%s: bad sizes", who);
  if (cap < h * w + 4 || cap > kSelMax) {
    set_error("%s: cap=%d must satisfy h*w+4=%d <= cap <= %d", who, cap, h * w + 4, kSelMax);
    return FOVEA_ERR_CAPACITY;
  }
  SelectParams p;
  if (int rc = make_select_params(p, h, w, H, W, nchan, cap, who)) return rc;
  int n2 = 1;
  while (n2 < cap) n2 <<= 1;
  const int smem = n2 * static_cast<int>(sizeof(unsigned long long));
  FOVEA_CUDA(cudaFuncSetAttribute(select_points_kernel<kNB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  select_points_kernel<kNB><<<B, kSelThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(grid), winner, pts, src, npts, p);
  return check_launch(who);
}

extern "C" int fovea_select_points(const float* grid, const int32_t* winner, int B, int h, int w, int H, int W,
                                   int nchan, int cap, int32_t* pts, int32_t* src, int32_t* npts,
                                   fovea_stream_t stream) {
  return launch_select<false>(grid, winner, B, h, w, H, W, nchan, cap, pts, src, npts, stream, "fovea_select_points");
}

extern "C" int fovea_select_points_sparse(const float* grid, int B, int h, int w, int H, int W, int nchan, int cap,
                                          int32_t* pts, int32_t* src, int32_t* npts, int32_t* targets,
                                          fovea_stream_t stream) {
  const char* who = "fovea_select_points_sparse";
  FOVEA_REQUIRE(grid && pts && src && npts && targets, "%s: null pointer", who);
  FOVEA_REQUIRE(B > 0 && h > 0 && w > 0 && H > 1 && W > 1 && nchan > 0, "%s: bad sizes", who);
  FOVEA_REQUIRE(H <= 32768 && W <= 65536, "%s: pixel keys are (row << 16 | column) in 31 bits", who);
  if (cap < h * w + 4 || cap > kSelMax) {
    set_error("%s: cap=%d must satisfy h*w+4=%d <= cap <= %d", who, cap, h * w + 4, kSelMax);
    return FOVEA_ERR_CAPACITY;
  }
  SelectParams p;
  if (int rc = make_select_params(p, h, w, H, W, nchan, cap, who)) return rc;
  int n2 = 1;
  while (n2 < h * w + 4) n2 <<= 1;
  const int smem = n2 * static_cast<int>(sizeof(unsigned long long));
  FOVEA_CUDA(cudaFuncSetAttribute(select_points_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  select_points_sparse_kernel<<<B, kSelThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(grid), pts, src, npts, targets, p);
  return check_launch(who);
}

extern "C" int fovea_select_points_nb(const float* grid, const int32_t* winner, int B, int h, int w, int H, int W,
                                      int nchan, int cap, int32_t* pts, int32_t* src, int32_t* npts,
                                      fovea_stream_t stream) {
  return launch_select<true>(grid, winner, B, h, w, H, W, nchan, cap, pts, src, npts, stream, "fovea_select_points_nb");
}

extern "C" int64_t fovea_locate_hints_workspace_bytes(int B, int H, int W) {
  return static_cast<int64_t>(B) * ceil_div(ceil_div(H, FOVEA_HINT_CELL_H), 4) * ceil_div(W, FOVEA_HINT_CELL_W) * 4;
}

extern "C" int fovea_locate_hints(const int32_t* pts, const int32_t* npts, const uint16_t* mesh, const int32_t* ntri,
                                  int B, int cap, int tcap, int H, int W, int32_t* hints, void* workspace,
                                  fovea_stream_t stream) {
  FOVEA_REQUIRE(pts && npts && mesh && ntri && hints && workspace && B > 0 && H > 0 && W > 0,
                "fovea_locate_hints: bad arguments");
  const size_t smem = static_cast<size_t>(tcap) * sizeof(uint4) + static_cast<size_t>(cap) * 4;
  if (smem > 231 * 1024) {
    set_error("fovea_locate_hints: tcap=%d needs %zu B of shared memory", tcap, smem);
    return FOVEA_ERR_CAPACITY;
  }
  FOVEA_CUDA(cudaFuncSetAttribute(locate_hints_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  locate_hints_kernel<<<B, kHintThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      pts, npts, reinterpret_cast<const uint4*>(mesh), ntri, cap, tcap, H, W, hints, static_cast<int32_t*>(workspace));
  return check_launch("fovea_locate_hints");
}

extern "C" int fovea_triangle_setup(const int32_t* pts, const int32_t* src, const uint16_t* mesh, const int32_t* ntri,
                                    int B, int cap, int tcap, int max_coord, int nan_row, void* trirec,
                                    fovea_stream_t stream) {
  FOVEA_REQUIRE(pts && src && mesh && ntri && trirec && B > 0 && cap > 0 && tcap > 0, "fovea_triangle_setup: bad arguments");
  FOVEA_REQUIRE(max_coord > 0 && max_coord <= 16384,
                "fovea_triangle_setup: coordinates must be < 16384 for exact int32 edge functions (got %d)", max_coord);
  FOVEA_REQUIRE(B <= 65535, "fovea_triangle_setup: B too large for the grid");
  dim3 grid(ceil_div(tcap, 256), B);
  triangle_setup_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pts, src, reinterpret_cast<const uint4*>(mesh), ntri, static_cast<TriRec*>(trirec), cap, tcap, nan_row);
  return check_launch("fovea_triangle_setup");
}

extern "C" int fovea_locate_pixels(const int32_t* winner, const void* trirec, const int32_t* ntri,
                                   const int32_t* hints, int B, int h, int w, int H, int W, int tcap, uint16_t* loc,
                                   fovea_stream_t stream) {
  FOVEA_REQUIRE(winner && trirec && ntri && hints && loc, "fovea_locate_pixels: null pointer");
  FOVEA_REQUIRE(B > 0 && h > 0 && w > 0 && H > 1 && W > 1, "fovea_locate_pixels: bad sizes");
  FOVEA_REQUIRE(H <= 16384 && W <= 16384, "fovea_locate_pixels: canvas side must be <= 16384");
  FOVEA_REQUIRE(tcap <= 32768 && h * w < 32767, "fovea_locate_pixels: triangle ids / table rows must fit 15 bits");
  FOVEA_REQUIRE(B <= 65535 && ceil_div(H, kLocTileH) <= 65535, "fovea_locate_pixels: B or H too large for the grid");
  dim3 grid(ceil_div(W, kLocTileW), ceil_div(H, kLocTileH), B);
  locate_pixels_kernel<<<grid, kLocThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      winner, static_cast<const TriRec*>(trirec), ntri, hints, loc, h * w, H, W, tcap);
  return check_launch("fovea_locate_pixels");
}

// Measured on B200 (64 frames of 1024^2, C = 51): G = 4 (128-bit loads, 64 registers, 4 CTAs/SM) 2.58 ms;
// G = 8 (256-bit loads, 128 registers, 2 CTAs/SM) 2.82 ms -- half the L1 requests, but too few warps to hide the
// L2 round trip of every newly touched 32-byte table sector.  FOVEA_FILL_G=8 selects the wide path for A/B runs.
static bool fill_wide_requested() {
  static const bool v = [] { const char* e = getenv("FOVEA_FILL_G"); return e && e[0] == '8'; }();
  return v;
}

template <int G, int WX>
static int launch_fill(const uint16_t* loc, const TriRec* recs, const float* table, float* scores, long long* mk,
                       const FillParams& p, int B, cudaStream_t s) {
  constexpr int kTileH = (32 / FOVEA_FILL_WL) * (kFillThreads / 32 / WX);
  dim3 grid(ceil_div(p.W, 4 * FOVEA_FILL_WL * WX), ceil_div(p.H, kTileH * kFillTilesY), B);
  // Experiment switch FOVEA_FILL_PAD_KB: unused dynamic shared memory per CTA, i.e. a cap on the fill's CTAs per SM (the
  // kernel alone is insensitive to 3 vs 4 CTAs/SM); the registers and thread slots it leaves free let the latency-bound plan
  // kernels of the next batch co-reside under the pipelined schedule instead of displacing fill CTAs.
  static const int pad = [] { const char* e = getenv("FOVEA_FILL_PAD_KB"); return e ? atoi(e) * 1024 : 0; }();
  if (scores && mk) {
    if (pad) cudaFuncSetAttribute(inverse_fill_kernel<true, true, G, WX>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
    inverse_fill_kernel<true, true, G, WX><<<grid, kFillThreads, pad, s>>>(loc, recs, table, scores, mk, p);
  } else if (scores) {
    if (pad) cudaFuncSetAttribute(inverse_fill_kernel<true, false, G, WX>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
    inverse_fill_kernel<true, false, G, WX><<<grid, kFillThreads, pad, s>>>(loc, recs, table, scores, mk, p);
  } else {
    inverse_fill_kernel<false, true, G, WX><<<grid, kFillThreads, 0, s>>>(loc, recs, table, scores, mk, p);
  }
  return check_launch("fovea_inverse_fill");
}

int fovea_launch_fill_smem(const uint16_t* loc, const void* trirec, const float* table, int B, int C, int Cs, int h, int w,
                           int H, int W, int tcap, int zero_residual, float* scores, int mode, cudaStream_t s);   // inverse_smem.cu

extern "C" int fovea_inverse_fill(const uint16_t* loc, const void* trirec, const float* table, int B, int C, int Cs, int h,
                                  int w, int H, int W, int tcap, int zero_residual, float* scores, void* mask,
                                  int mask_u8, fovea_stream_t stream) {
  FOVEA_REQUIRE(loc && trirec && table, "fovea_inverse_fill: null pointer");
  FOVEA_REQUIRE(scores || mask, "fovea_inverse_fill: neither scores nor mask requested");
  FOVEA_REQUIRE(B > 0 && C > 0 && Cs >= C && Cs % 4 == 0 && h > 0 && w > 0 && H > 1 && W > 1,
                "fovea_inverse_fill: bad sizes");
  FOVEA_REQUIRE(W % 4 == 0, "fovea_inverse_fill: W=%d must be a multiple of 4 (128-bit stores)", W);
  FOVEA_REQUIRE(H <= 16384 && W <= 16384, "fovea_inverse_fill: canvas side must be <= 16384");
  FOVEA_REQUIRE(B <= 65535 && ceil_div(H, 4) <= 65535, "fovea_inverse_fill: B or H too large for the grid");
  FOVEA_REQUIRE(static_cast<long long>(h) * w + 2 <= 32768 && static_cast<long long>(h) * w * Cs * 4 < (1ll << 32),
                "fovea_inverse_fill: value table too large (rows must fit 15 bits, bytes 32 bits)");
  FOVEA_REQUIRE(!mask_u8 || C <= 256, "fovea_inverse_fill: uint8 masks need C <= 256 (C=%d)", C);
  FillParams p{C, Cs, h, w, H, W, 0, tcap, zero_residual, mask_u8 ? 1 : 0};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const TriRec* recs = static_cast<const TriRec*>(trirec);
  long long* mk = reinterpret_cast<long long*>(mask);
  // experiment switch FOVEA_FILL_SMEM=1: scores-only fills stage the tile's table rows in shared memory (inverse_smem.cu)
  const char* smem_env = getenv("FOVEA_FILL_SMEM");   // (read per call: the A/B test flips it between launches)
  const int smem_rows = smem_env ? atoi(smem_env) : 0;   // 1: rows in shared memory; 2: + TMA tile stores
  if (smem_rows && scores && !mask) {
    const int rc = fovea_launch_fill_smem(loc, trirec, table, B, C, Cs, h, w, H, W, tcap, zero_residual, scores, smem_rows, s);
    if (rc >= 0) return rc;
  }
  // 8-channel groups (256-bit table loads) need 32-byte aligned rows: Cs % 8 == 0 and a 32-byte aligned table
  const bool wide = Cs % 8 == 0 && (reinterpret_cast<uintptr_t>(table) & 31u) == 0 && fill_wide_requested();
  if (wide) return launch_fill<8, 2>(loc, recs, table, scores, mk, p, B, s);
  return W >= 128 ? launch_fill<4, 2>(loc, recs, table, scores, mk, p, B, s)
                  : launch_fill<4, 1>(loc, recs, table, scores, mk, p, B, s);
}

extern "C" int fovea_probe_store_ceiling(float* scores, const int32_t* side_read, int B, int C, int H, int W,
                                         fovea_stream_t stream) {
  FOVEA_REQUIRE(scores && B > 0 && C > 0 && H > 0 && W > 0 && W % 4 == 0, "fovea_probe_store_ceiling: bad arguments");
  FOVEA_REQUIRE(B <= 65535 && ceil_div(H, kFillTileH) <= 65535, "fovea_probe_store_ceiling: B or H too large");
  dim3 grid(ceil_div(W, kFillTileW), ceil_div(H, kFillTileH), B);
  static const int tma_stages = [] { const char* e = getenv("FOVEA_PROBE_TMA"); return e ? atoi(e) : 0; }();
  if (tma_stages >= 2 && W >= kFillTileW) {  // experiment: the same stores through shared memory + TMA
    CUtensorMap map;
    if (int rc = make_plane_store_map(&map, scores, static_cast<long long>(B) * C, H, W, kFillTileW, kFillTileH)) return rc;
    const int st = tma_stages >= 3 ? 3 : 2;
    const size_t smem = static_cast<size_t>(st) * kTmaG * kFillTileH * kFillTileW * sizeof(float);
    auto kern = st == 3 ? store_ceiling_tma_kernel<3> : store_ceiling_tma_kernel<2>;
    FOVEA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, kFillThreads, smem, static_cast<cudaStream_t>(stream)>>>(map, reinterpret_cast<const int4*>(side_read), C, H, W);
    return check_launch("fovea_probe_store_ceiling (tma)");
  }
  store_ceiling_kernel<<<grid, kFillThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      scores, reinterpret_cast<const int4*>(side_read), C, H, W);
  return check_launch("fovea_probe_store_ceiling");
}

extern "C" int fovea_relabel_mask(const uint8_t* labels, const int64_t* lut, int B, int64_t HW, int nl, int64_t* mask,
                                  fovea_stream_t stream) {
  FOVEA_REQUIRE(labels && lut && mask && B > 0 && HW > 0 && nl > 0 && nl <= 256, "fovea_relabel_mask: bad arguments");
  FOVEA_REQUIRE(HW % 4 == 0 && B <= 65535, "fovea_relabel_mask: H*W must be a multiple of 4 and B <= 65535");
  const long long quads = HW / 4;
  const int bx = static_cast<int>(quads < 256ll * kNumSMs * 8 ? (quads + 255) / 256 : kNumSMs * 8);
  relabel_mask_kernel<<<dim3(bx, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uchar4*>(labels), reinterpret_cast<const long long*>(lut), quads, nl,
      reinterpret_cast<longlong2*>(mask));
  return check_launch("fovea_relabel_mask");
}

extern "C" int fovea_argmax_classes(const float* scores, int B, int C, int64_t HW, int64_t* mask,
                                    fovea_stream_t stream) {
  FOVEA_REQUIRE(scores && mask && B > 0 && C > 0 && HW > 0, "fovea_argmax_classes: bad arguments");
  const long long total = static_cast<long long>(B) * HW;
  const long long blocks = (total + 255) / 256;
  argmax_classes_kernel<<<static_cast<int>(blocks < kNumSMs * 32 ? blocks : kNumSMs * 32), 256, 0,
                          static_cast<cudaStream_t>(stream)>>>(scores, reinterpret_cast<long long*>(mask), C, HW, total);
  return check_launch("fovea_argmax_classes");
}

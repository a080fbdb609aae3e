// Point location by RASTERISATION: find_simplex for every pixel of the canvas (interp2d.py:58) computed triangle by
// triangle instead of pixel by pixel.
//
// fovea_locate_pixels walks the mesh from every 32-pixel run of every row: 360 M warp instructions per 64 frames of
// 1024^2 (0.39 ms, issue-bound; plus the walk-start hints another kernel has to prepare).  Here one warp takes one
// triangle: lane l owns rows ymin + l, ymin + l + 32, ... of its bounding box, solves the three edge inequalities
// e_i(y, x) = A_i y + B_i x + C_i >= m_i of the setup record for the row's span [lo, hi] in closed form (exact integer
// arithmetic, the SAME predicate the walker tests pixel by pixel, so the two kernels produce the same map) and stores
// the triangle id over the span.  Spans of different triangles are disjoint (the tie rule gives every pixel exactly one
// owner), so there is nothing to synchronise.  The pixels that received a node are stamped afterwards from the
// 6 400 nodes themselves (stamp_nodes_kernel) -- the 4-byte-per-pixel winner map is no longer read by stage 3's locate.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "fill.cuh"

namespace fovea {

constexpr int kRasThreads = 256;
constexpr int kRasMode = 64;       // default of FOVEA_RAS_MODE (see fovea_locate_raster)
constexpr int kRasTileMax = 1024;   // bounding boxes up to this many pixels are swept pixel by pixel (measured 32 .. 2048: flat,
                                    // 356 us at 1024 against 378 at 2048 and 396 at 320; FOVEA_RAS_TILE_MAX overrides)

// floor(a / b) for b > 0 and |a / b| < 2^22: float quotient, exact fix-up
__device__ __forceinline__ int floor_div_pos(int a, int b) {
  int q = __float2int_rd(__fdividef(static_cast<float>(a), static_cast<float>(b)));
  const long long r = static_cast<long long>(a) - static_cast<long long>(q) * b;
  if (r < 0) --q;
  else if (r >= b) ++q;
  return q;
}

__device__ __forceinline__ void store_span(uint16_t* row, int lo, int hi, unsigned id) {
  int x = lo;
  const unsigned v2 = id | (id << 16);
  while (x <= hi && (x & 7)) row[x++] = static_cast<uint16_t>(id);             // up to the next 16-byte boundary
  for (; x + 7 <= hi; x += 8) *reinterpret_cast<uint4*>(row + x) = make_uint4(v2, v2, v2, v2);
  while (x <= hi) row[x++] = static_cast<uint16_t>(id);
}

// Everything a lane needs about the triangle it works on.
struct RasTri {
  int A0, B0, C0, A1, B1, C1, A2, B2, C2;   // edge functions with the tie bit already subtracted: inside <=> all >= 0
  int ymin, ymax, xmin, xmax;
  int p0, p1, p2;                           // the vertices, packed (row << 16 | column)
  bool live;
};

__device__ __forceinline__ RasTri ras_load(const int32_t* __restrict__ pb, const uint4* __restrict__ mesh,
                                           const TriRec* __restrict__ recs, int t, int T) {
  RasTri R;
  R.live = t < T;
  R.A0 = R.B0 = R.C0 = R.A1 = R.B1 = R.C1 = R.A2 = R.B2 = R.C2 = 0;
  R.ymin = R.xmin = 0; R.ymax = R.xmax = -1;
  R.p0 = R.p1 = R.p2 = -1;
  if (!R.live) return R;
  const uint4* r = reinterpret_cast<const uint4*>(recs + t);
  const uint4 q0 = __ldg(r), q1 = __ldg(r + 1), q2 = __ldg(r + 2);
  const uint4 mq = __ldg(mesh + t);
  if (q2.w == 0u) { R.live = false; return R; }  // degenerate triangle (host meshes only): owns nothing
  const int p0 = __ldg(pb + (mq.x & 0xFFFFu)), p1 = __ldg(pb + (mq.x >> 16)), p2 = __ldg(pb + (mq.y & 0xFFFFu));
  R.p0 = p0; R.p1 = p1; R.p2 = p2;
  R.ymin = min(min(p0 >> 16, p1 >> 16), p2 >> 16); R.ymax = max(max(p0 >> 16, p1 >> 16), p2 >> 16);
  R.xmin = min(min(p0 & 0xFFFF, p1 & 0xFFFF), p2 & 0xFFFF); R.xmax = max(max(p0 & 0xFFFF, p1 & 0xFFFF), p2 & 0xFFFF);
  const unsigned m = q2.z >> 16;
  R.A0 = static_cast<int>(q0.x); R.B0 = static_cast<int>(q0.y); R.C0 = static_cast<int>(q0.z) - static_cast<int>(m & 1u);
  R.A1 = static_cast<int>(q0.w); R.B1 = static_cast<int>(q1.x); R.C1 = static_cast<int>(q1.y) - static_cast<int>((m >> 1) & 1u);
  R.A2 = static_cast<int>(q1.z); R.B2 = static_cast<int>(q1.w); R.C2 = static_cast<int>(q2.x) - static_cast<int>((m >> 2) & 1u);
  return R;
}

// One warp takes FOUR triangles, eight lanes each (the typical triangle covers ~80 pixels in a ~13 x 13 box: a whole warp
// per triangle leaves most lanes idle and pays the set-up 4 times over -- measured 290 instructions per triangle).  The
// eight lanes sweep their triangle's bounding box row by row, 8 pixels per step, each lane testing ITS pixel against the
// three edge functions (no divisions).  Triangles with a large box (the few hull / periphery triangles that cover a tenth
// of the canvas each) are then taken one after the other by the WHOLE warp on the row-span path: lane l owns rows
// ymin + l, + 32, ..., solves the three inequalities for the row's span in closed form and stores it 16 bytes at a time.
__global__ void __launch_bounds__(kRasThreads)
raster_locate_kernel(const int32_t* __restrict__ pts, const uint4* __restrict__ mesh, const TriRec* __restrict__ trirec,
                     const int32_t* __restrict__ ntri, uint16_t* __restrict__ loc, int hw, int H, int W, int cap, int tcap,
                     int tile_max) {
  const int b = blockIdx.y;
  const int T = ntri[b];
  uint16_t* lb = loc + static_cast<size_t>(b) * H * W;
  if (T <= 0) {  // no mesh for this frame (see fovea_delaunay): nothing owns anything
    const unsigned none = 0x8000u | static_cast<unsigned>(hw);
    for (size_t i = static_cast<size_t>(blockIdx.x) * kRasThreads + threadIdx.x; i < static_cast<size_t>(H) * W;
         i += static_cast<size_t>(gridDim.x) * kRasThreads)
      lb[i] = static_cast<uint16_t>(none);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = (blockIdx.x * (kRasThreads / 32) + warp) * 4;   // first of this warp's four triangles
  if (t0 >= T) return;
  const int sub = lane >> 3, sl = lane & 7;
  const int t = t0 + sub;
  const RasTri R = ras_load(pts + static_cast<size_t>(b) * cap, mesh + static_cast<size_t>(b) * tcap,
                            trirec + static_cast<size_t>(b) * tcap, t, T);
  const int bh = R.ymax - R.ymin + 1, bw = R.xmax - R.xmin + 1;
  const bool large = R.live && bh * bw > tile_max;
  if (R.live && !large) {
    int e0 = R.A0 * R.ymin + R.B0 * (R.xmin + sl) + R.C0;
    int e1 = R.A1 * R.ymin + R.B1 * (R.xmin + sl) + R.C1;
    int e2 = R.A2 * R.ymin + R.B2 * (R.xmin + sl) + R.C2;
    for (int y = R.ymin; y <= R.ymax; ++y) {
      int f0 = e0, f1 = e1, f2 = e2;
      uint16_t* row = lb + static_cast<size_t>(y) * W;
      for (int x = R.xmin + sl; x <= R.xmax; x += 8) {
        if ((f0 | f1 | f2) >= 0) row[x] = static_cast<uint16_t>(t);    // all three >= 0  <=>  no sign bit set
        f0 += 8 * R.B0; f1 += 8 * R.B1; f2 += 8 * R.B2;
      }
      e0 += R.A0; e1 += R.A1; e2 += R.A2;
    }
  }
  // the large ones: whole warp, one after the other (their parameters come from the first lane of their group)
  unsigned todo = __ballot_sync(0xffffffffu, large) & 0x01010101u;
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const int A[3] = {__shfl_sync(0xffffffffu, R.A0, src), __shfl_sync(0xffffffffu, R.A1, src), __shfl_sync(0xffffffffu, R.A2, src)};
    const int Bx[3] = {__shfl_sync(0xffffffffu, R.B0, src), __shfl_sync(0xffffffffu, R.B1, src), __shfl_sync(0xffffffffu, R.B2, src)};
    const int Cc[3] = {__shfl_sync(0xffffffffu, R.C0, src), __shfl_sync(0xffffffffu, R.C1, src), __shfl_sync(0xffffffffu, R.C2, src)};
    const int ymin = __shfl_sync(0xffffffffu, R.ymin, src), ymax = __shfl_sync(0xffffffffu, R.ymax, src);
    const int xmin = __shfl_sync(0xffffffffu, R.xmin, src), xmax = __shfl_sync(0xffffffffu, R.xmax, src);
    const unsigned id = static_cast<unsigned>(t0 + (src >> 3));
    for (int y = ymin + lane; y <= ymax; y += 32) {
      int lo = xmin, hi = xmax;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int k = A[i] * y + Cc[i];                                       // the pixel is inside iff B x + k >= 0
        if (Bx[i] > 0) lo = max(lo, -floor_div_pos(k, Bx[i]));                // x >= ceil(-k / B) = -floor(k / B)
        else if (Bx[i] < 0) hi = min(hi, floor_div_pos(k, -Bx[i]));           // x <= floor(k / -B)
        else if (k < 0) hi = -1;                                              // the whole row is outside
      }
      if (lo <= hi) store_span(lb + static_cast<size_t>(y) * W, lo, hi, id);
    }
  }
}

// The same map with NO per-pixel tests: every triangle is handled like the large ones above -- a lane takes a row of the
// bounding box, solves the three edge inequalities for the row's span in closed form and stores it -- only with LPT
// lanes per triangle (rows ymin + l, + LPT, ...) so that the typical 13-row triangle keeps its lanes busy.  Per
// (triangle, row): three float-quotient divisions with exact fix-up and ~4 stores, against 2 x 8 lane-steps of the sweep.
__device__ __forceinline__ void store_span_short(uint16_t* row, int lo, int hi, unsigned id) {
  if (hi - lo >= 23) { store_span(row, lo, hi, id); return; }
  int x = lo;
  if (x & 1) row[x++] = static_cast<uint16_t>(id);
  const unsigned v2 = id | (id << 16);
  for (; x < hi; x += 2) *reinterpret_cast<unsigned*>(row + x) = v2;
  if (x == hi) row[x] = static_cast<uint16_t>(id);
}

template <int LPT>
__global__ void __launch_bounds__(kRasThreads)
raster_span_kernel(const int32_t* __restrict__ pts, const uint4* __restrict__ mesh, const TriRec* __restrict__ trirec,
                   const int32_t* __restrict__ ntri, uint16_t* __restrict__ loc, int hw, int H, int W, int cap, int tcap,
                   int tile_max) {
  constexpr int TPW = 32 / LPT;   // triangles per warp
  const int b = blockIdx.y;
  const int T = ntri[b];
  uint16_t* lb = loc + static_cast<size_t>(b) * H * W;
  if (T <= 0) {  // no mesh for this frame (see fovea_delaunay): nothing owns anything
    const unsigned none = 0x8000u | static_cast<unsigned>(hw);
    for (size_t i = static_cast<size_t>(blockIdx.x) * kRasThreads + threadIdx.x; i < static_cast<size_t>(H) * W;
         i += static_cast<size_t>(gridDim.x) * kRasThreads)
      lb[i] = static_cast<uint16_t>(none);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = (blockIdx.x * (kRasThreads / 32) + warp) * TPW;
  if (t0 >= T) return;
  const int sub = lane / LPT, sl = lane % LPT;
  const int t = t0 + sub;
  const RasTri R = ras_load(pts + static_cast<size_t>(b) * cap, mesh + static_cast<size_t>(b) * tcap,
                            trirec + static_cast<size_t>(b) * tcap, t, T);
  const int bh = R.ymax - R.ymin + 1, bw = R.xmax - R.xmin + 1;
  const bool large = R.live && bh * bw > tile_max;
  auto row_span = [&](const int (&A)[3], const int (&Bx)[3], const int (&Cc)[3], int y, int& lo, int& hi) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int k = A[i] * y + Cc[i];                                       // the pixel is inside iff B x + k >= 0
      if (Bx[i] > 0) lo = max(lo, -floor_div_pos(k, Bx[i]));                // x >= ceil(-k / B) = -floor(k / B)
      else if (Bx[i] < 0) hi = min(hi, floor_div_pos(k, -Bx[i]));           // x <= floor(k / -B)
      else if (k < 0) hi = -1;                                              // the whole row is outside
    }
  };
  if (R.live && !large) {
    const int A[3] = {R.A0, R.A1, R.A2}, Bx[3] = {R.B0, R.B1, R.B2}, Cc[3] = {R.C0, R.C1, R.C2};
    for (int y = R.ymin + sl; y <= R.ymax; y += LPT) {
      int lo = R.xmin, hi = R.xmax;
      row_span(A, Bx, Cc, y, lo, hi);
      if (lo <= hi) store_span_short(lb + static_cast<size_t>(y) * W, lo, hi, static_cast<unsigned>(t));
    }
  }
  // the large ones: whole warp, one after the other (their parameters come from the first lane of their group)
  unsigned leaders = 0;
#pragma unroll
  for (int g = 0; g < TPW; ++g) leaders |= 1u << (g * LPT);
  unsigned todo = __ballot_sync(0xffffffffu, large) & leaders;
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const int A[3] = {__shfl_sync(0xffffffffu, R.A0, src), __shfl_sync(0xffffffffu, R.A1, src), __shfl_sync(0xffffffffu, R.A2, src)};
    const int Bx[3] = {__shfl_sync(0xffffffffu, R.B0, src), __shfl_sync(0xffffffffu, R.B1, src), __shfl_sync(0xffffffffu, R.B2, src)};
    const int Cc[3] = {__shfl_sync(0xffffffffu, R.C0, src), __shfl_sync(0xffffffffu, R.C1, src), __shfl_sync(0xffffffffu, R.C2, src)};
    const int ymin = __shfl_sync(0xffffffffu, R.ymin, src), ymax = __shfl_sync(0xffffffffu, R.ymax, src);
    const int xmin = __shfl_sync(0xffffffffu, R.xmin, src), xmax = __shfl_sync(0xffffffffu, R.xmax, src);
    const unsigned id = static_cast<unsigned>(t0 + src / LPT);
    for (int y = ymin + lane; y <= ymax; y += 32) {
      int lo = xmin, hi = xmax;
      row_span(A, Bx, Cc, y, lo, hi);
      if (lo <= hi) store_span(lb + static_cast<size_t>(y) * W, lo, hi, id);
    }
  }
}

// ---- Marker raster.  On a canvas the triangulation covers (forced corners: the convex hull IS the canvas) the spans of
// one row partition it, so the map is fixed by where every span STARTS: the map is cleared to "no start",
// raster_mark_kernel stores the triangle id at the first pixel of each (triangle, row) span, and
// raster_fill_rows_kernel sweeps every row once, carrying the last started id forward, with coalesced 16-byte loads and
// stores.  One lane takes one triangle and walks its rows with an exact integer DDA (quotient + remainder of every
// edge's crossing, advanced by one addition and one conditional carry per row -- two divisions per edge per TRIANGLE
// instead of one per row); ~35 lane-instructions and one 2-byte store per (triangle, row), against ~8 per PIXEL tested by
// the sweep.
constexpr int kMarkCoopRows = 64;   // taller triangles are ALWAYS taken by the whole warp, rows strided by 32 (direct divisions)
#ifndef FOVEA_TALL_CTAS
#define FOVEA_TALL_CTAS 4   // CTAs per SM of raster_mark_tall_kernel's grid (2 / 4 / 8 measured: flat)
#endif
constexpr int kMarkThreads = 128;   // (<= 256: the sorted order is kept in bytes)

__device__ __forceinline__ void mark_start(uint16_t* loc, unsigned lin, unsigned id) {   // lin: pixel index in the chunk (< 2^32)
  loc[lin] = static_cast<uint16_t>(id);
}

// At a vertex the tie rule is not exclusive: several of the triangles around a site can own its pixel (harmless for the
// sweep -- the pixel carries a node and is stamped afterwards -- but two span STARTS on one pixel would lose one).
// A one-pixel span on the triangle's own vertex (an apex) carries no information and is skipped.  That is enough:
// a longer span that starts on a vertex either continues onto pixels nobody else owns (its start is the only one
// kept there), or is a one-pixel-long horizontal edge, whose two triangles tie for a start that only ever reaches the
// edge's two vertex pixels -- both stamped.
__device__ __forceinline__ bool apex_only(const RasTri& R, int y, int lo, int hi) {
  const int pk = (y << 16) | lo;
  return lo == hi && (pk == R.p0 || pk == R.p1 || pk == R.p2);
}

struct EdgeDda {   // floor((A y + C) / m), m = |B| (1 when B == 0), as quotient + remainder, advanced row by row
  int q, r, dq, dr, m;
  // floor(a / m) and the remainder, m > 0: float quotient + one fix-up step, verified -- the remainder is formed modulo
  // 2^32 from the estimate (off by at most |a / m| * 2^-22 <= 2^8 units, i.e. the true remainder is within 2^8 * m < 2^23
  // of [0, m): no wrap) and an estimate that one step does not repair falls back to the integer division.
  static __device__ __forceinline__ void fdiv(int a, int m, int& q, int& r) {
    q = __float2int_rd(__fdividef(static_cast<float>(a), static_cast<float>(m)));
    r = static_cast<int>(static_cast<unsigned>(a) - static_cast<unsigned>(q) * static_cast<unsigned>(m));
    if (r < 0) { r += m; --q; }
    else if (r >= m) { r -= m; ++q; }
    if (r < 0 || r >= m) {   // (far from an edge's end points its crossing can leave the float quotient's exact range)
      q = a / m; r = a - q * m;
      if (r < 0) { r += m; --q; }
    }
  }
  __device__ __forceinline__ void init(int A, int B, int C, int y) {
    m = B > 0 ? B : (B < 0 ? -B : 1);
    const int k = A * y + C;
    if (B == 0) { q = k; r = 0; dq = A; dr = 0; return; }   // (k itself can be ~2^30: no division)
    fdiv(k, m, q, r);
    fdiv(A, m, dq, dr);
  }
  __device__ __forceinline__ void step() {
    q += dq; r += dr;
    if (r >= m) { r -= m; ++q; }
  }
  __device__ __forceinline__ void jump(int n) {   // n rows at once (n * dr < 2^10 * 2^14: the quotient below is <= n, exact)
    int qq, rr;
    fdiv(r + n * dr, m, qq, rr);
    q += n * dq + qq; r = rr;
  }
};

// The three edges of a triangle in roles: a counter-clockwise triangle of positive area has at least one edge that bounds
// its rows from the LEFT (B > 0: x >= -floor(k / B)) and one from the RIGHT (B < 0: x <= floor(k / -B)); the third is
// either kind or horizontal (B == 0: the row is inside iff k >= 0).
struct SpanWalker {
  EdgeDda L, Rr, X;
  int xkind, xmin, xmax;
  __device__ __forceinline__ void init(const RasTri& T, int y) {
    const int iL = T.B0 > 0 ? 0 : (T.B1 > 0 ? 1 : 2);
    const int iR = T.B0 < 0 ? 0 : (T.B1 < 0 ? 1 : 2);
    const int iX = 3 - iL - iR;
    auto pick = [&](int i, int& A, int& B, int& C) {
      A = i == 0 ? T.A0 : (i == 1 ? T.A1 : T.A2); B = i == 0 ? T.B0 : (i == 1 ? T.B1 : T.B2); C = i == 0 ? T.C0 : (i == 1 ? T.C1 : T.C2);
    };
    int A, B, C;
    pick(iL, A, B, C); L.init(A, B, C, y);
    pick(iR, A, B, C); Rr.init(A, B, C, y);
    pick(iX, A, B, C); X.init(A, B, C, y);
    xkind = B > 0 ? 1 : (B < 0 ? -1 : 0);
    xmin = T.xmin; xmax = T.xmax;
  }
  __device__ __forceinline__ void span(int& lo, int& hi) const {
    lo = max(xmin, -L.q); hi = min(xmax, Rr.q);
    lo = max(lo, xkind > 0 ? -X.q : lo);
    hi = min(hi, xkind < 0 ? X.q : (xkind == 0 && X.q < 0 ? -1 : hi));
  }
  __device__ __forceinline__ void step() { L.step(); Rr.step(); X.step(); }
  __device__ __forceinline__ void jump(int n) { L.jump(n); Rr.jump(n); X.jump(n); }
};

__global__ void __launch_bounds__(kMarkThreads)
raster_mark_kernel(const int32_t* __restrict__ pts, const uint4* __restrict__ mesh, const TriRec* __restrict__ trirec,
                   const int32_t* __restrict__ ntri, uint16_t* __restrict__ loc, unsigned* __restrict__ queue, int H,
                   int W, int cap, int tcap, int tall_cost) {
  const int b = blockIdx.y;
  const int T = ntri[b];
  if (T <= 0) return;  // no mesh for this frame: no starts, the row sweep leaves "no value" everywhere
  const int lane = threadIdx.x & 31;
  const int tbase = blockIdx.x * kMarkThreads;
  if (tbase >= T) return;
  // The lanes of a warp walk their triangles in lockstep: hand the CTA's 128 triangles out SORTED BY HEIGHT (a counting
  // sort on the row count, in shared memory), so that a warp's triangles are about equally tall.
  __shared__ int s_hist[kMarkCoopRows + 2];
  __shared__ unsigned char s_order[kMarkThreads];
  __shared__ RasTri s_tri[kMarkThreads];                        // (17 words each: conflict-free)
  if (threadIdx.x < kMarkCoopRows + 2) s_hist[threadIdx.x] = 0;
  // every thread loads the record of ITS triangle (all global loads of the kernel, issued together) ...
  const RasTri mine = ras_load(pts + static_cast<size_t>(b) * cap, mesh + static_cast<size_t>(b) * tcap,
                               trirec + static_cast<size_t>(b) * tcap, tbase + threadIdx.x, T);
  s_tri[threadIdx.x] = mine;
  __syncthreads();
  const bool past = tbase + static_cast<int>(threadIdx.x) >= T;
  const int bin = past ? kMarkCoopRows + 1 : min(max(mine.ymax - mine.ymin, 0), kMarkCoopRows);   // (past the end: last)
  const int pos = atomicAdd(&s_hist[bin], 1);
  __syncthreads();
  if (threadIdx.x < 32) {                                       // exclusive scan of the 66 bins by one warp
    int carry = 0;
    for (int i0 = 0; i0 < kMarkCoopRows + 2; i0 += 32) {
      const int i = i0 + lane;
      const int v = i < kMarkCoopRows + 2 ? s_hist[i] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
      }
      if (i < kMarkCoopRows + 2) s_hist[i] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
  s_order[s_hist[bin] + pos] = static_cast<unsigned char>(threadIdx.x);
  __syncthreads();
  // ... and walks the one the sort hands it
  const int t = tbase + s_order[threadIdx.x];
  const unsigned img = static_cast<unsigned>(b) * H * W;   // (a chunk holds < 2^32 pixels: checked by the launcher)
  const RasTri R = s_tri[s_order[threadIdx.x]];
  // A warp's lanes walk their triangles in lockstep, so the tallest one sets the price (~45 instructions per row for
  // the whole warp) -- while a queued triangle costs a warp of raster_mark_tall_kernel ~300.  The height limit above
  // which triangles leave the lane path is the cheapest of a few candidates under that model.
  const int bh = R.live ? R.ymax - R.ymin + 1 : 0;
  int limit = kMarkCoopRows, best = 0x7fffffff;
#pragma unroll
  for (int cand = kMarkCoopRows; cand >= 8; cand = cand * 3 / 4) {   // 64, 48, 36, 27, 20, 15, 11, 8
    const int above = __popc(__ballot_sync(0xffffffffu, bh > cand));
    const int tallest = __reduce_max_sync(0xffffffffu, bh > cand ? 0 : bh);
    const int cost = tallest * 45 + above * tall_cost;
    if (cost < best) { best = cost; limit = cand; }
  }
  const bool large = R.live && bh > limit;
  if (R.live && !large) {
    SpanWalker w;
    w.init(R, R.ymin);
    unsigned rowoff = img + static_cast<unsigned>(R.ymin) * W;
    for (int y = R.ymin; y <= R.ymax; ++y, rowoff += W) {
      int lo, hi;
      w.span(lo, hi);
      if (lo <= hi && !apex_only(R, y, lo, hi)) mark_start(loc, rowoff + lo, static_cast<unsigned>(t));
      w.step();
    }
  }
  // the tall ones go to a queue: raster_mark_tall_kernel spreads them over the whole device, one warp each (the corner
  // fans -- dozens of consecutive triangles a thousand rows tall -- would otherwise be walked one after the other by
  // the single warp that drew them)
  const unsigned tall = __ballot_sync(0xffffffffu, large);
  if (tall) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(queue, static_cast<unsigned>(__popc(tall)));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (large) queue[4 + base + __popc(tall & ((1u << lane) - 1u))] = (static_cast<unsigned>(b) << 16) | static_cast<unsigned>(t);
  }
}

__global__ void __launch_bounds__(256)
raster_mark_tall_kernel(const int32_t* __restrict__ pts, const uint4* __restrict__ mesh, const TriRec* __restrict__ trirec,
                        const int32_t* __restrict__ ntri, uint16_t* __restrict__ loc, unsigned* __restrict__ queue, int H,
                        int W, int cap, int tcap) {
  const int lane = threadIdx.x & 31;
  const unsigned n = queue[0];
  __shared__ unsigned s_next;
  for (;;) {     // CTAs pull eight items at a time as they finish (a few are 1000 rows tall, most ~100): one atomic per eight
    __syncthreads();                           // (one atomic per item on the one counter was the kernel's bottleneck)
    if (threadIdx.x == 0) s_next = atomicAdd(queue + 1, 8u);
    __syncthreads();
    const unsigned first = s_next;
    if (first >= n) break;
    const unsigned i = first + (threadIdx.x >> 5);
    if (i >= n) continue;
    const unsigned item = queue[4 + i];
    const int b = static_cast<int>(item >> 16), t = static_cast<int>(item & 0xFFFFu);
    const unsigned img = static_cast<unsigned>(b) * H * W;
    const RasTri R = ras_load(pts + static_cast<size_t>(b) * cap, mesh + static_cast<size_t>(b) * tcap,
                              trirec + static_cast<size_t>(b) * tcap, t, ntri[b]);   // (every lane the same triangle)
    // lane l walks rows [ymin + l * per, ymin + (l + 1) * per): the walker is started at ymin by every lane alike (the
    // only divisions with large operands: warp-uniform) and jumps to the lane's first row
    const int per = (R.ymax - R.ymin + 32) / 32;
    const int y0 = R.ymin + lane * per, y1 = min(y0 + per - 1, R.ymax);
    SpanWalker w;
    w.init(R, R.ymin);
    if (y0 <= y1) {
      if (lane) w.jump(lane * per);
      unsigned rowoff = img + static_cast<unsigned>(y0) * W;
      for (int y = y0; y <= y1; ++y, rowoff += W) {
        int lo, hi;
        w.span(lo, hi);
        if (lo <= hi && !apex_only(R, y, lo, hi)) mark_start(loc, rowoff + lo, static_cast<unsigned>(t));
        w.step();
      }
    }
  }
}

// One warp per canvas row: 256 pixels per step (8 per lane, one 16-byte load + store), the id of the last span start
// carried from lane to lane by a 5-step shuffle scan and from step to step in a register.  A pixel that is no span
// start still holds the kNoStart the chunk was cleared to.
constexpr unsigned kNoStart = 0xFFFFu;   // never a map value: node ids and triangle ids are < 32767 (checked on the host)
__global__ void __launch_bounds__(256)
raster_fill_rows_kernel(uint16_t* __restrict__ loc, long long rows, int W, unsigned none) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  uint16_t* lr = loc + row * W;
  unsigned carry = none;
  for (int x0 = 0; x0 < W; x0 += 256) {
    const int x = x0 + lane * 8;
    const bool act = x < W;
    uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
    if (act) v = *reinterpret_cast<const uint4*>(lr + x);
    unsigned px[8] = {v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16, v.z & 0xFFFFu, v.z >> 16, v.w & 0xFFFFu, v.w >> 16};
    unsigned inc = kNoStart;   // the id started last inside my 8 pixels ...
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (px[j] != kNoStart) inc = px[j];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {   // ... then: the last start at or before my pixels (within this step)
      const unsigned y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o && inc == kNoStart) inc = y;
    }
    unsigned cur = __shfl_up_sync(0xffffffffu, inc, 1);   // the last start strictly before my pixels
    if (lane == 0 || cur == kNoStart) cur = carry;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (px[j] != kNoStart) cur = px[j];
      px[j] = cur;
    }
    if (act)
      *reinterpret_cast<uint4*>(lr + x) = make_uint4(px[0] | (px[1] << 16), px[2] | (px[3] << 16), px[4] | (px[5] << 16), px[6] | (px[7] << 16));
    const unsigned tail = __shfl_sync(0xffffffffu, inc, 31);
    if (tail != kNoStart) carry = tail;
  }
}

// u = int(((gx+1)/2)*(W-1)), v = int(((gy+1)/2)*(H-1))  -- models/models.py:644-645, fp32 op for op (as inverse.cu)
__device__ __forceinline__ int raster_target(float g, int size) {
  const float f = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), static_cast<float>(size - 1));
  return f == f ? __float2int_rz(f) : -1;
}

// loc[v,u] = 0x8000 | n for every node n that won its target pixel (models/models.py:650-651), and the "no value" code
// at image corners no node landed on (corner sites carry the NaN row; several triangles meet there).
__global__ void stamp_nodes_kernel(const float2* __restrict__ grid, const int32_t* __restrict__ winner,
                                   uint16_t* __restrict__ loc, int B, int hw, int H, int W) {
  const int total = B * (hw + 4);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int b = idx / (hw + 4), node = idx - b * (hw + 4);
    const size_t img = static_cast<size_t>(b) * H * W;
    if (node >= hw) {
      const int c = node - hw, v = (c & 2) ? H - 1 : 0, u = (c & 1) ? W - 1 : 0;
      if (winner[img + static_cast<size_t>(v) * W + u] < 0) loc[img + static_cast<size_t>(v) * W + u] = static_cast<uint16_t>(0x8000u | hw);
      continue;
    }
    const float2 g = grid[static_cast<size_t>(b) * hw + node];
    const int u = raster_target(g.x, W), v = raster_target(g.y, H);
    if (u < 0 || u >= W || v < 0 || v >= H) continue;
    const size_t p = img + static_cast<size_t>(v) * W + u;
    if (winner[p] == node) loc[p] = static_cast<uint16_t>(0x8000u | node);
  }
}

// The same stamps from fovea_select_points_sparse's per-node targets (no winner map): targets[b][n] = (row << 16 | column)
// of node n if it won its pixel, else -1; entries hw .. hw+3 name image corners no node landed on.
__global__ void stamp_targets_kernel(const int32_t* __restrict__ targets, uint16_t* __restrict__ loc, int B, int hw, int H, int W) {
  const int total = B * (hw + 4);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int t = targets[idx];
    if (t < 0) continue;
    const int b = idx / (hw + 4), node = idx - b * (hw + 4);
    loc[(static_cast<size_t>(b) * H + (t >> 16)) * W + (t & 0xFFFF)] = static_cast<uint16_t>(0x8000u | min(node, hw));
  }
}

__global__ void fill_none_kernel(uint4* __restrict__ loc, size_t n16, unsigned none2) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    loc[i] = make_uint4(none2, none2, none2, none2);
}

}  // namespace fovea

using namespace fovea;

extern "C" int64_t fovea_locate_raster_workspace_bytes(int B, int H, int W, int tcap) {
  (void)H; (void)W;
  return 16 + static_cast<int64_t>(B) * tcap * 4;   // the queue of tall triangles: length, next, then one word per triangle
}

static int locate_raster(const int32_t* pts, const uint16_t* mesh, const void* trirec, const int32_t* ntri,
                         const float* grid, const int32_t* winner, const int32_t* targets, int B, int h, int w, int H,
                         int W, int cap, int tcap, int prefill, uint16_t* loc, void* workspace, fovea_stream_t stream) {
  FOVEA_REQUIRE(pts && mesh && trirec && ntri && loc, "fovea_locate_raster: null pointer");
  FOVEA_REQUIRE((grid == nullptr) == (winner == nullptr), "fovea_locate_raster: grid and winner go together");
  FOVEA_REQUIRE(B > 0 && h > 0 && w > 0 && H > 1 && W > 1 && cap > 0 && tcap > 0, "fovea_locate_raster: bad sizes");
  FOVEA_REQUIRE(H <= 16384 && W <= 16384 && W % 8 == 0, "fovea_locate_raster: canvas side must be <= 16384 and the width "
                "a multiple of 8 (16-byte span stores)");
  FOVEA_REQUIRE(tcap <= 32768 && static_cast<long long>(h) * w < 32767, "fovea_locate_raster: triangle ids / table rows must fit 15 bits");
  FOVEA_REQUIRE(B <= 65535, "fovea_locate_raster: B too large for the grid");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int hw = h * w;
  static const int tile_max = [] { const char* e = getenv("FOVEA_RAS_TILE_MAX"); return e ? atoi(e) : kRasTileMax; }();
  if (prefill) {  // canvases the triangulation does not cover (no forced corners): everything starts as "no value"
    const size_t n16 = static_cast<size_t>(B) * H * W / 8;
    const unsigned none = 0x8000u | static_cast<unsigned>(hw);
    fill_none_kernel<<<kNumSMs * 8, 256, 0, s>>>(reinterpret_cast<uint4*>(loc), n16, none | (none << 16));
  }
  // FOVEA_RAS_MODE (read per call, so one process can compare): 0 = the pixel sweep, 64 = span-start markers + row sweep, 1 / 2 / 4 / 8 / 16 / 32 = row spans with that many
  // lanes per triangle
  int mode = kRasMode;
  if (const char* e = getenv("FOVEA_RAS_MODE")) mode = atoi(e);
  const uint4* mesh4 = reinterpret_cast<const uint4*>(mesh);
  const TriRec* recs = static_cast<const TriRec*>(trirec);
#define FOVEA_RAS_SPAN(LPT)                                                                                      \
  raster_span_kernel<LPT><<<dim3(ceil_div(tcap, (32 / LPT) * (kRasThreads / 32)), B), kRasThreads, 0, s>>>(       \
      pts, mesh4, recs, ntri, loc, hw, H, W, cap, tcap, tile_max)
  if (mode == 64 && workspace && !prefill && (grid || targets)) {   // markers + row sweep (needs a canvas the mesh covers: every row starts a span)
    // Frame chunks small enough to stay in L2 from the clear to the sweep (the starts are scattered 2-byte stores: into
    // lines that left L2 each one costs a 32-byte read-modify-write in DRAM -- measured, that was 40 % of the mark kernel)
    unsigned* queue = static_cast<unsigned*>(workspace);   // [0] = length, [1] = next to take, [4 ...] = (frame << 16 | triangle)
    const size_t frame_px = static_cast<size_t>(H) * W;
    static const int tall_cost = [] { const char* e = getenv("FOVEA_RAS_TALL_COST"); return e ? atoi(e) : 300; }();
    size_t chunk_mb = 48;
    if (const char* e = getenv("FOVEA_RAS_CHUNK_MB")) chunk_mb = std::min<size_t>(std::max(atoi(e), 1), 4096);
    const int per = static_cast<int>(std::max<size_t>(1, (chunk_mb << 20) / (frame_px * 2)));   // (per * frame_px < 2^32)
    for (int b0 = 0; b0 < B; b0 += per) {
      const int nb = std::min(per, B - b0);
      uint16_t* lc = loc + b0 * frame_px;
      FOVEA_CUDA(cudaMemsetAsync(lc, 0xFF, static_cast<size_t>(nb) * frame_px * 2, s));
      FOVEA_CUDA(cudaMemsetAsync(queue, 0, 16, s));
      raster_mark_kernel<<<dim3(ceil_div(tcap, kMarkThreads), nb), kMarkThreads, 0, s>>>(
          pts + static_cast<size_t>(b0) * cap, mesh4 + static_cast<size_t>(b0) * tcap, recs + static_cast<size_t>(b0) * tcap,
          ntri + b0, lc, queue, H, W, cap, tcap, tall_cost);
      raster_mark_tall_kernel<<<kNumSMs * FOVEA_TALL_CTAS, 256, 0, s>>>(
          pts + static_cast<size_t>(b0) * cap, mesh4 + static_cast<size_t>(b0) * tcap, recs + static_cast<size_t>(b0) * tcap,
          ntri + b0, lc, queue, H, W, cap, tcap);
      const long long rows = static_cast<long long>(nb) * H;
      raster_fill_rows_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, s>>>(lc, rows, W, 0x8000u | static_cast<unsigned>(hw));
    }
  } else if (mode == 1) FOVEA_RAS_SPAN(1);
  else if (mode == 2) FOVEA_RAS_SPAN(2);
  else if (mode == 4) FOVEA_RAS_SPAN(4);
  else if (mode == 8) FOVEA_RAS_SPAN(8);
  else if (mode == 16) FOVEA_RAS_SPAN(16);
  else if (mode == 32) FOVEA_RAS_SPAN(32);
  else
    raster_locate_kernel<<<dim3(ceil_div(tcap, 4 * (kRasThreads / 32)), B), kRasThreads, 0, s>>>(
        pts, mesh4, recs, ntri, loc, hw, H, W, cap, tcap, tile_max);
#undef FOVEA_RAS_SPAN
  const int total = B * (hw + 4);
  if (targets)
    stamp_targets_kernel<<<min(ceil_div(total, 256), kNumSMs * 8), 256, 0, s>>>(targets, loc, B, hw, H, W);
  else if (grid)
    stamp_nodes_kernel<<<min(ceil_div(total, 256), kNumSMs * 8), 256, 0, s>>>(reinterpret_cast<const float2*>(grid), winner,
                                                                             loc, B, hw, H, W);
  return check_launch("fovea_locate_raster");
}

extern "C" int fovea_locate_raster(const int32_t* pts, const uint16_t* mesh, const void* trirec, const int32_t* ntri,
                                   const float* grid, const int32_t* winner, int B, int h, int w, int H, int W, int cap,
                                   int tcap, int prefill, uint16_t* loc, void* workspace, fovea_stream_t stream) {
  return locate_raster(pts, mesh, trirec, ntri, grid, winner, nullptr, B, h, w, H, W, cap, tcap, prefill, loc, workspace, stream);
}

extern "C" int fovea_locate_raster_targets(const int32_t* pts, const uint16_t* mesh, const void* trirec, const int32_t* ntri,
                                           const int32_t* targets, int B, int h, int w, int H, int W, int cap, int tcap,
                                           uint16_t* loc, void* workspace, fovea_stream_t stream) {
  FOVEA_REQUIRE(targets, "fovea_locate_raster_targets: null pointer");
  return locate_raster(pts, mesh, trirec, ntri, nullptr, nullptr, targets, B, h, w, H, W, cap, tcap, 0, loc, workspace, stream);
}

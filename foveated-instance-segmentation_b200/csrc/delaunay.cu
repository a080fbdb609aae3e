// Device Delaunay triangulation: one CTA (1024 threads) per image, the whole mesh in shared memory.
//
// Replaces the host Qhull call of the reference (interp2d.py:55, spatial/qhull.pyx:1679) for the <= ~6.4k integer
// pixel sites of one image.  Input points arrive sorted row-major and unique (fovea_select_points).
//
//   1. rows      : group points by pixel row (block scan).
//   2. strips    : the band between two consecutive occupied rows is triangulated by the merge ("zipper") of the
//                  two sorted rows -- closed form per row edge (two binary searches), so all strip triangles and
//                  their adjacency are produced in parallel with no hashing.
//   3. pockets   : the two regions between the left/right chain of row end points and the convex hull are
//                  monotone mountains; they are closed by PARALLEL ear clipping (independent sets of convex chain
//                  vertices per round), which also yields the hull edges.
//   4. Lawson    : parallel edge flips with exact int64 in-circle tests until every interior edge is locally
//                  Delaunay.  Each round every dirty triangle proposes one illegal edge; proposals claim the two
//                  triangles the flip rewrites with a random-priority atomicMin, winners flip.  The back links of the
//                  four outer neighbours -- which a flip re-points and which may be flipping themselves in the same
//                  round -- go through postings the winners leave in their own lock words (DT_CLAIM2 below; the first
//                  version claimed all six triangles instead: twice the rounds).
//                  Random priorities matter: with index priorities the skinny strips serialise (~10^4 rounds on a
//                  1024^2 frame, ~150 with random ones and six-triangle claims, ~75 with two; measured in the NumPy
//                  prototypes of this algorithm, tools/prototypes/dt_claims.py).
//
// Exactness: coordinates < 8192 keep the 4th-order in-circle determinant inside int64 (checked on the host).
// Co-circular point sets (ubiquitous on a pixel lattice) have no unique Delaunay triangulation; any locally
// Delaunay result is accepted (in-circle == 0 is legal), exactly as Qhull's 'Qt' picks an arbitrary one.
#include <stdlib.h>

#include "common.cuh"

namespace fovea {

constexpr int kDtThreads = 1024;
constexpr unsigned kNoneCode = 0xFFFFu;     // hull edge
constexpr unsigned kPendingCode = 0xFFFEu;  // chain edge of an empty strip (both pockets touch it)
constexpr unsigned kFixLeft = 0xFFFDu;      // strip triangle whose left neighbour slot is filled in pass 2

struct DtArrays {
  unsigned short *v0, *v1, *v2;  // vertex ids, counter-clockwise in (x=col, y=row)
  unsigned short *n0, *n1, *n2;  // neighbour codes: (triangle << 2) | slot-in-neighbour, opposite v0/v1/v2
  unsigned* lock;                // [tcap] flip claims (aliased by the construction scratch below)
  unsigned* dirty;               // [ceil(tcap/32)] bitmask
  int* spts;                     // [cap] packed points (every predicate of the flip phase reads them)
  unsigned short* rowStart;      // [n+1]  (global workspace: only the construction phases read it)
  // construction scratch, aliased onto `lock`
  unsigned short *stripBase, *cprev, *cnext, *cown;
};

__device__ __forceinline__ int pt_row(int p) { return p >> 16; }
__device__ __forceinline__ int pt_col(int p) { return p & 0xFFFF; }

__device__ __forceinline__ long long orient_pts(int a, int b, int c) {
  return static_cast<long long>(pt_col(b) - pt_col(a)) * (pt_row(c) - pt_row(a)) -
         static_cast<long long>(pt_row(b) - pt_row(a)) * (pt_col(c) - pt_col(a));
}

// > 0  <=>  d strictly inside the circumcircle of the counter-clockwise triangle (a,b,c).
// Coordinates < 8192 (checked on the host): differences < 2^13, so squares/cross products fit int32 (< 2^27) and only
// the three final products need 64 bits (< 2^55).
__device__ __forceinline__ long long incircle_pts(int a, int b, int c, int d) {
  const int ax = pt_col(a) - pt_col(d), ay = pt_row(a) - pt_row(d);
  const int bx = pt_col(b) - pt_col(d), by = pt_row(b) - pt_row(d);
  const int cx = pt_col(c) - pt_col(d), cy = pt_row(c) - pt_row(d);
  return static_cast<long long>(ax * ax + ay * ay) * (bx * cy - by * cx) -
         static_cast<long long>(bx * bx + by * by) * (ax * cy - ay * cx) +
         static_cast<long long>(cx * cx + cy * cy) * (ax * by - ay * bx);
}

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// position of the n-th (0-based) set bit of w; branch-free binary search on popcounts (w must have > n set bits)
__device__ __forceinline__ int nth_set_bit(unsigned w, int n) {
  int pos = 0;
  int c = __popc(w & 0xFFFFu);
  if (n >= c) { n -= c; w >>= 16; pos += 16; }
  c = __popc(w & 0xFFu);
  if (n >= c) { n -= c; w >>= 8; pos += 8; }
  c = __popc(w & 0xFu);
  if (n >= c) { n -= c; w >>= 4; pos += 4; }
  c = __popc(w & 0x3u);
  if (n >= c) { n -= c; w >>= 2; pos += 2; }
  if (n >= static_cast<int>(w & 1u)) pos += 1;
  return pos;
}

// The same scan with ONE barrier: every warp sums the totals of the warps before it itself (and all 32 for `total`).
// The caller must put a barrier between two calls (warp_sums is rewritten by the next one).
__device__ __forceinline__ int block_scan_excl_1bar(int v, int* warp_sums, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  const int w = warp_sums[lane];                       // (kDtThreads = 1024: exactly 32 warps)
  total = __reduce_add_sync(0xffffffffu, w);
  const int before = __reduce_add_sync(0xffffffffu, lane < warp ? w : 0);
  return before + incl - v;
}

// exclusive block scan of one int per thread; returns the exclusive prefix, `total` = block sum
__device__ int block_scan_excl(int v, int* warp_sums, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += y;
    }
    warp_sums[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) warp_sums[32] = wi;
  }
  __syncthreads();
  const int res = warp_sums[warp] + incl - v;
  total = warp_sums[32];
  __syncthreads();
  return res;
}

// number of x in [0,cnt) with col(pts[first+x]) < key  (strict)  /  <= key (non-strict)
__device__ __forceinline__ int count_less(const int* pts, int first, int cnt, int key, bool or_equal) {
  int lo = 0, hi = cnt;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int c = pt_col(pts[first + mid]);
    if (or_equal ? (c <= key) : (c < key)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

#ifdef DT_DEBUG
__device__ int* g_dt_dbg = nullptr;
__device__ __forceinline__ int ck(int idx, int lim, int line) {
  if (idx < 0 || idx >= lim) { if (g_dt_dbg) { g_dt_dbg[0] = -line; g_dt_dbg[1] = idx; g_dt_dbg[2] = lim; } return 0; }
  return idx;
}
#define CK(i, lim) ck((i), (lim), __LINE__)
#else
#define CK(i, lim) (i)
#endif

__device__ __forceinline__ int nx3(int k) { return (0x09 >> (2 * k)) & 3; }   // (k + 1) % 3 for k in 0..2
__device__ __forceinline__ int pv3(int k) { return (0x12 >> (2 * k)) & 3; }   // (k + 2) % 3
__device__ __forceinline__ unsigned short& nb_slot(const DtArrays& A, int t, int k) {
  return k == 0 ? A.n0[t] : (k == 1 ? A.n1[t] : A.n2[t]);
}
__device__ __forceinline__ unsigned short vert(const DtArrays& A, int t, int k) {
  return k == 0 ? A.v0[t] : (k == 1 ? A.v1[t] : A.v2[t]);
}
__device__ __forceinline__ void link_back(const DtArrays& A, unsigned code, unsigned me) {
  if (code < kPendingCode) nb_slot(A, code >> 2, code & 3) = static_cast<unsigned short>(me);
}

#ifndef DT_CLAIM2
#define DT_CLAIM2 1   // 1: a flip claims its two triangles (back links through postings); 0: the six-triangle claim
#endif
// Lock words of the block-wide flip rounds: [31:26] round tag (decreasing), then either a CLAIM -- bit 25 set, 11 random
// priority bits, the proposing triangle -- or, once a flip has won, a POSTING -- bit 25 clear, bit 16 = second triangle
// of the pair, [15:14] this triangle's slot of the old diagonal, [13:0] its partner.
constexpr unsigned kClaimBit = 1u << 25, kPostIsU = 1u << 16;
__device__ __forceinline__ unsigned claim_word(unsigned tag, int t, int round) {
  // (a cheaper two-multiplication hash was tried: same time, more rounds on the slowest frames)
  return tag | kClaimBit | ((hash32(t * 2654435761u + round * 0x9E3779B9u) & 0x7FFu) << 14) | static_cast<unsigned>(t);
}
#ifndef DT_MAXNREG
#define DT_MAXNREG 64   // (48 would let one inverse_fill CTA co-reside per SM under the pipelined schedule: measured, no gain)
#endif
// kBig: claims and dirty bits in global memory (lock_g); otherwise everything is shared memory and the template keeps
// the compiler's address-space inference intact (LDS / STS / ATOMS instead of generic LD / ST / ATOM).
template <bool kBig>
__global__ void __maxnreg__(DT_MAXNREG)
delaunay_kernel(const int32_t* __restrict__ pts_g, const int32_t* __restrict__ npts, int cap, int tcap,
                uint16_t* __restrict__ mesh_out, int32_t* __restrict__ ntri_out, int32_t* __restrict__ rounds_out,
                int max_rounds, int32_t* __restrict__ dbg, unsigned short* __restrict__ row_ws,
                int32_t* __restrict__ hints_out, int H, int W, int mesh_stride, unsigned* __restrict__ lock_g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int warp_sums[33];
  __shared__ int s_ntri;
  __shared__ int s_tail[2];
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const long long clk0 = clock64();
  const int32_t* pts_in = pts_g + static_cast<size_t>(b) * cap;
  const int n = npts[b];

  DtArrays A;
  {
    unsigned char* p = smem_raw;
    A.v0 = reinterpret_cast<unsigned short*>(p); p += 2 * tcap;
    A.v1 = reinterpret_cast<unsigned short*>(p); p += 2 * tcap;
    A.v2 = reinterpret_cast<unsigned short*>(p); p += 2 * tcap;
    A.n0 = reinterpret_cast<unsigned short*>(p); p += 2 * tcap;
    A.n1 = reinterpret_cast<unsigned short*>(p); p += 2 * tcap;
    A.n2 = reinterpret_cast<unsigned short*>(p); p += 2 * tcap;
    if (kBig) {  // large site sets (e.g. the 64 x 128 lattice): claims and dirty bits live in global memory (L2), the
                   // mesh itself and the points still fit one CTA's shared memory
      A.lock = lock_g + static_cast<size_t>(b) * (tcap + (tcap + 31) / 32);
      A.dirty = A.lock + tcap;
    } else {
      A.lock = reinterpret_cast<unsigned*>(p); p += 4 * tcap;
      A.dirty = reinterpret_cast<unsigned*>(p); p += 4 * ((tcap + 31) / 32);
    }
    A.spts = reinterpret_cast<int*>(p);
    A.rowStart = row_ws + static_cast<size_t>(b) * (cap + 2);
    unsigned short* s = reinterpret_cast<unsigned short*>(A.lock);  // 2*tcap shorts of scratch
    A.stripBase = s;
    A.cprev = s + cap;
    A.cnext = s + 2 * cap;
    A.cown = s + 3 * cap;
  }
  uint16_t* mesh = mesh_out + static_cast<size_t>(b) * mesh_stride * 8;
  for (int i = tid; i < n; i += kDtThreads) A.spts[i] = pts_in[i];
  const int* pts = A.spts;
  __syncthreads();
  if (dbg && tid == 0) { dbg[b * 8 + 0] = 1; dbg[b * 8 + 1] = n; }

  // ------------------------------------------------------------------ 1. rows
  const int items = (n + kDtThreads - 1) / kDtThreads;
  const int base = tid * items;
  int cnt = 0;
  for (int i = 0; i < items; ++i) {
    const int p = base + i;
    if (p < n && (p == 0 || pt_row(pts[p]) != pt_row(pts[p - 1]))) ++cnt;
  }
  int R;
  int off = block_scan_excl(cnt, warp_sums, R);
  for (int i = 0; i < items; ++i) {
    const int p = base + i;
    if (p < n && (p == 0 || pt_row(pts[p]) != pt_row(pts[p - 1]))) A.rowStart[off++] = static_cast<unsigned short>(p);
  }
  if (tid == 0) A.rowStart[R] = static_cast<unsigned short>(n);
  __syncthreads();
  if (n > cap || 2 * n - 5 > tcap) {  // more triangles than the 16-bit mesh encoding holds: report, never overrun
    if (tid == 0) { ntri_out[b] = 0; if (rounds_out) rounds_out[b] = -1; }
    if (hints_out) {
      const int nh = ceil_div(H, FOVEA_HINT_CELL_H) * ceil_div(W, FOVEA_HINT_CELL_W);
      for (int i = tid; i < nh; i += kDtThreads) hints_out[static_cast<size_t>(b) * nh + i] = 0;
    }
    return;
  }
  if (n < 3 || R < 2) {  // nothing to triangulate (all points collinear in one row)
    if (tid == 0) { ntri_out[b] = 0; if (rounds_out) rounds_out[b] = 0; }
    if (hints_out) {
      const int nh = ceil_div(H, FOVEA_HINT_CELL_H) * ceil_div(W, FOVEA_HINT_CELL_W);
      for (int i = tid; i < nh; i += kDtThreads) hints_out[static_cast<size_t>(b) * nh + i] = 0;
    }
    return;
  }

  // The row starts are read by every binary search of the strips and every ear test of the pockets: when they fit, they
  // move from the global workspace into the unused upper half of a scratch array (each array holds `cap` entries, the
  // construction uses the first R of them).
  if (!kBig && R + 1 <= cap - cap / 2) {
    unsigned short* rs = A.stripBase + cap / 2;
    for (int i = tid; i <= R; i += kDtThreads) rs[i] = A.rowStart[i];
    __syncthreads();
    A.rowStart = rs;
  }
  if (dbg && tid == 0) dbg[b * 8 + 1] = (int)((clock64() - clk0) >> 4);
  // ------------------------------------------------------------------ 2. strips
  const int nstrips = R - 1;
  const int sitems = (nstrips + kDtThreads - 1) / kDtThreads;
  const int sbase = tid * sitems;
  cnt = 0;
  for (int i = 0; i < sitems; ++i) {
    const int r = sbase + i;
    if (r < nstrips) cnt += static_cast<int>(A.rowStart[r + 2]) - static_cast<int>(A.rowStart[r]) - 2;
  }
  int nstrip_tris;
  off = block_scan_excl(cnt, warp_sums, nstrip_tris);
  for (int i = 0; i < sitems; ++i) {
    const int r = sbase + i;
    if (r < nstrips) {
      A.stripBase[r] = static_cast<unsigned short>(off);
      const int c = static_cast<int>(A.rowStart[r + 2]) - static_cast<int>(A.rowStart[r]) - 2;
      off += c;
      // chain-edge owners default to "empty strip"; overwritten below when the strip has triangles
      A.cown[r] = static_cast<unsigned short>(c == 0 ? kPendingCode : kNoneCode);   // left chain owners
      A.cnext[r] = static_cast<unsigned short>(c == 0 ? kPendingCode : kNoneCode);  // right chain owners (temp)
    }
  }
  __syncthreads();

  for (int p = tid; p + 1 < n; p += kDtThreads) {
    // row of p: largest r with rowStart[r] <= p
    int lo = 0, hi = R - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (A.rowStart[mid] <= p) lo = mid; else hi = mid - 1;
    }
    const int r = lo;
    const int rs = A.rowStart[r], re = A.rowStart[r + 1];
    if (p + 1 >= re) continue;  // p is the last point of its row: no row edge (p, p+1)
    const int a = p - rs;
    const int key = pt_col(pts[p + 1]);
    int idD = -1, idU = -1, j = 0, iu = 0;
    if (r + 1 < R) {  // (p,p+1) is a top edge of strip r: "down" triangle (p, p+1, bottom_j)
      const int bs = A.rowStart[r + 1], nb = A.rowStart[r + 2] - bs;
      j = count_less(pts, bs + 1, nb - 1, key, false);
      idD = A.stripBase[r] + a + j;
    }
    if (r > 0) {  // (p,p+1) is a bottom edge of strip r-1: "up" triangle (top_i, p+1, p)
      const int ts = A.rowStart[r - 1], nt = rs - ts;
      iu = count_less(pts, ts + 1, nt - 1, key, true);
      idU = A.stripBase[r - 1] + a + iu;
    }
    if (idD >= 0) {
      const int bs = A.rowStart[r + 1];
      const int scount = static_cast<int>(A.rowStart[r + 2]) - rs - 2;
      const int pos = idD - A.stripBase[r];
      A.v0[idD] = static_cast<unsigned short>(p);
      A.v1[idD] = static_cast<unsigned short>(p + 1);
      A.v2[idD] = static_cast<unsigned short>(bs + j);
      A.n2[idD] = static_cast<unsigned short>(idU >= 0 ? ((idU << 2) | 0) : kNoneCode);
      if (pos == scount - 1) { A.n0[idD] = kNoneCode; A.cnext[r] = static_cast<unsigned short>((idD << 2) | 0); }
      else A.n0[idD] = static_cast<unsigned short>(((idD + 1) << 2) | 1);
      if (pos == 0) { A.n1[idD] = kNoneCode; A.cown[r] = static_cast<unsigned short>((idD << 2) | 1); }
      else A.n1[idD] = kFixLeft;
    }
    if (idU >= 0) {
      const int ts = A.rowStart[r - 1];
      const int scount = re - ts - 2;
      const int pos = idU - A.stripBase[r - 1];
      A.v0[idU] = static_cast<unsigned short>(ts + iu);
      A.v1[idU] = static_cast<unsigned short>(p + 1);
      A.v2[idU] = static_cast<unsigned short>(p);
      A.n0[idU] = static_cast<unsigned short>(idD >= 0 ? ((idD << 2) | 2) : kNoneCode);
      if (pos == scount - 1) { A.n2[idU] = kNoneCode; A.cnext[r - 1] = static_cast<unsigned short>((idU << 2) | 2); }
      else A.n2[idU] = static_cast<unsigned short>(((idU + 1) << 2) | 1);
      if (pos == 0) { A.n1[idU] = kNoneCode; A.cown[r - 1] = static_cast<unsigned short>((idU << 2) | 1); }
      else A.n1[idU] = kFixLeft;
    }
  }
  if (tid == 0) { s_ntri = nstrip_tris; if (dbg) { dbg[b * 8 + 0] = 2; dbg[b * 8 + 2] = R; dbg[b * 8 + 3] = nstrip_tris; } }
  __syncthreads();
  // pass 2: left neighbour's slot depends on that neighbour's kind (down: right diagonal opposite v0, up: opposite v2)
  for (int t = tid; t < nstrip_tris; t += kDtThreads) {
    if (A.n1[t] == kFixLeft) {
      const bool prev_down = A.v1[t - 1] == A.v0[t - 1] + 1;
      A.n1[t] = static_cast<unsigned short>(((t - 1) << 2) | (prev_down ? 0 : 2));
    }
  }
  __syncthreads();
  // right-chain owners were parked in cnext; move them into the (now dead) stripBase array
  unsigned short* rown = A.stripBase;
  for (int r = tid; r < nstrips; r += kDtThreads) {
    rown[r] = A.cnext[r];
  }
  __syncthreads();

  if (dbg && tid == 0) dbg[b * 8 + 2] = (int)((clock64() - clk0) >> 4);
  // ------------------------------------------------------------------ 3. pockets (left, then right)
  // cprev[m]: bits 0..12 = previous alive chain vertex (kCNone = chain start), bit 13 DEAD
  constexpr unsigned kCIdx = 0x1FFFu, kCNone = 0x1FFFu, kCDead = 0x2000u;
  int ntri_run = nstrip_tris;  // triangles so far (identical in every thread)
  bool converged = true;       // false: a safety bound ended a loop early (block-uniform) -> the frame reports no mesh
  for (int side = 0; side < 2; ++side) {
    auto cpt = [&](int m) { return side == 0 ? static_cast<int>(A.rowStart[m]) : static_cast<int>(A.rowStart[m + 1]) - 1; };
    auto cpri = [&](int m, int round) {
      return (hash32(static_cast<unsigned>(m) * 2654435761u + static_cast<unsigned>(round) * 40503u + side) << 13) |
             static_cast<unsigned>(m);
    };
    for (int m = tid; m < R; m += kDtThreads) {
      A.cprev[m] = static_cast<unsigned short>(m == 0 ? kCNone : m - 1);
      A.cnext[m] = static_cast<unsigned short>(m == R - 1 ? kCNone : m + 1);
      if (side == 1 && m < nstrips) A.cown[m] = rown[m];
    }
    __syncthreads();
    bool closed = false;
    for (int round = 0; round < 4 * R + 64; ++round) {
      if (dbg && tid == 0) dbg[b * 8 + 4 + side] = round;
      // phase 1: an alive interior chain vertex is an EAR if it is strictly convex towards the pocket; an ear is selected
      // iff it beats both neighbouring ears (random priority: an independent set).  One pass: every thread also tests
      // its two neighbours (two more orientation tests, but no barrier between "which are ears" and "who wins"), and
      // keeps its selections in registers.  Nothing is written in this phase.
      auto is_ear = [&](unsigned pv, unsigned m, unsigned nx) {
        if (pv == kCNone || nx == kCNone) return false;
        const long long o = orient_pts(pts[cpt(pv)], pts[cpt(m)], pts[cpt(nx)]);
        return side == 0 ? (o > 0) : (o < 0);
      };
      unsigned selmask = 0;   // bit i: my i-th chain vertex (m = tid + i * kDtThreads) is clipped this round
      int mysel = 0, si = 0;
      for (int m = tid; m < R; m += kDtThreads, ++si) {
        const unsigned cp = A.cprev[m];
        if (cp & kCDead) continue;
        const unsigned pv = cp & kCIdx, nx = A.cnext[m];
        if (!is_ear(pv, m, nx)) continue;
        const unsigned pri = cpri(m, round);
        if (cpri(pv, round) < pri && is_ear(A.cprev[pv] & kCIdx, pv, m)) continue;
        if (cpri(nx, round) < pri && is_ear(m, nx, A.cnext[nx])) continue;
        selmask |= 1u << si;
        ++mysel;
      }
      // phase 2: clip the selected ears (their neighbours are not selected, so the list surgery is race-free).
      // Triangle ids come from a block scan over the selected ears, not from an atomic counter: the numbering -- and
      // with it the flip priorities and the final choice among co-circular alternatives -- is the same on every run.
      // (The scan's barriers also separate the reads above from the writes below.)
      int nsel;
      int E_next = ntri_run + block_scan_excl_1bar(mysel, warp_sums, nsel);   // (the round ends with a barrier)
      if (nsel == 0) { closed = true; break; }
      ntri_run += nsel;
      si = 0;
      for (int m = tid; m < R; m += kDtThreads, ++si) {
        if (!((selmask >> si) & 1u)) continue;
        const unsigned cp = A.cprev[m];
        const int pv = cp & kCIdx, nx = A.cnext[m];
        const int E = E_next++;
        const int P = cpt(pv), M = cpt(m), N = cpt(nx);
        // left : (P,M,N) is ccw: edge(P,M) opposite v2, edge(M,N) opposite v0, edge(N,P) opposite v1
        // right: (P,N,M) is ccw: edge(P,M) opposite v1, edge(M,N) opposite v0, edge(P,N) opposite v2
        const int s_pm = side == 0 ? 2 : 1, s_mn = 0, s_pn = side == 0 ? 1 : 2;
        A.v0[E] = static_cast<unsigned short>(P);
        A.v1[E] = static_cast<unsigned short>(side == 0 ? M : N);
        A.v2[E] = static_cast<unsigned short>(side == 0 ? N : M);
        unsigned own_pm = A.cown[pv], own_mn = A.cown[m];
        if (own_pm == kPendingCode) {  // edge of an empty strip: the right pocket owns its far side (if anyone)
          if (side == 0) rown[pv] = static_cast<unsigned short>((E << 2) | s_pm);
          own_pm = kNoneCode;
        }
        if (own_mn == kPendingCode) {
          if (side == 0) rown[m] = static_cast<unsigned short>((E << 2) | s_mn);
          own_mn = kNoneCode;
        }
        nb_slot(A, E, s_pm) = static_cast<unsigned short>(own_pm);
        nb_slot(A, E, s_mn) = static_cast<unsigned short>(own_mn);
        nb_slot(A, E, s_pn) = static_cast<unsigned short>(kNoneCode);
        link_back(A, own_pm, (E << 2) | s_pm);
        link_back(A, own_mn, (E << 2) | s_mn);
        A.cown[pv] = static_cast<unsigned short>((E << 2) | s_pn);
        A.cnext[pv] = static_cast<unsigned short>(nx);
        A.cprev[nx] = static_cast<unsigned short>((A.cprev[nx] & ~kCIdx) | static_cast<unsigned>(pv));
        A.cprev[m] = static_cast<unsigned short>(kCDead);
      }
      __syncthreads();
    }
    converged = converged && closed;
    __syncthreads();
  }
  const int T = ntri_run;
  if (dbg && tid == 0) { dbg[b * 8 + 0] = 3; dbg[b * 8 + 6] = T; }
  __syncthreads();

  if (dbg && tid == 0) dbg[b * 8 + 3] = (int)((clock64() - clk0) >> 4);
  // ------------------------------------------------------------------ 4. Lawson flips
  // Flat accessors: the three vertex (neighbour) arrays are consecutive, so slot k of triangle t is one load.
  unsigned short* const VV = A.v0;
  unsigned short* const NN = A.n0;
#define DT_V(t, k) VV[(k) * tcap + (t)]
#define DT_N(t, k) NN[(k) * tcap + (t)]
  const int nwords = (tcap + 31) / 32;
  for (int t = tid; t < tcap; t += kDtThreads) A.lock[t] = 0xFFFFFFFFu;
  for (int w = tid; w < nwords; w += kDtThreads) {
    const int lo = w * 32;
    A.dirty[w] = lo + 32 <= T ? 0xFFFFFFFFu : (lo >= T ? 0u : ((1u << (T - lo)) - 1u));
  }
  __syncthreads();
  // Every warp owns a contiguous range of dirty words and processes ITS set bits densely (rank -> lane), so a round
  // costs ceil(dirty/32) full-warp test iterations instead of one mostly-idle iteration per 32 triangles.
  const int lane = tid & 31, warp = tid >> 5;
  const int wpw = (nwords + 31) / 32;  // dirty words per warp (<= 16 for tcap <= 16383)
#ifndef DT_STRIDED
#define DT_STRIDED 1
#endif
#if DT_STRIDED   // word j of a warp: interleaved over the mesh (a hot region's words spread over all warps) ...
#define DT_WORD(j) (warp + 32 * (j))
#else            // ... or one contiguous range per warp
#define DT_WORD(j) (warp * wpw + (j))
#endif
  int round = 0, tail_hold = 0;
  bool flips_done = false;
  for (; round < max_rounds; ++round) {
    // Claims carry a 6-bit round tag that DEcreases every round, so this round's atomicMin always beats the stale
    // claims of earlier rounds and the lock array only needs a reset when the tag wraps (every 64 rounds).
    if ((round & 63) == 0 && round > 0) {
      for (int t = tid; t < T; t += kDtThreads) A.lock[t] = 0xFFFFFFFFu;
      __syncthreads();
    }
#ifdef DT_PROFILE
    if (tid == 0 && round < 512) {  // per-round clock + dirty count, appended after the regular workspace
      int nd = 0;
      for (int w = 0; w < nwords; ++w) nd += __popc(A.dirty[w]);
      int* prof = reinterpret_cast<int*>(row_ws + static_cast<size_t>(gridDim.x) * (cap + 2)) + (blockIdx.x * 512 + round) * 2;
      prof[0] = (int)((clock64() - clk0) >> 4);
      prof[1] = nd;
    }
#endif
    const unsigned tag = static_cast<unsigned>(63 - (round & 63)) << 26;
    // P0: snapshot my warp's dirty words and clear them; triangles that stay illegal re-set their bit below
    // While a warp's range is densely dirty (the first rounds after the strip construction) only a random quarter
    // of its dirty triangles competes per round: six-triangle claims make ~12 candidates contend for every win, so
    // thinning costs few wins per round but removes 3/4 of the tests and claims.  Unselected bits stay dirty.
    unsigned myword = 0;
    bool deferred = false;
    if (lane < wpw && DT_WORD(lane) < nwords) {
      const unsigned w = A.dirty[DT_WORD(lane)];
#ifndef DT_THIN
#define DT_THIN 4   // keep one dirty triangle in DT_THIN of a dense word per round (2, 4 or 8)
#endif
#ifndef DT_DENSE
#define DT_DENSE 32  // a word is dense when more than this many of its 32 triangles are dirty (32: no thinning)
#endif
      myword = w;
      if (DT_DENSE < 32) {   // (compile-time: no thinning with the default)
        const bool dense = __popc(w) > DT_DENSE;
        const unsigned h = hash32(DT_WORD(lane) * 0x9E3779B9u + round * 0x85EBCA6Bu);
        const unsigned keep = DT_THIN == 2 ? h : (DT_THIN == 4 ? (h & hash32(h)) : (h & hash32(h) & hash32(h ^ 0x5bd1e995u)));
        if (dense) myword = w & keep;
        if (dense && myword == 0u) myword = w & (0u - w);  // keep at least one (lowest) bit so progress is guaranteed
      }
      A.dirty[DT_WORD(lane)] = w & ~myword;
      deferred = (w & ~myword) != 0u;  // unselected dirty triangles: the loop must not terminate this round
    }
    // inclusive prefix of the per-word popcounts across the warp (lane j <-> word j)
    const int mycnt = __popc(myword);
    int incl = mycnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    const int excl = incl - mycnt;
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    // triangle of rank r + lane among this warp's dirty bits: the word holding the rank-th bit is the smallest lane j with
    // incl_j > rank (binary search by shuffles -- every lane must take part)
    auto ranked = [&](int r) -> int {
      const int rank = r + lane;
      int jsel = 0;
#pragma unroll
      for (int sstep = 16; sstep > 0; sstep >>= 1) {
        const int v = __shfl_sync(0xffffffffu, incl, jsel + sstep - 1);
        if (v <= rank) jsel += sstep;
      }
      jsel = min(jsel, 31);
      const unsigned wsel = __shfl_sync(0xffffffffu, myword, jsel);
      const int ex = __shfl_sync(0xffffffffu, excl, jsel);
      return rank < total ? DT_WORD(jsel) * 32 + nth_set_bit(wsel, rank - ex) : -1;
    };
    int tc0 = -1, tc1 = -1, tc2 = -1, tc3 = -1;
    unsigned cand = 0;  // 2 bits per iteration: 0 = none, else illegal edge slot + 1
    bool any = deferred;
    int it = 0;
    for (int base = 0; base < total; base += 32, ++it) {
      // triangle of rank base+lane among this warp's dirty bits (kept for the two later passes of this round while it
      // fits the four cache registers: most rounds have at most 128 dirty triangles per warp)
      const int t = ranked(base);
      if (it == 0) tc0 = t; else if (it == 1) tc1 = t; else if (it == 2) tc2 = t; else if (it == 3) tc3 = t;
      if (t < 0) continue;
      const int pa = pts[DT_V(t, 0)], pb = pts[DT_V(t, 1)], pc = pts[DT_V(t, 2)];
      // all three neighbours are fetched and tested together (no early exit): the round's critical path is this
      // dependent chain, and three independent chains cost one latency instead of up to three
      unsigned code3[3];
      int d3[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) code3[k] = DT_N(t, k);
#pragma unroll
      for (int k = 0; k < 3; ++k) d3[k] = code3[k] < kPendingCode ? pts[DT_V(code3[k] >> 2, code3[k] & 3)] : 0;
      int found = -1;
      unsigned ucode = 0;
#pragma unroll
      for (int k = 2; k >= 0; --k)
        if (code3[k] < kPendingCode && incircle_pts(pa, pb, pc, d3[k]) > 0) { found = k; ucode = code3[k]; }
      if (found < 0) continue;
      any = true;
      cand |= static_cast<unsigned>(found + 1) << (2 * it);
      atomicOr(&A.dirty[t >> 5], 1u << (t & 31));  // stays dirty until it wins
      const unsigned pri = claim_word(tag, t, round);
      const int u = ucode >> 2;
#if !DT_CLAIM2
      const int ku = ucode & 3;
#endif
      atomicMin(&A.lock[t], pri);
      atomicMin(&A.lock[u], pri);
#if !DT_CLAIM2
      const unsigned o1 = DT_N(t, (found + 1) % 3), o2 = DT_N(t, (found + 2) % 3);
      const unsigned o3 = DT_N(u, (ku + 1) % 3), o4 = DT_N(u, (ku + 2) % 3);
      if (o1 < kPendingCode) atomicMin(&A.lock[o1 >> 2], pri);
      if (o2 < kPendingCode) atomicMin(&A.lock[o2 >> 2], pri);
      if (o3 < kPendingCode) atomicMin(&A.lock[o3 >> 2], pri);
      if (o4 < kPendingCode) atomicMin(&A.lock[o4 >> 2], pri);
#endif
    }
    if (!__syncthreads_or(any)) { flips_done = true; break; }
#if DT_CLAIM2
    // P2 with TWO-triangle claims.  A flip claims only the two triangles it rewrites, so flips of neighbouring quads win in
    // the same round (tools/prototypes/dt_claims.py: a third of the rounds with > 100 illegal edges).  What the
    // six-triangle claim protected were the back links: a flip re-points the slots of its four outer neighbours, and a
    // neighbour that flips in the same round is being rewritten itself.  Hence two steps:
    //   P2a  every winner replaces the claim words of its two triangles by a POSTING (same round tag, bit 25 clear):
    //        which triangle of the pair this is, its slot of the old diagonal, and its partner.  Other proposers still
    //        comparing claim words are not disturbed: a posting equals nobody's claim (bit 25), and both words were
    //        already lost to them.  The next round's claims (smaller tag) overwrite postings like any stale claim.
    //   P2b  every winner rewrites its two triangles; across each outer edge it reads the neighbour's lock word: a
    //        posting of this round means the neighbour flipped too -- the edge's new owner on that side follows from the
    //        posting, and the neighbour learns mine the same way -- else the unchanged neighbour's slot is re-pointed.
    unsigned wins = 0;
    it = 0;
    for (int base = 0; base < total; base += 32, ++it) {
      const int sel = (cand >> (2 * it)) & 3;
      if (!__any_sync(0xffffffffu, sel != 0)) continue;  // warp-uniform: the shuffles below need every lane
      const int t = it < 4 ? (it == 0 ? tc0 : (it == 1 ? tc1 : (it == 2 ? tc2 : tc3))) : ranked(base);
      if (!sel) continue;
      const int k = sel - 1;
      const unsigned lt = A.lock[t];
      const unsigned ucode = DT_N(t, k);
      const int u = ucode < kPendingCode ? static_cast<int>(ucode >> 2) : t, ku = ucode & 3;  // (a hull marker is no index)
      const unsigned lu = A.lock[u];
      // both words hold MY claim of this round (the priority bits need not be recomputed: a claim names its proposer)
      const bool mine = (lt & 0x3FFFu) == static_cast<unsigned>(t) && (lt & (0xFC000000u | kClaimBit)) == (tag | kClaimBit);
      if (!mine || ucode >= kPendingCode || lu != lt) continue;
      wins |= 1u << it;
      A.lock[t] = tag | static_cast<unsigned>(k << 14) | static_cast<unsigned>(u);
      A.lock[u] = tag | kPostIsU | static_cast<unsigned>(ku << 14) | static_cast<unsigned>(t);
    }
    __syncthreads();
    // Winners are sparse among a warp's proposals (a few lanes per 32-triangle iteration), and the rewrite is the longest
    // straight-line piece of a round: gather them into dense lanes first -- 32 at a time, across iterations -- so the
    // rewrite is issued once per 32 winners instead of once per iteration.
    auto rewrite = [&](int tk) {   // tk = (t << 2 | slot of the illegal edge), or -1
      if (tk < 0) return;
      const int t = tk >> 2, k = tk & 3;
      // (my two triangles are mine alone in this step: neighbours that flipped read my posting, not my slots)
      const unsigned ucode = DT_N(t, k);
      const int u = static_cast<int>(ucode >> 2), ku = ucode & 3;
      unsigned n_ca = DT_N(t, nx3(k)), n_ab = DT_N(t, pv3(k));
      unsigned n_bd = DT_N(u, nx3(ku)), n_dc = DT_N(u, pv3(ku));
      const unsigned short a = DT_V(t, k), bq = DT_V(t, nx3(k)), c = DT_V(t, pv3(k));
      const unsigned short d = DT_V(u, ku);
      auto across = [&](unsigned code, unsigned mine) -> unsigned {   // the neighbour's side of an outer edge after this round
        if (code >= kPendingCode) return code;                       // hull edge
        const unsigned x = code >> 2, sx = code & 3;
        const unsigned L = A.lock[x];
        if ((L & (0xFC000000u | kClaimBit)) == tag) {                // a posting of this round: x flipped too
          const unsigned partner = L & 0x3FFFu;
          // sx is x's slot kd+1 or kd+2 (kd = its slot of the old diagonal, L[15:14]); second <=> sx == (kd + 2) % 3
          const bool second = (0x214u >> (((L >> 12) & 0xCu) | sx)) & 1u;
          // first triangle of a pair (a,b,c | d): slot k+1 = edge (c,a) -> (partner, 1); slot k+2 = edge (a,b) -> (itself, 2)
          // second triangle:                      slot ku+1 = edge (b,d) -> (partner, 0); slot ku+2 = edge (d,c) -> (itself, 0)
          if (L & kPostIsU) return ((second ? x : partner) << 2) | 0u;
          return second ? ((x << 2) | 2u) : ((partner << 2) | 1u);
        }
        DT_N(x, sx) = static_cast<unsigned short>(mine);             // unchanged: re-point its slot at me
        return code;
      };
      n_bd = across(n_bd, (t << 2) | 0);
      n_ab = across(n_ab, (t << 2) | 2);
      n_dc = across(n_dc, (u << 2) | 0);
      n_ca = across(n_ca, (u << 2) | 1);
      // t <- (a,b,d), u <- (a,d,c); the new diagonal (a,d) is opposite v1 in t and opposite v2 in u
      A.v0[t] = a; A.v1[t] = bq; A.v2[t] = d;
      A.n0[t] = static_cast<unsigned short>(n_bd); A.n1[t] = static_cast<unsigned short>((u << 2) | 2);
      A.n2[t] = static_cast<unsigned short>(n_ab);
      A.v0[u] = a; A.v1[u] = d; A.v2[u] = c;
      A.n0[u] = static_cast<unsigned short>(n_dc); A.n1[u] = static_cast<unsigned short>(n_ca);
      A.n2[u] = static_cast<unsigned short>((t << 2) | 1);
      atomicOr(&A.dirty[u >> 5], 1u << (u & 31));  // t's bit is already set
    };
    int pend = -1, npend = 0;   // gathered winners: lane j < npend holds one
    it = 0;
    for (int base = 0; base < total; base += 32, ++it) {
      const bool win = (wins >> it) & 1u;
      const unsigned wm = __ballot_sync(0xffffffffu, win);
      if (!wm) continue;
      const int t = it < 4 ? (it == 0 ? tc0 : (it == 1 ? tc1 : (it == 2 ? tc2 : tc3))) : ranked(base);
      const int mine = win ? ((t << 2) | (static_cast<int>((cand >> (2 * it)) & 3) - 1)) : -1;
      const int cnt = __popc(wm);
      const int j = lane - npend;                                   // lanes npend.. take this iteration's winners in order
      const int got = __shfl_sync(0xffffffffu, mine, (j >= 0 && j < cnt) ? nth_set_bit(wm, j) : 0);
      if (j >= 0 && j < cnt) pend = got;
      if (npend + cnt >= 32) {
        rewrite(pend);
        const int j2 = lane + 32 - npend;                           // the winners that did not fit start the next set
        const int got2 = __shfl_sync(0xffffffffu, mine, j2 < cnt ? nth_set_bit(wm, j2) : 0);
        pend = j2 < cnt ? got2 : -1;
        npend += cnt - 32;
      } else {
        npend += cnt;
      }
    }
    if (npend > 0) rewrite(lane < npend ? pend : -1);
#else
    // P2: winners flip
    it = 0;
    for (int base = 0; base < total; base += 32, ++it) {
      const int sel = (cand >> (2 * it)) & 3;
      if (!__any_sync(0xffffffffu, sel != 0)) continue;  // warp-uniform: the shuffles below need every lane
      const int rank = base + lane;
      // word holding the rank-th dirty bit: smallest lane j with incl_j > rank (binary search by shuffles)
      int jsel = 0;
#pragma unroll
      for (int sstep = 16; sstep > 0; sstep >>= 1) {
        const int v = __shfl_sync(0xffffffffu, incl, jsel + sstep - 1);
        if (v <= rank) jsel += sstep;
      }
      jsel = min(jsel, 31);
      const unsigned wsel = __shfl_sync(0xffffffffu, myword, jsel);
      const int ex = __shfl_sync(0xffffffffu, excl, jsel);
      const int t = rank < total ? DT_WORD(jsel) * 32 + nth_set_bit(wsel, rank - ex) : -1;
      if (!sel) continue;
      const int k = sel - 1;
      const unsigned pri = tag | ((hash32(t * 2654435761u + round * 0x9E3779B9u) & 0xFFFu) << 14) | static_cast<unsigned>(t);
      // The claims are checked in dependency order, but the loads are issued eagerly (three dependent levels instead
      // of five): a triangle whose lock is not mine may be rewritten by its owner right now, so what is read from it
      // may be mid-update -- such values are only used as (always in-range: 14-bit) indices of further loads and
      // discarded with the claim.
      const unsigned lt = A.lock[t];
      const unsigned ucode = DT_N(t, k);
      const unsigned n_ca = DT_N(t, (k + 1) % 3), n_ab = DT_N(t, (k + 2) % 3);
      const int u = ucode < kPendingCode ? static_cast<int>(ucode >> 2) : t, ku = ucode & 3;  // (a hull marker is no index)
      const unsigned lu = A.lock[u];
      const unsigned n_bd = DT_N(u, (ku + 1) % 3), n_dc = DT_N(u, (ku + 2) % 3);
      const unsigned l1 = n_ca < kPendingCode ? A.lock[n_ca >> 2] : pri, l2 = n_ab < kPendingCode ? A.lock[n_ab >> 2] : pri;
      const unsigned l3 = n_bd < kPendingCode ? A.lock[n_bd >> 2] : pri, l4 = n_dc < kPendingCode ? A.lock[n_dc >> 2] : pri;
      if (lt != pri || ucode >= kPendingCode || lu != pri || l1 != pri || l2 != pri || l3 != pri || l4 != pri) continue;
      const unsigned short a = DT_V(t, k), bq = DT_V(t, (k + 1) % 3), c = DT_V(t, (k + 2) % 3);
      const unsigned short d = DT_V(u, ku);
      // t <- (a,b,d), u <- (a,d,c); the new diagonal (a,d) is opposite v1 in t and opposite v2 in u
      A.v0[t] = a; A.v1[t] = bq; A.v2[t] = d;
      A.n0[t] = static_cast<unsigned short>(n_bd); A.n1[t] = static_cast<unsigned short>((u << 2) | 2);
      A.n2[t] = static_cast<unsigned short>(n_ab);
      A.v0[u] = a; A.v1[u] = d; A.v2[u] = c;
      A.n0[u] = static_cast<unsigned short>(n_dc); A.n1[u] = static_cast<unsigned short>(n_ca);
      A.n2[u] = static_cast<unsigned short>((t << 2) | 1);
      link_back(A, n_bd, (t << 2) | 0);
      link_back(A, n_ab, (t << 2) | 2);
      link_back(A, n_dc, (u << 2) | 0);
      link_back(A, n_ca, (u << 2) | 1);
      atomicOr(&A.dirty[u >> 5], 1u << (u & 31));  // t's bit is already set
    }
#endif
#ifndef DT_TAIL
#define DT_TAIL 10   // enter the single-warp tail when at most this many threads proposed a flip (0 = never)
#endif
    const int nprop = __syncthreads_count(cand != 0 || deferred);
    // ---- Tail.  The last rounds of a frame (up to ~160 of the worst frame's 307) carry a handful of dirty triangles -- a
    // few cascades advancing one flip per round -- and a round then costs its fixed price: two 1024-thread barriers, two
    // rank searches, the word scans.  When few flips were proposed, warp 0 takes the whole dirty set into its lanes'
    // registers (one triangle per lane) and runs the same kind of round -- same tests, priorities, claims and winner rule,
    // only without the thinning of densely dirty words -- with warp-level synchronisation, until the set is empty or
    // outgrows the warp.  Deterministic like the block-wide rounds (nothing depends on lane order or timing).
    // Measured: 1.146 -> 1.036 ms per 64 frames at 1024^2, 1.95 -> 1.62 ms at 2048^2; thresholds 5..24 give the same.
    if (DT_TAIL > 0 && nprop <= DT_TAIL && round >= tail_hold) {
      if (warp == 0) {
        int mine = -1, n = 0;
        bool fits = true;
        for (int w0 = 0; w0 < nwords && fits; w0 += 32) {
          unsigned word = w0 + lane < nwords ? A.dirty[w0 + lane] : 0u;
          for (;;) {
            const unsigned has = __ballot_sync(0xffffffffu, word != 0u);
            if (!has) break;
            const int cnt = __popc(has);
            if (n + cnt > 32) { fits = false; break; }
            int entry = -1;
            if (word) { entry = (w0 + lane) * 32 + (__ffs(word) - 1); word &= word - 1; }
            const int j = lane - n;
            const int got = __shfl_sync(0xffffffffu, entry, (j >= 0 && j < cnt) ? nth_set_bit(has, j) : 0);
            if (j >= 0 && j < cnt) mine = got;
            n += cnt;
          }
        }
        int rr = round;
        if (fits) {
          while (n > 0 && n <= 32 && rr + 1 < max_rounds) {
            ++rr;
            if ((rr & 63) == 0) {
              for (int t = lane; t < T; t += 32) A.lock[t] = 0xFFFFFFFFu;
              __syncwarp();
            }
            const unsigned tag = static_cast<unsigned>(63 - (rr & 63)) << 26;
            const int t = lane < n ? mine : -1;
            if (t >= 0) atomicAnd(&A.dirty[t >> 5], ~(1u << (t & 31)));   // P0: snapshot and clear
            __syncwarp();
            int found = -1;
            unsigned ucode = 0, pri = 0;
            if (t >= 0) {                                                   // P1: test, claim
              const int pa = pts[DT_V(t, 0)], pb = pts[DT_V(t, 1)], pc = pts[DT_V(t, 2)];
              unsigned code3[3];
              int d3[3];
#pragma unroll
              for (int k = 0; k < 3; ++k) code3[k] = DT_N(t, k);
#pragma unroll
              for (int k = 0; k < 3; ++k) d3[k] = code3[k] < kPendingCode ? pts[DT_V(code3[k] >> 2, code3[k] & 3)] : 0;
#pragma unroll
              for (int k = 2; k >= 0; --k)
                if (code3[k] < kPendingCode && incircle_pts(pa, pb, pc, d3[k]) > 0) { found = k; ucode = code3[k]; }
              if (found >= 0) {
                atomicOr(&A.dirty[t >> 5], 1u << (t & 31));
                pri = tag | ((hash32(t * 2654435761u + rr * 0x9E3779B9u) & 0xFFFu) << 14) | static_cast<unsigned>(t);
                const int u = ucode >> 2, ku = ucode & 3;
                atomicMin(&A.lock[t], pri);
                atomicMin(&A.lock[u], pri);
                const unsigned o1 = DT_N(t, (found + 1) % 3), o2 = DT_N(t, (found + 2) % 3);
                const unsigned o3 = DT_N(u, (ku + 1) % 3), o4 = DT_N(u, (ku + 2) % 3);
                if (o1 < kPendingCode) atomicMin(&A.lock[o1 >> 2], pri);
                if (o2 < kPendingCode) atomicMin(&A.lock[o2 >> 2], pri);
                if (o3 < kPendingCode) atomicMin(&A.lock[o3 >> 2], pri);
                if (o4 < kPendingCode) atomicMin(&A.lock[o4 >> 2], pri);
              }
            }
            __syncwarp();
            int newu = -1;
            if (found >= 0) {                                               // P2: winners flip
              const int k = found;
              const unsigned lt = A.lock[t];
              const unsigned uc = DT_N(t, k);
              const unsigned n_ca = DT_N(t, (k + 1) % 3), n_ab = DT_N(t, (k + 2) % 3);
              const int u = uc < kPendingCode ? static_cast<int>(uc >> 2) : t, ku = uc & 3;
              const unsigned lu = A.lock[u];
              const unsigned n_bd = DT_N(u, (ku + 1) % 3), n_dc = DT_N(u, (ku + 2) % 3);
              const unsigned l1 = n_ca < kPendingCode ? A.lock[n_ca >> 2] : pri, l2 = n_ab < kPendingCode ? A.lock[n_ab >> 2] : pri;
              const unsigned l3 = n_bd < kPendingCode ? A.lock[n_bd >> 2] : pri, l4 = n_dc < kPendingCode ? A.lock[n_dc >> 2] : pri;
              if (lt == pri && uc < kPendingCode && lu == pri && l1 == pri && l2 == pri && l3 == pri && l4 == pri) {
                const unsigned short a = DT_V(t, k), bq = DT_V(t, (k + 1) % 3), c = DT_V(t, (k + 2) % 3);
                const unsigned short d = DT_V(u, ku);
                A.v0[t] = a; A.v1[t] = bq; A.v2[t] = d;
                A.n0[t] = static_cast<unsigned short>(n_bd); A.n1[t] = static_cast<unsigned short>((u << 2) | 2);
                A.n2[t] = static_cast<unsigned short>(n_ab);
                A.v0[u] = a; A.v1[u] = d; A.v2[u] = c;
                A.n0[u] = static_cast<unsigned short>(n_dc); A.n1[u] = static_cast<unsigned short>(n_ca);
                A.n2[u] = static_cast<unsigned short>((t << 2) | 1);
                link_back(A, n_bd, (t << 2) | 0);
                link_back(A, n_ab, (t << 2) | 2);
                link_back(A, n_dc, (u << 2) | 0);
                link_back(A, n_ca, (u << 2) | 1);
                const unsigned ubit = 1u << (u & 31);
                if (!(atomicOr(&A.dirty[u >> 5], ubit) & ubit)) newu = u;   // newly dirty: joins the set
              }
            }
            __syncwarp();
            // next set: the proposers (still dirty) first, then the winners' newly dirty partners
            const unsigned ba = __ballot_sync(0xffffffffu, found >= 0), bb = __ballot_sync(0xffffffffu, newu >= 0);
            const int ca = __popc(ba), cb = __popc(bb);
            n = ca + cb;
            if (n > 32) break;   // every member has its dirty bit set: the block-wide rounds take over
            const int jb = lane - ca;
            const int va = __shfl_sync(0xffffffffu, t, lane < ca ? nth_set_bit(ba, lane) : 0);
            const int vb = __shfl_sync(0xffffffffu, newu, (jb >= 0 && jb < cb) ? nth_set_bit(bb, jb) : 0);
            mine = lane < ca ? va : ((jb >= 0 && jb < cb) ? vb : -1);
          }
        }
        if (lane == 0) { s_tail[0] = rr; s_tail[1] = (fits && n == 0) ? 1 : 0; }
      }
      __syncthreads();
      const int rr = s_tail[0];
      const bool finished = s_tail[1] != 0;
      __syncthreads();           // s_tail may be rewritten by the next attempt
      if (finished) { round = rr + 1; flips_done = true; break; }
      tail_hold = rr + 8;        // the set outgrew a warp (or never fitted): a few block-wide rounds before the next try
      round = rr;
    }
  }
#undef DT_V
#undef DT_N
#undef DT_WORD

  if (dbg && tid == 0) dbg[b * 8 + 0] = (int)((clock64() - clk0) >> 4);
  // A loop that ran into its safety bound (flips: max_rounds; pockets: 4R + 64 rounds) leaves a mesh that is not
  // Delaunay.  Never observed; if it happens the frame reports NO mesh (ntri = 0: every unfilled pixel stays NaN) and
  // rounds = -1, which fovea.ops.check_plan turns into an exception -- never a silently wrong interpolation.
  converged = converged && flips_done;
  if (!converged) {
    if (tid == 0) { ntri_out[b] = 0; if (rounds_out) rounds_out[b] = -1; }
    if (hints_out) {
      const int nh = ceil_div(H, FOVEA_HINT_CELL_H) * ceil_div(W, FOVEA_HINT_CELL_W);
      for (int i = tid; i < nh; i += kDtThreads) hints_out[static_cast<size_t>(b) * nh + i] = 0;
    }
    return;
  }
  // ------------------------------------------------------------------ output: 16-byte records (v0,v1,v2,0,n0,n1,n2,0)
  for (int t = tid; t < T; t += kDtThreads) {
    const unsigned c0 = A.n0[t], c1 = A.n1[t], c2 = A.n2[t];
    uint4 rec;
    rec.x = static_cast<unsigned>(A.v0[t]) | (static_cast<unsigned>(A.v1[t]) << 16);
    rec.y = static_cast<unsigned>(A.v2[t]);
    rec.z = (c0 >= kPendingCode ? 0xFFFFu : (c0 >> 2)) | ((c1 >= kPendingCode ? 0xFFFFu : (c1 >> 2)) << 16);
    rec.w = (c2 >= kPendingCode ? 0xFFFFu : (c2 >> 2));
    reinterpret_cast<uint4*>(mesh)[t] = rec;
  }
  if (tid == 0) {
    ntri_out[b] = T;
    if (rounds_out) rounds_out[b] = round;
  }

  // ------------------------------------------------------------------ walk-start hints (fovea_locate_hints, fused)
  // The mesh is still in shared memory: the three coarse-to-fine levels of locate_hints_kernel run here without
  // re-staging it.  A hint only has to be NEAR its cell centre (fovea_locate_pixels walks from it), so the walk below
  // stops at the first triangle with no negative edge function (no tie rule needed).  Scratch: the dead lock array.
  if (hints_out) {
    __syncthreads();
    const int ch = ceil_div(H, FOVEA_HINT_CELL_H), cw = ceil_div(W, FOVEA_HINT_CELL_W);
    const int mh = ceil_div(ch, 4);
    int* coarse = reinterpret_cast<int*>(A.lock);  // [16*16]
    int* mid = coarse + 256;                       // [mh*cw]   (the host checked that it fits)
    int32_t* hb = hints_out + static_cast<size_t>(b) * ch * cw;
    auto walk = [&](int qr, int qc, int t) {
      for (int step = 0; step < T + 8; ++step) {
        const int p0 = pts[A.v0[t]], p1 = pts[A.v1[t]], p2 = pts[A.v2[t]];
        const int q = (qr << 16) | qc;
        unsigned code;
        if (orient_pts(p1, p2, q) < 0) code = A.n0[t];         // triangles are counter-clockwise: inside <=> all >= 0
        else if (orient_pts(p2, p0, q) < 0) code = A.n1[t];
        else if (orient_pts(p0, p1, q) < 0) code = A.n2[t];
        else return t;
        if (code >= kPendingCode) return t;                    // left the hull: the last triangle is the nearest known
        t = static_cast<int>(code >> 2);
      }
      return t;
    };
    if (tid < 256) {
      const int py = tid / 16, px = tid % 16;
      coarse[tid] = walk(min(H - 1, (2 * py + 1) * H / 32), min(W - 1, (2 * px + 1) * W / 32), 0);
    }
    __syncthreads();
    for (int i = tid; i < mh * cw; i += kDtThreads) {
      const int cy = i / cw, cx = i - cy * cw;
      const int qr = min(H - 1, cy * 32 + 16), qc = min(W - 1, cx * FOVEA_HINT_CELL_W + FOVEA_HINT_CELL_W / 2);
      mid[i] = walk(qr, qc, coarse[min(15, qr * 16 / H) * 16 + min(15, qc * 16 / W)]);
    }
    __syncthreads();
    for (int i = tid; i < ch * cw; i += kDtThreads) {
      const int cy = i / cw, cx = i - cy * cw;
      hb[i] = walk(min(H - 1, cy * FOVEA_HINT_CELL_H + FOVEA_HINT_CELL_H / 2),
                   min(W - 1, cx * FOVEA_HINT_CELL_W + FOVEA_HINT_CELL_W / 2), mid[(cy / 4) * cw + cx]);
    }
  }
}

constexpr int kDtMaxTri = 16383;     // triangle ids are stored as (id << 2 | slot) in 16 bits; 0xFFFD..0xFFFF are markers
constexpr size_t kDtMaxSmem = 227 * 1024;

// triangles the kernel can address for this capacity (a full triangulation of n sites has at most 2n - 5 of them)
static int dt_kernel_tcap(int tcap) { return tcap < kDtMaxTri ? tcap : kDtMaxTri; }
static size_t dt_smem_bytes(int cap, int tk, bool big) {
  return static_cast<size_t>(12) * tk + (big ? 0 : 4 * static_cast<size_t>(tk) + 4 * static_cast<size_t>((tk + 31) / 32)) +
         4 * static_cast<size_t>(cap);
}
static bool dt_big(int cap, int tcap) { return dt_smem_bytes(cap, dt_kernel_tcap(tcap), false) > kDtMaxSmem; }

}  // namespace fovea

using namespace fovea;

extern "C" int64_t fovea_delaunay_workspace_bytes(int B, int cap) {
  // [B] flip rounds + [B,8] stage counters (int32), then [B, cap+2] row starts (uint16); site sets too large for the
  // all-shared-memory layout add [B, tk + tk/32] claim words + dirty bits (uint32)
  int64_t bytes = static_cast<int64_t>(B) * 4 * 9 + (static_cast<int64_t>(B) * (cap + 2) * 2 + 3) / 4 * 4;
  const int tk = dt_kernel_tcap(2 * cap);
  if (dt_big(cap, 2 * cap)) bytes += static_cast<int64_t>(B) * (tk + (tk + 31) / 32) * 4;
  return bytes;
}

static int launch_delaunay(const int32_t* pts, const int32_t* npts, int B, int cap, int tcap, int max_coord,
                           uint16_t* mesh, int32_t* ntri, void* workspace, int32_t* hints, int H, int W,
                           cudaStream_t stream, const char* who) {
  FOVEA_REQUIRE(pts && npts && mesh && ntri && workspace, "%s: null pointer", who);
  FOVEA_REQUIRE(B > 0 && cap >= 4 && tcap >= 2 * cap, "%s: need tcap >= 2*cap (cap=%d tcap=%d)", who, cap, tcap);
  FOVEA_REQUIRE(max_coord > 0 && max_coord <= 8192,
                "%s: coordinates must be < 8192 for the exact int64 in-circle test (got %d)", who, max_coord);
  if (cap > 8196) {   // (8196 = the 64 x 128 lattice of config/deform.yaml + the four corners)
    set_error("%s: cap=%d exceeds the 16-bit mesh encoding (at most 8196 sites, 16383 triangles)", who, cap);
    return FOVEA_ERR_CAPACITY;
  }
  const int tk = dt_kernel_tcap(tcap);
  const bool big = dt_big(cap, tcap);
  FOVEA_REQUIRE(!(big && hints), "%s: the fused walk hints need the all-shared-memory layout (cap=%d)", who, cap);
  const size_t smem = dt_smem_bytes(cap, tk, big);
  if (smem > kDtMaxSmem) {
    set_error("%s: %zu B of shared memory needed for cap=%d (> 227 KB); use the host triangulation", who, smem, cap);
    return FOVEA_ERR_CAPACITY;
  }
  FOVEA_CUDA(cudaFuncSetAttribute(big ? delaunay_kernel<true> : delaunay_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  // safety bound on the flip rounds (worst frame seen: ~310); FOVEA_DT_MAX_ROUNDS lowers it to exercise the
  // non-convergence report in the tests
  int max_rounds = 20000;
  if (const char* e = getenv("FOVEA_DT_MAX_ROUNDS")) max_rounds = atoi(e) > 0 ? atoi(e) : max_rounds;
  int32_t* ws32 = static_cast<int32_t*>(workspace);
  unsigned short* row_ws = reinterpret_cast<unsigned short*>(ws32 + 9 * B);
  unsigned* lock_g = big ? reinterpret_cast<unsigned*>(ws32 + 9 * B + (static_cast<size_t>(B) * (cap + 2) * 2 + 3) / 4) : nullptr;
  if (big)
    delaunay_kernel<true><<<B, kDtThreads, smem, stream>>>(pts, npts, cap, tk, mesh, ntri, ws32, max_rounds, ws32 + B, row_ws, hints, H,
                                                   W, tcap, lock_g);
  else
    delaunay_kernel<false><<<B, kDtThreads, smem, stream>>>(pts, npts, cap, tk, mesh, ntri, ws32, max_rounds, ws32 + B, row_ws, hints, H,
                                                   W, tcap, lock_g);
  return check_launch(who);
}

extern "C" int fovea_delaunay(const int32_t* pts, const int32_t* npts, int B, int cap, int tcap, int max_coord,
                              uint16_t* mesh, int32_t* ntri, void* workspace, fovea_stream_t stream) {
  return launch_delaunay(pts, npts, B, cap, tcap, max_coord, mesh, ntri, workspace, nullptr, 0, 0,
                         static_cast<cudaStream_t>(stream), "fovea_delaunay");
}

extern "C" int fovea_delaunay_hints_fused(int tcap, int H, int W) {
  // the two coarse levels live in the kernel's dead lock array (4*tcap bytes) -- when that array is in shared memory
  if (dt_big(tcap / 2, tcap)) return 0;
  const long long ch = ceil_div(H, FOVEA_HINT_CELL_H), cw = ceil_div(W, FOVEA_HINT_CELL_W);
  return (256 + ceil_div(static_cast<int>(ch), 4) * cw) * 4 <= 4ll * tcap ? 1 : 0;
}

extern "C" int fovea_delaunay_with_hints(const int32_t* pts, const int32_t* npts, int B, int cap, int tcap, int H,
                                         int W, uint16_t* mesh, int32_t* ntri, int32_t* hints, void* workspace,
                                         fovea_stream_t stream) {
  FOVEA_REQUIRE(hints && H > 1 && W > 1, "fovea_delaunay_with_hints: bad arguments");
  FOVEA_REQUIRE(fovea_delaunay_hints_fused(tcap, H, W), "fovea_delaunay_with_hints: %dx%d hints do not fit the kernel's "
                "scratch (tcap=%d); call fovea_delaunay + fovea_locate_hints", H, W, tcap);
  return launch_delaunay(pts, npts, B, cap, tcap, H > W ? H : W, mesh, ntri, workspace, hints, H, W,
                         static_cast<cudaStream_t>(stream), "fovea_delaunay_with_hints");
}

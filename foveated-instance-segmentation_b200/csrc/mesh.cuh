// Triangle mesh access + exact point location shared by the stage-3 kernels.
//
// Points are packed (row<<16 | col); a triangle is one 16-byte record of uint16 (v0,v1,v2,0,n0,n1,n2,0): v* index
// the per-image point list, n_k is the triangle across the edge OPPOSITE vertex k (SciPy's `neighbors` convention,
// spatial/qhull.pyx:1367-1380), 0xFFFF = convex-hull edge.  Orientation may be either sign (host Qhull emits
// both); all predicates are exact integer arithmetic.
#pragma once
#include "common.cuh"

namespace fovea {

constexpr unsigned kNoTri = 0xFFFFu;

struct Mesh {
  const int32_t* pts;
  const uint4* rec;  // global or shared memory
  int ntri;
};

__device__ __forceinline__ long long orient2d(int ar, int ac, int br, int bc, int qr, int qc) {
  // z of (b-a) x (q-a) with x = col, y = row
  return static_cast<long long>(bc - ac) * (qr - ar) - static_cast<long long>(br - ar) * (qc - ac);
}

struct Located {
  int tri;            // -1: outside the triangulation
  long long a0, a1;   // orientation-normalised sub-areas opposite vertex 0 and 1
  long long area;     // |orient(v0,v1,v2)|
  unsigned v01, v2;    // packed vertex ids (v0 | v1<<16, v2)
};

// Tie rule for a query that lies EXACTLY on an edge (or vertex): the pixel belongs to the triangle that contains
// the symbolically perturbed point q + (d_row = eps^2, d_col = -eps) -- a top-left fill rule, so every pixel inside
// the hull has exactly one owner and the result does not depend on where the walk started.  (The reference's
// find_simplex accepts whichever eps-tolerant triangle its sequential warm-started walk meets first,
// spatial/qhull.pyx:1367-1467; this rule reproduces that choice for ~82% of on-edge pixels.  The choice only
// shows where one of the two triangles has a NaN vertex, because the interpolant is continuous across edges.)
__device__ __forceinline__ bool edge_owned(long long e, long long s, int ar, int ac, int br, int bc) {
  if (e != 0) return e > 0;
  const int dr = br - ar;
  return dr != 0 ? (s * dr > 0) : (s * (bc - ac) > 0);
}

__device__ __forceinline__ bool test_triangle(const Mesh& m, int t, int qr, int qc, Located& out, int& next) {
  const uint4 rec = m.rec[t];
  const int p0 = m.pts[rec.x & 0xFFFFu], p1 = m.pts[rec.x >> 16], p2 = m.pts[rec.y & 0xFFFFu];
  const int r0 = p0 >> 16, c0 = p0 & 0xFFFF, r1 = p1 >> 16, c1 = p1 & 0xFFFF, r2 = p2 >> 16, c2 = p2 & 0xFFFF;
  long long A = orient2d(r0, c0, r1, c1, r2, c2);
  const long long s = A < 0 ? -1 : 1;
  A *= s;
  const long long e0 = s * orient2d(r1, c1, r2, c2, qr, qc);
  const long long e1 = s * orient2d(r2, c2, r0, c0, qr, qc);
  const long long e2 = A - e0 - e1;  // == s*orient(p0,p1,q)
  next = -1;
  if (A == 0) return false;
  const bool strict = e0 > 0 && e1 > 0 && e2 > 0;  // the common case
  if (!strict) {
    const unsigned n0 = rec.z & 0xFFFFu, n1 = rec.z >> 16, n2 = rec.w & 0xFFFFu;
    // an edge on the convex hull owns its pixels (there is no neighbour to hand them to)
    if (!(edge_owned(e0, s, r1, c1, r2, c2) || (e0 == 0 && n0 == kNoTri))) { next = n0; return false; }
    if (!(edge_owned(e1, s, r2, c2, r0, c0) || (e1 == 0 && n1 == kNoTri))) { next = n1; return false; }
    if (!(edge_owned(e2, s, r0, c0, r1, c1) || (e2 == 0 && n2 == kNoTri))) { next = n2; return false; }
  }
  out.tri = t; out.a0 = e0; out.a1 = e1; out.area = A; out.v01 = rec.x; out.v2 = rec.y & 0xFFFFu;
  return true;
}

// Visibility walk from `start`; falls back to a scan of all triangles if the walk meets a degenerate
// triangle or exceeds its step budget (cannot happen on a Delaunay mesh, may on a host mesh with flat facets).
__device__ __noinline__ Located locate_bruteforce(const Mesh& m, int qr, int qc) {
  Located L; L.tri = -1; L.a0 = L.a1 = 0; L.area = 1; L.v01 = 0; L.v2 = 0;
  int next;
  for (int t = 0; t < m.ntri; ++t)
    if (test_triangle(m, t, qr, qc, L, next)) return L;
  L.tri = -1;
  return L;
}

__device__ __forceinline__ Located locate(const Mesh& m, int qr, int qc, int start) {
  Located L; L.tri = -1; L.a0 = L.a1 = 0; L.area = 1; L.v01 = 0; L.v2 = 0;
  int t = (start >= 0 && start < m.ntri) ? start : 0;
  const int budget = m.ntri + 8;
  for (int step = 0; step < budget; ++step) {
    int next;
    if (test_triangle(m, t, qr, qc, L, next)) return L;
    if (next < 0) break;                                   // degenerate triangle
    if (static_cast<unsigned>(next) == kNoTri) { L.tri = -1; return L; }  // left the hull
    t = next;
  }
  return locate_bruteforce(m, qr, qc);
}

}  // namespace fovea

// Stage 3 in MASK mode: torch.argmax(pred_sampled, dim=1) (models/models.py:1044) without interpolating all C channels.
//
// The reference materialises pred_sampled [B,C,H,W] (models/models.py:933-940) and arg-maxes it; the fused argmax of
// fovea_inverse_fill still interpolates every channel of every pixel to emit one index (1.98 ms per 64 frames of
// 1024^2 for 8 bytes per pixel written: 4 % of the HBM peak, bound by L1 loads and FP issue).  This file prunes first:
//
//   * a pixel that received a node carries that node's table row unchanged -> its label is the row's argmax, computed
//     once per NODE (node_argmax_kernel);
//   * inside a triangle every channel's score is  fl(fl(fl(a*w0) + fl(b*w1)) + fl(c*w2))  with weights >= 0 and IEEE
//     rounding monotone, so a channel that is <= another channel at all three vertices can never beat it at any pixel
//     of the triangle.  triangle_candidates_kernel keeps, per TRIANGLE, the channels that survive this dominance test
//     against the two strongest channels of each vertex (typically 3-10 of 51 on i.i.d. predictions, 2 on the reference's
//     C1 decoder) together with their three vertex values; inverse_mask_kernel evaluates only those -- with exactly the
//     arithmetic of fovea_inverse_fill.
//
// Exactness (the mask equals the fused argmax of fovea_inverse_fill bit for bit, ties included):
//   - torch.argmax returns the FIRST maximum.  A channel may be dropped because of a LOWER-index channel that is >= at
//     all three vertices (that one wins every tie anyway); because of a HIGHER-index channel only if that one is
//     greater at all three vertices by a margin (2^-20 of the larger magnitude) that exceeds the worst-case rounding of
//     both evaluations (3 roundings of relative 2^-24 each, weights summing to 1 +- 2^-23), so it is STRICTLY greater
//     at every pixel;
//   - all three weights are >= 0: w0, w1 are quotients of non-negative integers; the third, float(1 - c0 - c1), is pure
//     rounding noise (|.| ~ 1e-16) when the pixel lies exactly on the edge opposite vertex 2 and is clamped at zero by
//     every stage-3 kernel (fill_tile) -- the reference's own value there is noise of its own LU-based formula;
//   - a triangle whose survivors overflow the 16-slot record evaluates all C channels (0.3 % of the triangles on
//     i.i.d. N(0,1) predictions);
//   - survivors are kept in increasing channel order and compared with `>`, like the full loop.
#include <math_constants.h>

#include "common.cuh"
#include "fill.cuh"

namespace fovea {

constexpr int kCandMax = 32;          // survivors kept per triangle (one 512-byte record, only the used slots are ever
                                      // touched); more -> full evaluation.  (16 slots overflow for ~10 % of the huge hull
                                      // triangles, whose three nodes are far apart -- and those cover a fifth of a canvas.)
constexpr int kCandThreads = 128;
constexpr int kCandRefine = 10;       // lists longer than this are pruned exactly (warp-cooperatively)
constexpr unsigned kCandFull = 0xFFu;  // ncand marker: evaluate every channel
constexpr unsigned kCandNaN = 0xFEu;   // ncand marker: a vertex has no value -> NaN in every channel -> label 0

// label of a table row = argmax over its C channels (first maximum; rows h*w (NaN) and h*w+1 (zeros) give 0)
__global__ void __launch_bounds__(256)
node_argmax_kernel(const float* __restrict__ table, uint8_t* __restrict__ nodearg, int rows, int C, int Cs) {
  const int b = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* t = table + (static_cast<size_t>(b) * rows + r) * Cs;
  float best = __ldg(t);
  int bi = 0;
  for (int c = 1; c < C; ++c) {
    const float v = __ldg(t + c);
    if (v > best) { best = v; bi = c; }
  }
  nodearg[static_cast<size_t>(b) * rows + r] = static_cast<uint8_t>(best == best ? bi : 0);
}

// One thread per triangle, uniform control flow (the 32 triangles of a warp walk the same C channels in lock step; an
// earlier version kept an exact survivor list per thread and ran at 8-11 active lanes per instruction).
//   pass 1: per vertex the two largest channels (value and index) -> up to six "seeds", and the margin
//           M = 2^-20 * (largest magnitude among the triangle's vertex values);
//   pass 2: a channel is dropped if some seed makes it irrelevant on the whole triangle:
//             seed index lower : seed >= channel at the three vertices (the seed wins every tie as well);
//             seed index higher: seed exceeds the channel at the three vertices by more than M -- more than the rounding of
//                                both evaluations, i.e. strictly greater at every pixel;
//           what is left is written out in index order.  The seeds are real channels, so being beaten by one is reason
//           enough whatever becomes of the seed itself (the relation is transitive).  On i.i.d. N(0,1) predictions 6.8
//           channels are left per triangle where an exact all-pairs pruning leaves 5.9.
__global__ void __launch_bounds__(kCandThreads)
triangle_candidates_kernel(const TriRec* __restrict__ trirec, const int32_t* __restrict__ ntri,
                           const float* __restrict__ table, float4* __restrict__ cand, uint8_t* __restrict__ ncand,
                           int hw, int C, int Cs, int tcap) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * kCandThreads + threadIdx.x;
  const bool exists = t < ntri[b];            // (no early return: pass 3 needs every lane of the warp)
  const uint4 q3 = exists ? __ldg(reinterpret_cast<const uint4*>(trirec + static_cast<size_t>(b) * tcap + t) + 3)
                          : make_uint4(0, 0, 0, 0);
  const int r0 = static_cast<int>(q3.x & 0xFFFFu), r1 = static_cast<int>(q3.x >> 16), r2 = static_cast<int>(q3.y);
  uint8_t* nc = ncand + static_cast<size_t>(b) * tcap + (exists ? t : 0);
  const bool valid = exists && r0 < hw && r1 < hw && r2 < hw;   // a vertex without a value: NaN in every channel
  float M = 0.f;
  float4* out = cand + (static_cast<size_t>(b) * tcap + (exists ? t : 0)) * kCandMax;
  int n = 0;
  if (valid) {
  const float* tb = table + static_cast<size_t>(b) * (hw + 2) * Cs;
  const float4* p0 = reinterpret_cast<const float4*>(tb + static_cast<size_t>(r0) * Cs);
  const float4* p1 = reinterpret_cast<const float4*>(tb + static_cast<size_t>(r1) * Cs);
  const float4* p2 = reinterpret_cast<const float4*>(tb + static_cast<size_t>(r2) * Cs);
  // ---- pass 1
  float mag = 0.f;
  float top[3][2] = {{-CUDART_INF_F, -CUDART_INF_F}, {-CUDART_INF_F, -CUDART_INF_F}, {-CUDART_INF_F, -CUDART_INF_F}};
  int topi[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  for (int c4 = 0; c4 < C; c4 += 4) {
    const float4 A = __ldg(p0 + (c4 >> 2)), Bv = __ldg(p1 + (c4 >> 2)), Cv = __ldg(p2 + (c4 >> 2));
    const float v[3][4] = {{A.x, A.y, A.z, A.w}, {Bv.x, Bv.y, Bv.z, Bv.w}, {Cv.x, Cv.y, Cv.z, Cv.w}};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool in = c4 + e < C;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float x = v[k][e];
        mag = fmaxf(mag, in ? fabsf(x) : 0.f);
        const bool g1 = in & (x > top[k][0]), g2 = in & (x > top[k][1]);
        top[k][1] = g1 ? top[k][0] : (g2 ? x : top[k][1]);
        topi[k][1] = g1 ? topi[k][0] : (g2 ? c4 + e : topi[k][1]);
        top[k][0] = g1 ? x : top[k][0];
        topi[k][0] = g1 ? c4 + e : topi[k][0];
      }
    }
  }
  M = 9.5367431640625e-07f * mag;  // 2^-20 * mag   (NaN / Inf magnitudes: no comparison below succeeds)
  // the seeds' values at the three vertices (nine scalar loads each way; duplicates among the six are harmless)
  const float* f0 = reinterpret_cast<const float*>(p0);
  const float* f1 = reinterpret_cast<const float*>(p1);
  const float* f2 = reinterpret_cast<const float*>(p2);
  int si[6];
  float sa[6], sb[6], sc[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    si[q] = topi[q >> 1][q & 1];
    sa[q] = __ldg(f0 + si[q]); sb[q] = __ldg(f1 + si[q]); sc[q] = __ldg(f2 + si[q]);
  }
  // ---- pass 2
  for (int c4 = 0; c4 < C; c4 += 4) {
    const float4 A = __ldg(p0 + (c4 >> 2)), Bv = __ldg(p1 + (c4 >> 2)), Cv = __ldg(p2 + (c4 >> 2));
    const float a4[4] = {A.x, A.y, A.z, A.w}, b4[4] = {Bv.x, Bv.y, Bv.z, Bv.w}, c4v[4] = {Cv.x, Cv.y, Cv.z, Cv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = c4 + e;
      const float a = a4[e], bb = b4[e], cc = c4v[e];
      bool beaten = c >= C;
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        // (bitwise & | on purpose: short-circuit operators compile to branches, and the lanes of a warp -- 32 different
        //  triangles -- would take them apart)
        const bool lower = si[q] < c, higher = si[q] > c;
        beaten = beaten | (lower & (sa[q] >= a) & (sb[q] >= bb) & (sc[q] >= cc)) |
                 (higher & (sa[q] - M > a) & (sb[q] - M > bb) & (sc[q] - M > cc));
      }
      if (!beaten) {
        if (n < kCandMax) out[n] = make_float4(a, bb, cc, __int_as_float(c));
        ++n;
      }
    }
  }
  }
  // ---- pass 3: long lists (the large hull / periphery triangles, whose three nodes are far apart and whose strongest
  // channels beat little -- and which cover a good part of the canvas) are pruned exactly, one after the other, by the
  // whole warp: lane i holds entry i and drops it if ANY other entry makes it irrelevant (same two rules)
  const int lane = threadIdx.x & 31;
  const unsigned peers = 0xffffffffu;
  unsigned todo = __ballot_sync(peers, n > kCandRefine && n <= kCandMax);
  __syncwarp(peers);   // the lists were written by their owners: order those stores before the peers' loads
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const int ns = __shfl_sync(peers, n, src);
    const float Ms = __shfl_sync(peers, M, src);
    const unsigned long long base = __shfl_sync(peers, reinterpret_cast<unsigned long long>(out), src);
    float4* list = reinterpret_cast<float4*>(base);
    const float4 me = lane < ns ? list[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    const int my = __float_as_int(me.w);
    bool keep = lane < ns;
    for (int j = 0; j < ns; ++j) {
      const float ja = __shfl_sync(peers, me.x, j), jb = __shfl_sync(peers, me.y, j), jc = __shfl_sync(peers, me.z, j);
      const int jd = __shfl_sync(peers, my, j);
      const bool lower = jd < my, higher = jd > my;
      keep = keep & !((lower & (ja >= me.x) & (jb >= me.y) & (jc >= me.z)) |
                      (higher & (ja - Ms > me.x) & (jb - Ms > me.y) & (jc - Ms > me.z)));
    }
    const unsigned kept = __ballot_sync(peers, keep);
    __syncwarp(peers);                               // every lane has read its entry before the list is rewritten
    if (keep) list[__popc(kept & ((1u << lane) - 1u))] = me;
    if (lane == src) n = __popc(kept);
    __syncwarp(peers);
  }
  if (exists) *nc = static_cast<uint8_t>(!valid ? kCandNaN : (n <= kCandMax ? n : kCandFull));
}

constexpr int kMaskThreads = 256;
// Lanes of a warp across a row (x 4 pixels each).  The candidate loop runs to the longest list among the warp's
// triangles, so a compact footprint (fewer triangles per warp) is worth more here than the long row segments the
// store-bound score fill wants: pruned mask fill on N(0,1) predictions 1.197 / 1.155 / 1.099 / 1.122 ms for
// WL = 16 / 8 / 4 / 2 (64 x 2, 32 x 4, 16 x 8, 8 x 16 pixels per warp).
#ifndef FOVEA_MASK_WL
#define FOVEA_MASK_WL 4
#endif
constexpr int kMaskWL = FOVEA_MASK_WL;

__device__ __forceinline__ float interp3(float a, float b, float c, float w0, float w1, float w2) {
  // interp2d.py:85-89 as fill_tile evaluates it: three products, summed in vertex order, every step rounded
  return __fadd_rn(__fadd_rn(__fmul_rn(a, w0), __fmul_rn(b, w1)), __fmul_rn(c, w2));
}

__global__ void __launch_bounds__(kMaskThreads, 3)
inverse_mask_kernel(const uint16_t* __restrict__ loc, const TriRec* __restrict__ trirec, const float* __restrict__ table,
                    const float4* __restrict__ cand, const uint8_t* __restrict__ ncand,
                    const uint8_t* __restrict__ nodearg, void* __restrict__ mask, FillParams p) {
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int WX = 2, kWarpW = 4 * kMaskWL, kWarpH = 32 / kMaskWL, kTileH = kWarpH * (kMaskThreads / 32 / WX);
  const int x0 = blockIdx.x * (kWarpW * WX) + (warp % WX) * kWarpW + (lane % kMaskWL) * 4;
  const int y = blockIdx.y * kTileH + (warp / WX) * kWarpH + (lane / kMaskWL);
  if (x0 >= p.W || y >= p.H) return;
  const int hw = p.h * p.w;
  const size_t plane = static_cast<size_t>(p.H) * p.W;
  const unsigned pixoff = static_cast<unsigned>(y) * p.W + x0;
  const TriRec* recs = trirec + static_cast<size_t>(b) * p.tcap;
  const uint8_t* na = nodearg + static_cast<size_t>(b) * (hw + 2);

  // ---- phase 1: what produces each of my four pixels (the arithmetic of fill_tile, inverse.cu)
  const uint2 l2 = __ldcs(reinterpret_cast<const uint2*>(loc + static_cast<size_t>(b) * plane + pixoff));
  const int lc[4] = {decode_loc(l2.x & 0xFFFFu), decode_loc(l2.x >> 16), decode_loc(l2.y & 0xFFFFu), decode_loc(l2.y >> 16)};
  int label[4] = {0, 0, 0, 0};
  float w0[4], w1[4], w2[4];
  int cur = -1, e0 = 0, e1 = 0, d0 = 0, d1 = 0;
  double inv_area = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    w0[k] = w1[k] = w2[k] = 0.f;
    if (lc[k] < 0) {                       // the pixel received a node: the label of that node's row
      label[k] = na[-(lc[k] + 1)];
    } else {
      if (lc[k] != cur) {
        cur = lc[k];
        const uint4* r = reinterpret_cast<const uint4*>(recs + cur);
        const uint4 q0 = __ldg(r), q1 = __ldg(r + 1), q3 = __ldg(r + 3);
        d0 = static_cast<int>(q0.y);
        d1 = static_cast<int>(q1.x);
        e0 = static_cast<int>(q0.x) * y + d0 * (x0 + k) + static_cast<int>(q0.z);
        e1 = static_cast<int>(q0.w) * y + d1 * (x0 + k) + static_cast<int>(q1.y);
        inv_area = __hiloint2double(static_cast<int>(q3.w), static_cast<int>(q3.z));
      }
      // interp2d.py:58-65 / qhull.pyx:1210-1264: c0, c1 in float64, c2 = 1 - c0 - c1, then cast to float32
      const double c0 = static_cast<double>(e0) * inv_area, c1 = static_cast<double>(e1) * inv_area;
      w0[k] = static_cast<float>(c0); w1[k] = static_cast<float>(c1);
      w2[k] = fmaxf(static_cast<float>(1.0 - c0 - c1), 0.f);   // (see fill_tile: rounding noise below zero is clamped)
    }
    e0 += d0;
    e1 += d1;
  }

  // ---- phase 2: every pixel walks the surviving channels of ITS triangle in one loop that all lanes share (a pixel in
  // the same triangle as its left neighbour reuses the neighbour's record instead of loading it again)
  unsigned nk[4], raw[4];
  const float4* cl[4];
  float best[4] = {0.f, 0.f, 0.f, 0.f};
  unsigned nmax = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    nk[k] = 0;
    raw[k] = 0;
    cl[k] = cand;
    if (lc[k] < 0) continue;
    raw[k] = (k > 0 && lc[k] == lc[k - 1]) ? raw[k - 1] : ncand[static_cast<size_t>(b) * p.tcap + lc[k]];
    cl[k] = cand + (static_cast<size_t>(b) * p.tcap + lc[k]) * kCandMax;
    if (raw[k] == kCandNaN) continue;      // NaN (or zeroed) in every channel: torch.argmax gives 0
    if (raw[k] == kCandFull) {             // more survivors than a record holds (rare): every channel, from the table
      const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(recs + lc[k]) + 3);
      const float* tb = table + static_cast<size_t>(b) * (hw + 2) * p.Cs;
      const float* ra = tb + static_cast<size_t>(q3.x & 0xFFFFu) * p.Cs;
      const float* rb = tb + static_cast<size_t>(q3.x >> 16) * p.Cs;
      const float* rc = tb + static_cast<size_t>(q3.y) * p.Cs;
      float bst = 0.f;
      for (int c = 0; c < p.C; ++c) {
        const float v = interp3(__ldg(ra + c), __ldg(rb + c), __ldg(rc + c), w0[k], w1[k], w2[k]);
        if (c == 0 || v > bst) { bst = v; label[k] = c; }
      }
      continue;
    }
    nk[k] = raw[k];
    nmax = max(nmax, raw[k]);
  }
  bool own[4];   // the pixel loads its triangle's record itself (its left neighbour lies in another triangle)
#pragma unroll
  for (int k = 0; k < 4; ++k) own[k] = (k == 0) | (lc[k] != lc[max(k - 1, 0)]);
  for (unsigned j = 0; j < nmax; ++j) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {          // branch-free: predicated load, selects
      const bool live = j < nk[k];
      if (live & own[k]) s = __ldg(cl[k] + j);
      const float v = interp3(s.x, s.y, s.z, w0[k], w1[k], w2[k]);
      const bool take = live & ((j == 0) | (v > best[k]));
      best[k] = take ? v : best[k];
      label[k] = take ? __float_as_int(s.w) : label[k];
    }
  }
  if (p.mask_u8) {
    *reinterpret_cast<uchar4*>(static_cast<unsigned char*>(mask) + static_cast<size_t>(b) * plane + pixoff) =
        make_uchar4(label[0], label[1], label[2], label[3]);
  } else {
    longlong2* mp = reinterpret_cast<longlong2*>(static_cast<long long*>(mask) + static_cast<size_t>(b) * plane + pixoff);
    __stcs(mp, make_longlong2(label[0], label[1]));
    __stcs(mp + 1, make_longlong2(label[2], label[3]));
  }
}

}  // namespace fovea

using namespace fovea;

// workspace layout: [B, tcap, kCandMax] float4 survivors | [B, tcap] uint8 counts | [B, h*w+2] uint8 node labels
static size_t cand_bytes(int B, int tcap) { return static_cast<size_t>(B) * tcap * kCandMax * sizeof(float4); }

extern "C" int64_t fovea_inverse_mask_workspace_bytes(int B, int h, int w, int tcap) {
  const size_t rows = static_cast<size_t>(h) * w + 2;
  return static_cast<int64_t>(cand_bytes(B, tcap) + ((static_cast<size_t>(B) * tcap + 15) / 16) * 16 + B * rows + 16);
}

extern "C" int fovea_inverse_mask(const uint16_t* loc, const void* trirec, const int32_t* ntri, const float* table,
                                  int B, int C, int Cs, int h, int w, int H, int W, int tcap, void* workspace,
                                  void* mask, int mask_u8, fovea_stream_t stream) {
  FOVEA_REQUIRE(loc && trirec && table && workspace && mask, "fovea_inverse_mask: null pointer");
  FOVEA_REQUIRE(B > 0 && C > 0 && C <= 256 && Cs >= C && Cs % 4 == 0 && h > 0 && w > 0 && H > 1 && W > 1 && tcap > 0,
                "fovea_inverse_mask: bad sizes (labels are staged as bytes: C <= 256)");
  FOVEA_REQUIRE(W % 4 == 0, "fovea_inverse_mask: W=%d must be a multiple of 4", W);
  FOVEA_REQUIRE(H <= 16384 && W <= 16384 && B <= 65535, "fovea_inverse_mask: canvas or batch too large");
  FOVEA_REQUIRE(static_cast<long long>(h) * w + 2 <= 32768, "fovea_inverse_mask: value table rows must fit 15 bits");
  FOVEA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0 && (reinterpret_cast<uintptr_t>(table) & 15u) == 0,
                "fovea_inverse_mask: workspace and table must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rows = h * w + 2;
  float4* cand = static_cast<float4*>(workspace);
  uint8_t* ncand = reinterpret_cast<uint8_t*>(workspace) + cand_bytes(B, tcap);
  uint8_t* nodearg = ncand + ((static_cast<size_t>(B) * tcap + 15) / 16) * 16;
  node_argmax_kernel<<<dim3(ceil_div(rows, 256), B), 256, 0, s>>>(table, nodearg, rows, C, Cs);
  if (ntri)  // 'nearest' plans carry no triangles: every pixel is a direct row
    triangle_candidates_kernel<<<dim3(ceil_div(tcap, kCandThreads), B), kCandThreads, 0, s>>>(
        static_cast<const TriRec*>(trirec), ntri, table, cand, ncand, h * w, C, Cs, tcap);
  FillParams p{C, Cs, h, w, H, W, 0, tcap, 1, mask_u8 ? 1 : 0};
  dim3 grid(ceil_div(W, 4 * kMaskWL * 2), ceil_div(H, (32 / kMaskWL) * (kMaskThreads / 32 / 2)), B);
  inverse_mask_kernel<<<grid, kMaskThreads, 0, s>>>(loc, static_cast<const TriRec*>(trirec), table, cand, ncand, nodearg,
                                                    mask, p);
  return check_launch("fovea_inverse_mask");
}

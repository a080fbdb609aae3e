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
//     (typically 2-8 of 51 on i.i.d. predictions, 2 on the reference's C1 decoder) together with their three vertex
//     values; inverse_mask_kernel evaluates only those -- with exactly the arithmetic of fovea_inverse_fill.
//
// Exactness (the mask equals the fused argmax of fovea_inverse_fill bit for bit, ties included):
//   - torch.argmax returns the FIRST maximum.  A channel may be dropped because of a LOWER-index channel that is >= at
//     all three vertices (that one wins every tie anyway); because of a HIGHER-index channel only if that one is
//     greater at all three vertices by a margin (2^-20 of the larger magnitude) that exceeds the worst-case rounding of
//     both evaluations (3 roundings of relative 2^-24 each, weights summing to 1 +- 2^-23), so it is STRICTLY greater
//     at every pixel;
//   - the third weight float(1 - c0 - c1) can round to a tiny negative number on an edge; monotonicity then fails in
//     principle, so such a pixel (and any triangle whose survivors overflow the record) evaluates all C channels;
//   - survivors are kept in increasing channel order and compared with `>`, like the full loop.
#include <math_constants.h>

#include "common.cuh"
#include "fill.cuh"

namespace fovea {

constexpr int kCandMax = 16;          // survivors kept per triangle (one 256-byte record); more -> full evaluation
constexpr int kCandThreads = 128;
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

// One thread per triangle.  The survivor list lives in shared memory, [slot][field][thread] (conflict-free).
__global__ void __launch_bounds__(kCandThreads)
triangle_candidates_kernel(const TriRec* __restrict__ trirec, const int32_t* __restrict__ ntri,
                           const float* __restrict__ table, float4* __restrict__ cand, uint8_t* __restrict__ ncand,
                           int hw, int C, int Cs, int tcap) {
  __shared__ float sv[kCandMax][4][kCandThreads];
  const int b = blockIdx.y;
  const int t = blockIdx.x * kCandThreads + threadIdx.x;
  if (t >= ntri[b]) return;
  const int tid = threadIdx.x;
  const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(trirec + static_cast<size_t>(b) * tcap + t) + 3);
  const int r0 = static_cast<int>(q3.x & 0xFFFFu), r1 = static_cast<int>(q3.x >> 16), r2 = static_cast<int>(q3.y);
  uint8_t* nc = ncand + static_cast<size_t>(b) * tcap + t;
  if (r0 >= hw || r1 >= hw || r2 >= hw) { *nc = static_cast<uint8_t>(kCandNaN); return; }
  const float* tb = table + static_cast<size_t>(b) * (hw + 2) * Cs;
  const float4* p0 = reinterpret_cast<const float4*>(tb + static_cast<size_t>(r0) * Cs);
  const float4* p1 = reinterpret_cast<const float4*>(tb + static_cast<size_t>(r1) * Cs);
  const float4* p2 = reinterpret_cast<const float4*>(tb + static_cast<size_t>(r2) * Cs);
  int n = 0;
  bool overflow = false;
  for (int c4 = 0; c4 < C && !overflow; c4 += 4) {
    const float4 A = __ldg(p0 + (c4 >> 2)), Bv = __ldg(p1 + (c4 >> 2)), Cv = __ldg(p2 + (c4 >> 2));
    const float a4[4] = {A.x, A.y, A.z, A.w}, b4[4] = {Bv.x, Bv.y, Bv.z, Bv.w}, c4v[4] = {Cv.x, Cv.y, Cv.z, Cv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (c4 + e >= C || overflow) continue;
      const float a = a4[e], bb = b4[e], cc = c4v[e];
      const float mag = fmaxf(fmaxf(fabsf(a), fabsf(bb)), fabsf(cc));
      bool dominated = false;
      for (int k = 0; k < n && !dominated; ++k)      // a lower index that is never smaller wins every tie anyway
        dominated = sv[k][0][tid] >= a && sv[k][1][tid] >= bb && sv[k][2][tid] >= cc;
      if (dominated) continue;
      int keep = 0;
      for (int k = 0; k < n; ++k) {                  // survivors the newcomer (a higher index) beats STRICTLY everywhere
        const float ka = sv[k][0][tid], kb = sv[k][1][tid], kc = sv[k][2][tid];
        const float m = 9.5367431640625e-07f * fmaxf(mag, fmaxf(fmaxf(fabsf(ka), fabsf(kb)), fabsf(kc)));
        if ((a - ka > m) && (bb - kb > m) && (cc - kc > m)) continue;
        if (keep != k) {
          sv[keep][0][tid] = ka; sv[keep][1][tid] = kb; sv[keep][2][tid] = kc; sv[keep][3][tid] = sv[k][3][tid];
        }
        ++keep;
      }
      n = keep;
      if (n == kCandMax) { overflow = true; continue; }
      sv[n][0][tid] = a; sv[n][1][tid] = bb; sv[n][2][tid] = cc; sv[n][3][tid] = __int_as_float(c4 + e);
      ++n;
    }
  }
  if (overflow) { *nc = static_cast<uint8_t>(kCandFull); return; }
  float4* out = cand + (static_cast<size_t>(b) * tcap + t) * kCandMax;
  for (int k = 0; k < n; ++k) out[k] = make_float4(sv[k][0][tid], sv[k][1][tid], sv[k][2][tid], sv[k][3][tid]);
  *nc = static_cast<uint8_t>(n);
}

constexpr int kMaskThreads = 256;
constexpr int kMaskWL = 16;  // the fill's mapping: a warp covers 64 px x 2 rows, a thread 4 consecutive pixels

__global__ void __launch_bounds__(kMaskThreads)
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
  const float* tb = table + static_cast<size_t>(b) * (hw + 2) * p.Cs;

  const uint2 l2 = __ldcs(reinterpret_cast<const uint2*>(loc + static_cast<size_t>(b) * plane + pixoff));
  const int lc[4] = {decode_loc(l2.x & 0xFFFFu), decode_loc(l2.x >> 16), decode_loc(l2.y & 0xFFFFu), decode_loc(l2.y >> 16)};
  int label[4];
  int cur = -1, e0 = 0, e1 = 0, d0 = 0, d1 = 0, sn0 = hw, sn1 = hw, sn2 = hw;
  unsigned n = 0;
  double inv_area = 0.0;
  const float4* cl = nullptr;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
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
        sn0 = static_cast<int>(q3.x & 0xFFFFu); sn1 = static_cast<int>(q3.x >> 16); sn2 = static_cast<int>(q3.y);
        inv_area = __hiloint2double(static_cast<int>(q3.w), static_cast<int>(q3.z));
        n = ncand[static_cast<size_t>(b) * p.tcap + cur];
        cl = cand + (static_cast<size_t>(b) * p.tcap + cur) * kCandMax;
      }
      // interp2d.py:58-65 / qhull.pyx:1210-1264: c0, c1 in float64, c2 = 1 - c0 - c1, then cast to float32
      const double c0 = static_cast<double>(e0) * inv_area, c1 = static_cast<double>(e1) * inv_area;
      const float w0 = static_cast<float>(c0), w1 = static_cast<float>(c1), w2 = static_cast<float>(1.0 - c0 - c1);
      int bi = 0;
      if (n == kCandNaN) {
        bi = 0;                              // NaN (or zeroed) in every channel: torch.argmax gives 0
      } else if (n == kCandFull || w2 < 0.f || w0 < 0.f || w1 < 0.f) {
        float best = 0.f;                    // every channel, the arithmetic of fill_tile (inverse.cu)
        const float* ra = tb + static_cast<size_t>(sn0) * p.Cs;
        const float* rb = tb + static_cast<size_t>(sn1) * p.Cs;
        const float* rc = tb + static_cast<size_t>(sn2) * p.Cs;
        for (int c = 0; c < p.C; ++c) {
          const float v = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(ra + c), w0), __fmul_rn(__ldg(rb + c), w1)),
                                    __fmul_rn(__ldg(rc + c), w2));
          if (c == 0 || v > best) { best = v; bi = c; }
        }
      } else {
        float best = 0.f;
        for (unsigned j = 0; j < n; ++j) {
          const float4 s = __ldg(cl + j);
          const float v = __fadd_rn(__fadd_rn(__fmul_rn(s.x, w0), __fmul_rn(s.y, w1)), __fmul_rn(s.z, w2));
          if (j == 0 || v > best) { best = v; bi = __float_as_int(s.w); }
        }
      }
      label[k] = bi;
    }
    e0 += d0;
    e1 += d1;
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

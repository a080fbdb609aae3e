// Stage 3 backward: the transpose of fovea_inverse_fill with respect to the value table.
//
// Reference: the autograd graph of models/models.py:933-940 (MODEL.loss_at_high_res / MODEL.upsample) --
// F.grid_sample(pred, grid_inv) -> NaN mask -> fillMissingValues_tensor -> Interp2D.forward, whose docstring
// (interp2d.py:38-47) promises gradients w.r.t. `values`.  Forward:  out[c,p] = sum_k w_k(p) * table[row_k(p), c];
// here:  grad_table[row, c] += w_k(p) * grad_out[c,p]  over every pixel p and vertex k that reads `row`.
// Pixels the forward leaves NaN (a vertex without a value) or sets to zero (models_instance.py:940) are constants:
// they send no gradient.  The 2x2 box mean that builds the table from `pred` is F.grid_sample at the node
// coordinates, so its transpose is fovea_grid_sample_bwd (host side: fovea/ops.py).
//
// Work decomposition: the forward fill's (one thread = 4 consecutive pixels, a warp = 64 px x 2 rows).  A table row
// is read by every pixel of every triangle around its node -- 10^3..10^5 pixels in the periphery -- so the adds are
// aggregated twice before they reach memory: inside the thread (pixels that share their rows with the thread's first
// pixel) and across the warp by a segmented shuffle reduction over runs of lanes with equal rows (a row of pixels
// crosses a triangle in ONE run, so runs are contiguous in lane order); only run heads issue red.global.add.f32.
#include "common.cuh"
#include "fill.cuh"

namespace fovea {

constexpr int kBwdThreads = 256;
constexpr int kBwdWL = 16;  // lanes of a warp across a row (x 4 pixels); the warp covers 2 rows -- the fill's mapping

__global__ void __launch_bounds__(kBwdThreads)
inverse_fill_bwd_kernel(const uint16_t* __restrict__ loc, const TriRec* __restrict__ trirec,
                        const float* __restrict__ gout, float* __restrict__ gtable, FillParams p) {
  const int b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int WX = 2, kWarpW = 4 * kBwdWL, kWarpH = 32 / kBwdWL, kTileH = kWarpH * (kBwdThreads / 32 / WX);
  const int x0 = blockIdx.x * (kWarpW * WX) + (warp % WX) * kWarpW + (lane % kBwdWL) * 4;
  const int y = blockIdx.y * kTileH + (warp / WX) * kWarpH + (lane / kBwdWL);
  const bool live = x0 < p.W && y < p.H;  // dead lanes still take part in the shuffles (with empty keys)
  const int hw = p.h * p.w;
  const size_t plane = static_cast<size_t>(p.H) * p.W;
  const unsigned pixoff = live ? static_cast<unsigned>(y) * p.W + x0 : 0u;
  const TriRec* recs = trirec + static_cast<size_t>(b) * p.tcap;

  // ---- rows and weights of my four pixels: the same arithmetic as fill_tile (inverse.cu)
  unsigned nd0[4] = {0, 0, 0, 0}, nd1[4] = {0, 0, 0, 0}, nd2[4] = {0, 0, 0, 0};
  float w0[4] = {0, 0, 0, 0}, w1[4] = {0, 0, 0, 0}, w2[4] = {0, 0, 0, 0};
  unsigned dead = live ? 0u : 0xFu;  // bit k: pixel k sends no gradient
  if (live) {
    const uint2 l2 = __ldg(reinterpret_cast<const uint2*>(loc + static_cast<size_t>(b) * plane + pixoff));
    const int lc[4] = {decode_loc(l2.x & 0xFFFFu), decode_loc(l2.x >> 16), decode_loc(l2.y & 0xFFFFu), decode_loc(l2.y >> 16)};
    int cur = -1, sn0 = hw, sn1 = hw, sn2 = hw, e0 = 0, e1 = 0, d0 = 0, d1 = 0;
    double inv_area = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int n0, n1, n2;
      float a0 = 1.f, a1 = 0.f, a2 = 0.f;
      if (lc[k] < 0) {
        n0 = n1 = n2 = -(lc[k] + 1);
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
        }
        const double c0 = static_cast<double>(e0) * inv_area, c1 = static_cast<double>(e1) * inv_area;
        a0 = static_cast<float>(c0);
        a1 = static_cast<float>(c1);
        a2 = fmaxf(static_cast<float>(1.0 - c0 - c1), 0.f);
        n0 = sn0; n1 = sn1; n2 = sn2;
      }
      e0 += d0;
      e1 += d1;
      if (n0 >= hw || n1 >= hw || n2 >= hw) dead |= 1u << k;  // NaN (or zeroed) output: a constant
      nd0[k] = static_cast<unsigned>(n0); nd1[k] = static_cast<unsigned>(n1); nd2[k] = static_cast<unsigned>(n2);
      w0[k] = a0; w1[k] = a1; w2[k] = a2;
    }
  }

  // ---- group 0: my first live pixel and the pixels that share its three rows; the others add on their own
  const int lead = dead == 0xFu ? 0 : __ffs(~dead & 0xFu) - 1;
  unsigned g0 = 0;  // bit k: pixel k is in group 0
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (!((dead >> k) & 1u) && nd0[k] == nd0[lead] && nd1[k] == nd1[lead] && nd2[k] == nd2[lead]) g0 |= 1u << k;
  const unsigned solo = ~dead & ~g0 & 0xFu;
  // run structure of the warp over group-0 keys (45 bits: three 15-bit rows); a thread with no live pixel has key 0
  // with bit 63 set and never merges (each such lane is its own run and adds nothing)
  const unsigned long long key = g0 ? (static_cast<unsigned long long>(nd0[lead]) | (static_cast<unsigned long long>(nd1[lead]) << 15) |
                                       (static_cast<unsigned long long>(nd2[lead]) << 30))
                                    : (0x8000000000000000ull | lane);
  const unsigned long long up = __shfl_up_sync(0xffffffffu, key, 1);
  const bool head = lane == 0 || up != key;
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  const int seg = __popc(heads & (0xffffffffu >> (31 - lane)));  // run number of my lane (1-based)
  unsigned okmask = 0;  // bit i: lane + 2^i belongs to my run
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int other = __shfl_down_sync(0xffffffffu, seg, 1 << i);
    if (lane + (1 << i) < 32 && other == seg) okmask |= 1u << i;
  }
  const bool single = nd0[lead] == nd1[lead] && nd1[lead] == nd2[lead];  // a pixel that received a node: one row, weight 1

  // group-0 weights with the other pixels' zeroed: the channel loop below is then free of per-pixel predicates
  float u0[4], u1[4], u2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool in = (g0 >> k) & 1u;
    u0[k] = in ? w0[k] : 0.f; u1[k] = in ? w1[k] : 0.f; u2[k] = in ? w2[k] : 0.f;
  }
  const bool emit = head && g0;
  // byte offsets of the three rows inside the frame's table (rows < 2^15, Cs <= 2^16: fits 32 bits)
  const unsigned o0 = nd0[lead] * static_cast<unsigned>(p.Cs) * 4u, o1 = nd1[lead] * static_cast<unsigned>(p.Cs) * 4u,
                 o2 = nd2[lead] * static_cast<unsigned>(p.Cs) * 4u;
  const float* gp = gout + static_cast<size_t>(b) * p.C * plane + pixoff;
  char* tb = reinterpret_cast<char*>(gtable + static_cast<size_t>(b) * (hw + 2) * p.Cs);
  for (int c = 0; c < p.C; ++c, gp += plane, tb += 4) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) v = __ldcs(reinterpret_cast<const float4*>(gp));
    float s0 = fmaf(u0[3], v.w, fmaf(u0[2], v.z, fmaf(u0[1], v.y, u0[0] * v.x)));
    float s1 = fmaf(u1[3], v.w, fmaf(u1[2], v.z, fmaf(u1[1], v.y, u1[0] * v.x)));
    float s2 = fmaf(u2[3], v.w, fmaf(u2[2], v.z, fmaf(u2[1], v.y, u2[0] * v.x)));
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float t0 = __shfl_down_sync(0xffffffffu, s0, 1 << i);
      const float t1 = __shfl_down_sync(0xffffffffu, s1, 1 << i);
      const float t2 = __shfl_down_sync(0xffffffffu, s2, 1 << i);
      const bool ok = (okmask >> i) & 1u;
      s0 += ok ? t0 : 0.f; s1 += ok ? t1 : 0.f; s2 += ok ? t2 : 0.f;
    }
    if (emit) {
      atomicAdd(reinterpret_cast<float*>(tb + o0), s0);   // (a pixel that received a node: weights (1,0,0), one row)
      if (!single) {
        atomicAdd(reinterpret_cast<float*>(tb + o1), s1);
        atomicAdd(reinterpret_cast<float*>(tb + o2), s2);
      }
    }
  }
  // the pixels of this thread that do not share their rows with its first live pixel (triangle boundaries inside the
  // thread's 4 pixels, the densely filled fovea): their own adds, outside the warp-synchronous loop
  if (solo) {
    const float* gq = gout + static_cast<size_t>(b) * p.C * plane + pixoff;
    float* tq = gtable + static_cast<size_t>(b) * (hw + 2) * p.Cs;
    for (int c = 0; c < p.C; ++c, gq += plane, ++tq) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(gq));
      const float g[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (!((solo >> k) & 1u)) continue;
        if (nd0[k] == nd1[k] && nd1[k] == nd2[k]) {
          atomicAdd(tq + static_cast<size_t>(nd0[k]) * p.Cs, g[k]);
        } else {
          atomicAdd(tq + static_cast<size_t>(nd0[k]) * p.Cs, w0[k] * g[k]);
          atomicAdd(tq + static_cast<size_t>(nd1[k]) * p.Cs, w1[k] * g[k]);
          atomicAdd(tq + static_cast<size_t>(nd2[k]) * p.Cs, w2[k] * g[k]);
        }
      }
    }
  }
}

}  // namespace fovea

using namespace fovea;

extern "C" int fovea_inverse_fill_bwd(const uint16_t* loc, const void* trirec, const float* grad_scores, int B, int C,
                                      int Cs, int h, int w, int H, int W, int tcap, float* grad_table,
                                      fovea_stream_t stream) {
  FOVEA_REQUIRE(loc && trirec && grad_scores && grad_table, "fovea_inverse_fill_bwd: null pointer");
  FOVEA_REQUIRE(B > 0 && C > 0 && Cs >= C && h > 0 && w > 0 && H > 1 && W > 1, "fovea_inverse_fill_bwd: bad sizes");
  FOVEA_REQUIRE(W % 4 == 0, "fovea_inverse_fill_bwd: W=%d must be a multiple of 4 (128-bit loads)", W);
  FOVEA_REQUIRE(H <= 16384 && W <= 16384 && B <= 65535, "fovea_inverse_fill_bwd: canvas or batch too large");
  FOVEA_REQUIRE(static_cast<long long>(h) * w + 2 <= 32768, "fovea_inverse_fill_bwd: value table rows must fit 15 bits");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FOVEA_CUDA(cudaMemsetAsync(grad_table, 0, sizeof(float) * static_cast<size_t>(B) * (static_cast<size_t>(h) * w + 2) * Cs, s));
  FillParams p{C, Cs, h, w, H, W, 0, tcap, 0, 0};
  dim3 grid(ceil_div(W, 4 * kBwdWL * 2), ceil_div(H, (32 / kBwdWL) * (kBwdThreads / 32 / 2)), B);
  inverse_fill_bwd_kernel<<<grid, kBwdThreads, 0, s>>>(loc, static_cast<const TriRec*>(trirec), grad_scores, grad_table, p);
  return check_launch("fovea_inverse_fill_bwd");
}

// Experiment (FOVEA_FILL_SMEM=1): fovea_inverse_fill's scores kernel with the table rows of a CTA's tile staged in SHARED
// memory.
//
// Why: inverse_fill_kernel is at 0.92 of the HBM peak alone, but it is SM-heavy -- ncu: L1 data pipe 75 % busy, because every
// 128-bit table-row load of a warp (64 px x 2 rows, ~10 triangles) costs 10-16 L1 wavefronts (one per distinct row).  Under
// the pipelined schedule the next batch's plan kernels take SMs away and the fill, having no slack on the SM side, stretches
// from 2.28 to 3.2 ms.  A 128 x 8 pixel tile touches only ~25 distinct table rows (5 KB): staged once per CTA through a small
// hash in shared memory, the per-channel loads become conflict-free LDS.128 (4 wavefronts per warp).
//
// Same arithmetic as fill_tile (inverse.cu) operation for operation: the outputs are bit-identical.  Tiles that touch more
// than kSmemRows distinct rows (the densely filled fovea, where every pixel carries its own node) fall back to global loads.
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "fill.cuh"
#include "tma.cuh"

namespace fovea {

constexpr int kSfThreads = 256;
constexpr int kSfHash = 512;     // hash slots (power of two) for the distinct rows of one tile
constexpr int kSmemRows = 128;   // rows staged per CTA (Cs * 4 bytes each)

// kTma: the scores leave through shared memory and cp.async.bulk.tensor tile stores (one 128 x 8 tile per channel plane,
// issued by one thread, two 4-plane stages in flight) instead of 51 STG.128 per thread -- the stores are half of the L1 data
// pipe's load in the default kernel (ncu: 52 % of its wavefronts remain when the table rows come from shared memory).
constexpr int kSfStageFloats = 4 * 8 * 128;   // 4 channel planes of one 128 x 8 tile

template <bool kTma>
__global__ void __launch_bounds__(kSfThreads, kTma ? 3 : 4)
inverse_fill_smem_kernel(const __grid_constant__ CUtensorMap tmap, const uint16_t* __restrict__ loc,
                         const TriRec* __restrict__ trirec, const float* __restrict__ table, float* __restrict__ scores,
                         FillParams p) {
  extern __shared__ __align__(128) float smem_dyn[];   // [kTma: 2 stages x 4 planes x 8 x 128] [kSmemRows][Cs]
  float* srows = smem_dyn + (kTma ? 2 * kSfStageFloats : 0);
  __shared__ int keys[kSfHash];                    // table row stored in a hash slot, -1 = empty
  __shared__ unsigned short ids[kSfHash];          // dense index of the slot's row in srows
  __shared__ int list[kSmemRows];                  // dense index -> table row
  __shared__ int wsum[9];
  const int b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int WL = 16, WX = 2, kWarpW = 4 * WL, kWarpH = 32 / WL;
  const int x0 = blockIdx.x * (kWarpW * WX) + (warp % WX) * kWarpW + (lane % WL) * 4;
  const int y = blockIdx.y * (kWarpH * (kSfThreads / 32 / WX)) + (warp / WX) * kWarpH + (lane / WL);
  const bool live = x0 < p.W && y < p.H;
  const int hw = p.h * p.w;
  const size_t plane = static_cast<size_t>(p.H) * p.W;
  const unsigned pixoff = live ? static_cast<unsigned>(y) * p.W + x0 : 0u;
  const TriRec* recs = trirec + static_cast<size_t>(b) * p.tcap;
  for (int i = tid; i < kSfHash; i += kSfThreads) keys[i] = -1;
  __syncthreads();

  // ---- rows and weights of my four pixels: fill_tile's arithmetic
  unsigned nd0[4] = {0, 0, 0, 0}, nd1[4] = {0, 0, 0, 0}, nd2[4] = {0, 0, 0, 0};
  float w0[4] = {1.f, 1.f, 1.f, 1.f}, w1[4] = {0.f, 0.f, 0.f, 0.f}, w2[4] = {0.f, 0.f, 0.f, 0.f};
  unsigned reload[4] = {live ? 1u : 0u, 0u, 0u, 0u};   // (a thread outside the canvas loads nothing)
  if (live) {
    const uint2 l2 = __ldcs(reinterpret_cast<const uint2*>(loc + static_cast<size_t>(b) * plane + pixoff));
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
      if (n0 == hw || n1 == hw || n2 == hw) {  // a NaN vertex poisons every channel
        n0 = n1 = n2 = p.zero_residual ? hw + 1 : hw;
        a0 = 1.f; a1 = 0.f; a2 = 0.f;
      }
      nd0[k] = static_cast<unsigned>(n0); nd1[k] = static_cast<unsigned>(n1); nd2[k] = static_cast<unsigned>(n2);
      w0[k] = a0; w1[k] = a1; w2[k] = a2;
      if (k > 0) reload[k] = (nd0[k] != nd0[k - 1]) | (nd1[k] != nd1[k - 1]) | (nd2[k] != nd2[k - 1]);
    }
  }

  // ---- is this a dense tile (the filled fovea: every pixel its own node -> more distinct rows than the hash holds)?
  // Count the runs of pixels with equal rows (a thread's first pixel starts a run if it differs from its left neighbour's
  // last pixel or the lane starts a row of the warp's footprint): distinct rows <= 3 * runs.
  {
    const unsigned p0 = __shfl_up_sync(0xffffffffu, nd0[3], 1), p1 = __shfl_up_sync(0xffffffffu, nd1[3], 1),
                   p2 = __shfl_up_sync(0xffffffffu, nd2[3], 1);
    const bool fresh0 = (lane % WL == 0) | (p0 != nd0[0]) | (p1 != nd1[0]) | (p2 != nd2[0]);
    int runs = live ? static_cast<int>(fresh0) + static_cast<int>(reload[1] + reload[2] + reload[3]) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) runs += __shfl_xor_sync(0xffffffffu, runs, o);
    if (lane == 0) wsum[warp] = runs;
  }
  __syncthreads();
  int nruns = 0;
#pragma unroll
  for (int i = 0; i < kSfThreads / 32; ++i) nruns += wsum[i];
  const bool dense = 3 * nruns > kSfHash - 32;   // block-uniform; (also keeps the open-addressing probes short)
  __syncthreads();                                // wsum is reused below

  // ---- the tile's distinct rows: insert into the hash (the pixel's nd* become hash slots)
  auto insert = [&](unsigned row) -> unsigned {
    unsigned h = (row * 2654435761u) >> 23;   // 9 bits
    for (;;) {
      const int prev = atomicCAS(&keys[h], -1, static_cast<int>(row));
      if (prev == -1 || prev == static_cast<int>(row)) return h;
      h = (h + 1) & (kSfHash - 1);
    }
  };
  if (live && !dense) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (reload[k]) {
        const unsigned r0 = nd0[k], r1 = nd1[k], r2 = nd2[k];
        nd0[k] = insert(r0);
        nd1[k] = r1 == r0 ? nd0[k] : insert(r1);
        nd2[k] = r2 == r0 ? nd0[k] : (r2 == r1 ? nd1[k] : insert(r2));
      } else {
        nd0[k] = nd0[k - (k > 0)]; nd1[k] = nd1[k - (k > 0)]; nd2[k] = nd2[k - (k > 0)];
      }
    }
  }
  __syncthreads();
  // ---- dense numbering of the occupied slots (two slots per thread), the row list, the verdict
  const int o0 = keys[2 * tid] >= 0, o1 = keys[2 * tid + 1] >= 0;
  int incl = o0 + o1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int i = 0; i < kSfThreads / 32; ++i) { const int v = wsum[i]; wsum[i] = acc; acc += v; }
    wsum[8] = acc;
  }
  __syncthreads();
  const int count = wsum[8];
  const bool staged = !dense && count <= kSmemRows;   // block-uniform
  const unsigned row_bytes = static_cast<unsigned>(p.Cs) * 4u;
  if (staged) {
    int id = wsum[warp] + incl - (o0 + o1);
    if (o0) { ids[2 * tid] = static_cast<unsigned short>(id); list[id] = keys[2 * tid]; ++id; }
    if (o1) { ids[2 * tid + 1] = static_cast<unsigned short>(id); list[id] = keys[2 * tid + 1]; }
    __syncthreads();
    // stage: one warp copies one row at a time (Cs / 4 lanes, 16 bytes each: coalesced)
    const float4* t4 = reinterpret_cast<const float4*>(table + static_cast<size_t>(b) * (hw + 2) * p.Cs);
    float4* s4 = reinterpret_cast<float4*>(srows);
    const int q = p.Cs >> 2;
    for (int id2 = warp; id2 < count; id2 += kSfThreads / 32)
      if (lane < q) s4[id2 * q + lane] = __ldg(t4 + static_cast<size_t>(list[id2]) * q + lane);
    if (q > 32)
      for (int id2 = warp; id2 < count; id2 += kSfThreads / 32)
        for (int j = 32 + lane; j < q; j += 32) s4[id2 * q + j] = __ldg(t4 + static_cast<size_t>(list[id2]) * q + j);
    __syncthreads();
  }
  if (!kTma && !live) return;   // (the TMA variant keeps every thread for the per-group barriers)
  // byte offsets of my pixels' rows: in shared memory (staged) or in the frame's table (fallback)
  unsigned off0[4], off1[4], off2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // (staged: dense index of the hash slot; hash used but too many rows: the slot's row; dense tile: nd* still are rows)
    off0[k] = !live ? 0u : (staged ? ids[nd0[k]] : (dense ? nd0[k] : static_cast<unsigned>(keys[nd0[k]]))) * row_bytes;
    off1[k] = !live ? 0u : (staged ? ids[nd1[k]] : (dense ? nd1[k] : static_cast<unsigned>(keys[nd1[k]]))) * row_bytes;
    off2[k] = !live ? 0u : (staged ? ids[nd2[k]] : (dense ? nd2[k] : static_cast<unsigned>(keys[nd2[k]]))) * row_bytes;
  }
  unsigned long long obase = reinterpret_cast<unsigned long long>(scores + static_cast<size_t>(b) * p.C * plane + pixoff);
  const unsigned long long ostep = static_cast<unsigned long long>(plane) * 4ull;
  // TMA variant: my 4 pixels' place inside the 128 x 8 stage tile, and the emit step shared by both row sources
  const int tx = (warp % WX) * kWarpW + (lane % WL) * 4, ty = (warp / WX) * kWarpH + (lane / WL);
  int stage = 0;
  auto emit = [&](float (&v)[4][4], int c) {
    if (!kTma) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (c + e < p.C) {
          __stcs(reinterpret_cast<float4*>(obase), make_float4(v[0][e], v[1][e], v[2][e], v[3][e]));
          obase += ostep;
        }
      }
    } else {
      float* buf = smem_dyn + stage * kSfStageFloats;
      if (tid == 0) tma_wait_read<1>();   // the stores that read THIS stage (committed two groups ago) are done reading
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 4; ++e)
        *reinterpret_cast<float4*>(buf + e * (8 * 128) + ty * 128 + tx) = make_float4(v[0][e], v[1][e], v[2][e], v[3][e]);
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c + e < p.C) tma_store_tile(&tmap, buf + e * (8 * 128), blockIdx.x * 128, blockIdx.y * 8, b * p.C + c + e);
        tma_commit();
      }
      stage ^= 1;
    }
  };
  if (staged) {
    unsigned sbase = static_cast<unsigned>(__cvta_generic_to_shared(srows));
    for (int c = 0; c < p.Cs; c += 4, sbase += 16u) {
      float v[4][4];
      float ra[4] = {0.f, 0.f, 0.f, 0.f}, rb[4] = {0.f, 0.f, 0.f, 0.f}, rc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %15, 0;\n\t"
            "@q ld.shared.v4.f32 {%0,%1,%2,%3}, [%12];\n\t"
            "@q ld.shared.v4.f32 {%4,%5,%6,%7}, [%13];\n\t"
            "@q ld.shared.v4.f32 {%8,%9,%10,%11}, [%14];\n\t}"
            : "+f"(ra[0]), "+f"(ra[1]), "+f"(ra[2]), "+f"(ra[3]), "+f"(rb[0]), "+f"(rb[1]), "+f"(rb[2]), "+f"(rb[3]),
              "+f"(rc[0]), "+f"(rc[1]), "+f"(rc[2]), "+f"(rc[3])
            : "r"(sbase + off0[k]), "r"(sbase + off1[k]), "r"(sbase + off2[k]), "r"(reload[k]));
#pragma unroll
        for (int e = 0; e < 4; ++e)
          v[k][e] = __fadd_rn(__fadd_rn(__fmul_rn(ra[e], w0[k]), __fmul_rn(rb[e], w1[k])), __fmul_rn(rc[e], w2[k]));
      }
      emit(v, c);
    }
  } else {
    unsigned long long tbase = reinterpret_cast<unsigned long long>(table + static_cast<size_t>(b) * (hw + 2) * p.Cs);
    for (int c = 0; c < p.Cs; c += 4, tbase += 16ull) {
      float v[4][4];
      float ra[4] = {0.f, 0.f, 0.f, 0.f}, rb[4] = {0.f, 0.f, 0.f, 0.f}, rc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %15, 0;\n\t"
            "@q ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%12];\n\t"
            "@q ld.global.nc.v4.f32 {%4,%5,%6,%7}, [%13];\n\t"
            "@q ld.global.nc.v4.f32 {%8,%9,%10,%11}, [%14];\n\t}"
            : "+f"(ra[0]), "+f"(ra[1]), "+f"(ra[2]), "+f"(ra[3]), "+f"(rb[0]), "+f"(rb[1]), "+f"(rb[2]), "+f"(rb[3]),
              "+f"(rc[0]), "+f"(rc[1]), "+f"(rc[2]), "+f"(rc[3])
            : "l"(tbase + off0[k]), "l"(tbase + off1[k]), "l"(tbase + off2[k]), "r"(reload[k]));
#pragma unroll
        for (int e = 0; e < 4; ++e)
          v[k][e] = __fadd_rn(__fadd_rn(__fmul_rn(ra[e], w0[k]), __fmul_rn(rb[e], w1[k])), __fmul_rn(rc[e], w2[k]));
      }
      emit(v, c);
    }
  }
  if (kTma && tid == 0) tma_wait_read<0>();   // shared memory must outlive the last tile stores
}

}  // namespace fovea

using namespace fovea;

// scores-only launch of the shared-memory variant (mode 1) or the shared-memory + TMA-store variant (mode 2); returns
// FOVEA_OK, or -1 if the shape is not eligible (the caller then uses the default kernel)
int fovea_launch_fill_smem(const uint16_t* loc, const void* trirec, const float* table, int B, int C, int Cs, int h, int w,
                           int H, int W, int tcap, int zero_residual, float* scores, int mode, cudaStream_t s) {
  if (W < 128 || Cs % 4 != 0 || Cs > 256) return -1;
  FillParams p{C, Cs, h, w, H, W, 0, tcap, zero_residual, 0};
  dim3 grid(ceil_div(W, 128), ceil_div(H, 8), B);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (mode == 2) {
    if (W % 4 != 0 || make_plane_store_map(&map, scores, static_cast<long long>(B) * C, H, W, 128, 8)) return -1;
    const int smem = kSmemRows * Cs * 4 + 2 * kSfStageFloats * 4;
    if (cudaFuncSetAttribute(inverse_fill_smem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    inverse_fill_smem_kernel<true><<<grid, kSfThreads, smem, s>>>(map, loc, static_cast<const TriRec*>(trirec), table, scores, p);
  } else {
    const int smem = kSmemRows * Cs * 4;
    if (cudaFuncSetAttribute(inverse_fill_smem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    inverse_fill_smem_kernel<false><<<grid, kSfThreads, smem, s>>>(map, loc, static_cast<const TriRec*>(trirec), table, scores, p);
  }
  return check_launch("fovea_inverse_fill (smem rows)");
}

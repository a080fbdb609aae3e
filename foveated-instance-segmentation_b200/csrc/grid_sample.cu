// Stage 2: F.grid_sample(input, grid) forward + backward with the reference's defaults
// (mode='bilinear', padding_mode='zeros', align_corners=False): models/models.py:865, 880, 909, 937.
//
// The arithmetic follows aten's grid_sampler_2d bit for bit (verified against torch CPU in
// tests/test_oracle_golden.py): ix = fma(x+1, W/2, -0.5); corner weights nw,ne,sw,se from the distances to
// the opposite corner; accumulation acc = v_nw*nw, then fma in the order ne, sw, se; out-of-bounds taps add 0.
//
// One thread owns one output pixel (b, oy, ox) and loops over the C channels: the 4 tap addresses and weights
// are computed once, lanes of a warp cover 32 consecutive ox so output stores are coalesced, and in the fovea
// (tap spacing < 1 px) neighbouring lanes hit the same 32 B sectors.  The source image is only touched at
// 4*h*w points per channel, so the kernel is sector-gather bound, not streaming.
#include <stdlib.h>

#include "common.cuh"
#include "taps.cuh"
#include "tma.cuh"

namespace fovea {

__global__ void __launch_bounds__(256)
grid_sample_fwd_kernel(const float* __restrict__ in, const float2* __restrict__ grid, float* __restrict__ out, int B,
                       int C, int H, int W, int hw_out) {
  const long long total = static_cast<long long>(B) * hw_out;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(idx / hw_out);
    const int pix = static_cast<int>(idx - static_cast<long long>(b) * hw_out);
    const float2 g = grid[idx];
    const Taps t = make_taps(g.x, g.y, H, W);
    const size_t plane = static_cast<size_t>(H) * W;
    const float* src = in + static_cast<size_t>(b) * C * plane + static_cast<long long>(t.y0) * W + t.x0;
    float* dst = out + static_cast<size_t>(b) * C * hw_out + pix;
#pragma unroll 3
    for (int c = 0; c < C; ++c) {
      const float* s = src + c * plane;
      const float v_nw = t.ok_nw ? __ldg(s) : 0.f;
      const float v_ne = t.ok_ne ? __ldg(s + 1) : 0.f;
      const float v_sw = t.ok_sw ? __ldg(s + W) : 0.f;
      const float v_se = t.ok_se ? __ldg(s + W + 1) : 0.f;
      float acc = v_nw * t.nw;
      acc = fmaf(v_ne, t.ne, acc);
      acc = fmaf(v_sw, t.sw, acc);
      acc = fmaf(v_se, t.se, acc);
      dst[static_cast<size_t>(c) * hw_out] = acc;
    }
  }
}

// Experiment (FOVEA_GS_TMA=1; north_star: "TMA-staged source tiles around each foveal region").  One CTA = a 16 x 16 tile
// of output pixels of one frame.  The CTA reduces the bounding box of its 4 x 256 taps; where the sampling grid is dense
// (the fovea: neighbouring outputs less than a few source pixels apart) the box fits 64 x 32 source pixels and ONE
// cp.async.bulk.tensor per channel stages it in shared memory -- out-of-image elements arrive as zeros, which is
// padding_mode='zeros' -- and the taps are read from there; everywhere else (the periphery: one output every 10-200 source
// pixels) the CTA gathers straight from global memory as the default kernel does.  Same arithmetic, same order: the two
// kernels are bit-identical.  Measured result in DESIGN.md section 4.
constexpr int kGsTile = 16, kGsBoxW = 64, kGsBoxH = 32, kGsMaxC = 4;

__global__ void __launch_bounds__(kGsTile * kGsTile)
grid_sample_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ in,
                           const float2* __restrict__ grid, float* __restrict__ out, int C, int H, int W, int h, int w,
                           int* __restrict__ staged_count) {
  __shared__ __align__(128) float box[kGsMaxC][kGsBoxH][kGsBoxW];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ int red[4][8];
  const int b = blockIdx.z;
  const int ox = blockIdx.x * kGsTile + (threadIdx.x & (kGsTile - 1)), oy = blockIdx.y * kGsTile + (threadIdx.x / kGsTile);
  const bool live = ox < w && oy < h;
  const int hw_out = h * w, pix = oy * w + ox;
  Taps t;
  int xlo = 1 << 30, xhi = -(1 << 30), ylo = 1 << 30, yhi = -(1 << 30);
  if (live) {
    const float2 g = grid[static_cast<size_t>(b) * hw_out + pix];
    t = make_taps(g.x, g.y, H, W);
    if (t.ok_nw || t.ok_ne || t.ok_sw || t.ok_se) { xlo = t.x0; xhi = t.x0 + 1; ylo = t.y0; yhi = t.y0 + 1; }
  }
  // block-wide bounding box of the taps
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xlo = min(xlo, __shfl_xor_sync(0xffffffffu, xlo, o)); xhi = max(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
    ylo = min(ylo, __shfl_xor_sync(0xffffffffu, ylo, o)); yhi = max(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = xlo; red[1][warp] = xhi; red[2][warp] = ylo; red[3][warp] = yhi; }
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    xlo = min(xlo, red[0][i]); xhi = max(xhi, red[1][i]); ylo = min(ylo, red[2][i]); yhi = max(yhi, red[3][i]);
  }
  xlo &= ~3;   // the box must start on a 16-byte boundary of the row (TMA: coordinate * element size % 16 == 0; -1 -> -4)
  const bool staged = xhi >= xlo && xhi - xlo < kGsBoxW && yhi - ylo < kGsBoxH;   // block-uniform
  if (staged) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, static_cast<unsigned>(C) * kGsBoxH * kGsBoxW * 4u);
      for (int c = 0; c < C; ++c) tma_load_tile(&tmap, &box[c][0][0], &bar, xlo, ylo, b * C + c);
      if (staged_count) atomicAdd(staged_count, 1);
    }
    mbar_wait(&bar, 0);
  }
  if (!live) return;
  const size_t plane = static_cast<size_t>(H) * W;
  const float* src = in + static_cast<size_t>(b) * C * plane + static_cast<long long>(t.y0) * W + t.x0;
  float* dst = out + static_cast<size_t>(b) * C * hw_out + pix;
  const int bx = t.x0 - xlo, by = t.y0 - ylo;
  for (int c = 0; c < C; ++c) {
    float v_nw, v_ne, v_sw, v_se;
    if (staged) {
      v_nw = t.ok_nw ? box[c][by][bx] : 0.f;
      v_ne = t.ok_ne ? box[c][by][bx + 1] : 0.f;
      v_sw = t.ok_sw ? box[c][by + 1][bx] : 0.f;
      v_se = t.ok_se ? box[c][by + 1][bx + 1] : 0.f;
    } else {
      const float* s = src + c * plane;
      v_nw = t.ok_nw ? __ldg(s) : 0.f;
      v_ne = t.ok_ne ? __ldg(s + 1) : 0.f;
      v_sw = t.ok_sw ? __ldg(s + W) : 0.f;
      v_se = t.ok_se ? __ldg(s + W + 1) : 0.f;
    }
    float acc = v_nw * t.nw;
    acc = fmaf(v_ne, t.ne, acc);
    acc = fmaf(v_sw, t.sw, acc);
    acc = fmaf(v_se, t.se, acc);
    dst[static_cast<size_t>(c) * hw_out] = acc;
  }
}

// The same gather from a uint8 image: the loader's ToTensor() (uint8 -> fp32 / 255) is
// folded into the tap loads, so the full-resolution frame crosses PCIe and sits in HBM at 1 byte per sample.  The
// division is the fp32 division ToTensor performs, so the result is bit-identical to sampling the converted image.
__global__ void __launch_bounds__(256)
grid_sample_fwd_u8_kernel(const unsigned char* __restrict__ in, const float2* __restrict__ grid, float* __restrict__ out,
                          int B, int C, int H, int W, int hw_out, float divisor) {
  const long long total = static_cast<long long>(B) * hw_out;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(idx / hw_out);
    const int pix = static_cast<int>(idx - static_cast<long long>(b) * hw_out);
    const float2 g = grid[idx];
    const Taps t = make_taps(g.x, g.y, H, W);
    const size_t plane = static_cast<size_t>(H) * W;
    const unsigned char* src = in + static_cast<size_t>(b) * C * plane + static_cast<long long>(t.y0) * W + t.x0;
    float* dst = out + static_cast<size_t>(b) * C * hw_out + pix;
#pragma unroll 3
    for (int c = 0; c < C; ++c) {
      const unsigned char* s = src + c * plane;
      const float v_nw = t.ok_nw ? __fdiv_rn(static_cast<float>(__ldg(s)), divisor) : 0.f;
      const float v_ne = t.ok_ne ? __fdiv_rn(static_cast<float>(__ldg(s + 1)), divisor) : 0.f;
      const float v_sw = t.ok_sw ? __fdiv_rn(static_cast<float>(__ldg(s + W)), divisor) : 0.f;
      const float v_se = t.ok_se ? __fdiv_rn(static_cast<float>(__ldg(s + W + 1)), divisor) : 0.f;
      float acc = v_nw * t.nw;
      acc = fmaf(v_ne, t.ne, acc);
      acc = fmaf(v_sw, t.sw, acc);
      acc = fmaf(v_se, t.se, acc);
      dst[static_cast<size_t>(c) * hw_out] = acc;
    }
  }
}

// Warp-aggregated atomic add: lanes of the warp that target the same address elect a leader that adds the
// group's sum with ONE red.global.add.f32 (the stock kernel issues one atomic per lane and tap).
__device__ __forceinline__ void warp_aggregated_add(float* addr, float val, bool active) {
  const unsigned live = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const unsigned long long key = reinterpret_cast<unsigned long long>(addr);
  const unsigned peers = __match_any_sync(live, key);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(peers) - 1;
  float sum = 0.f;
  unsigned rest = peers;
  while (rest) {
    const int src = __ffs(rest) - 1;
    sum += __shfl_sync(peers, val, src);
    rest &= rest - 1;
  }
  if (lane == leader) atomicAdd(addr, sum);
}

template <bool kGradIn, bool kGradGrid>
__global__ void __launch_bounds__(256)
grid_sample_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ in, const float2* __restrict__ grid,
                       float* __restrict__ gin, float2* __restrict__ ggrid, int B, int C, int H, int W, int hw_out) {
  const long long total = static_cast<long long>(B) * hw_out;
  // the loop bound is rounded up to whole warps so that every lane reaches the warp-collective adds
  const long long total_w = (total + 31) / 32 * 32;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total_w;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const bool live = idx < total;
    const long long sidx = live ? idx : total - 1;
    const int b = static_cast<int>(sidx / hw_out);
    const int pix = static_cast<int>(sidx - static_cast<long long>(b) * hw_out);
    const float2 g = grid[sidx];
    const Taps t = make_taps(g.x, g.y, H, W);
    const size_t plane = static_cast<size_t>(H) * W;
    const long long off = static_cast<long long>(t.y0) * W + t.x0;
    const float* src = in + static_cast<size_t>(b) * C * plane + off;
    float* gsrc = kGradIn ? gin + static_cast<size_t>(b) * C * plane + off : nullptr;
    const float* go = gout + static_cast<size_t>(b) * C * hw_out + pix;
    // distances used by aten's backward: d/d ix of the weights
    const float fx = floorf(t.ix), fy = floorf(t.iy);
    const float dx_e = (fx + 1.f) - t.ix, dx_w = t.ix - fx;  // (ix_se - ix), (ix - ix_nw)
    const float dy_s = (fy + 1.f) - t.iy, dy_n = t.iy - fy;  // (iy_se - iy), (iy - iy_nw)
    float gix = 0.f, giy = 0.f;
    for (int c = 0; c < C; ++c) {
      const float go_c = live ? go[static_cast<size_t>(c) * hw_out] : 0.f;
      if (kGradGrid) {
        const float* s = src + c * plane;
        const float v_nw = t.ok_nw ? __ldg(s) : 0.f;
        const float v_ne = t.ok_ne ? __ldg(s + 1) : 0.f;
        const float v_sw = t.ok_sw ? __ldg(s + W) : 0.f;
        const float v_se = t.ok_se ? __ldg(s + W + 1) : 0.f;
        gix -= v_nw * dy_s * go_c;
        giy -= v_nw * dx_e * go_c;
        gix += v_ne * dy_s * go_c;
        giy -= v_ne * dx_w * go_c;
        gix -= v_sw * dy_n * go_c;
        giy += v_sw * dx_e * go_c;
        gix += v_se * dy_n * go_c;
        giy += v_se * dx_w * go_c;
      }
      if (kGradIn) {
        float* d = gsrc + c * plane;
        warp_aggregated_add(d, t.nw * go_c, live && t.ok_nw);
        warp_aggregated_add(d + 1, t.ne * go_c, live && t.ok_ne);
        warp_aggregated_add(d + W, t.sw * go_c, live && t.ok_sw);
        warp_aggregated_add(d + W + 1, t.se * go_c, live && t.ok_se);
      }
    }
    if (kGradGrid && live) {
      float2 r;
      r.x = (0.5f * static_cast<float>(W)) * gix;
      r.y = (0.5f * static_cast<float>(H)) * giy;
      ggrid[idx] = r;
    }
  }
}

static int launch_blocks(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(kNumSMs) * 16;
  return static_cast<int>(blocks < cap ? blocks : cap);
}

}  // namespace fovea

using namespace fovea;

extern "C" int fovea_grid_sample_fwd(const float* in, const float* grid, int B, int C, int H, int W, int h, int w,
                                     float* out, fovea_stream_t stream) {
  FOVEA_REQUIRE(in && grid && out, "fovea_grid_sample_fwd: null pointer");
  FOVEA_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && h > 0 && w > 0, "fovea_grid_sample_fwd: bad sizes");
  // experiment switch: FOVEA_GS_TMA=1 routes dense (foveal) output tiles through TMA-staged source boxes
  const char* gs_env = getenv("FOVEA_GS_TMA");   // (read per call: tools/probe_grid_sample.py flips it between launches)
  const int use_tma = gs_env ? atoi(gs_env) : 0;
  bool in_hbm = true;   // a tensor map over pinned HOST memory encodes, but the copy engine of the TMA unit faults on it
  if (use_tma == 1) {   // (illegal instruction, measured): FOVEA_GS_TMA=2 skips this check to reproduce that
    cudaPointerAttributes attr;
    in_hbm = cudaPointerGetAttributes(&attr, in) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
  }
  if (use_tma && in_hbm && C <= kGsMaxC && W % 4 == 0 && (reinterpret_cast<uintptr_t>(in) & 15u) == 0 && B <= 65535) {
    CUtensorMap map;
    if (int rc = make_plane_load_map(&map, in, static_cast<long long>(B) * C, H, W, kGsBoxW, kGsBoxH)) return rc;
    dim3 grid_dim(ceil_div(w, kGsTile), ceil_div(h, kGsTile), B);
    grid_sample_fwd_tma_kernel<<<grid_dim, kGsTile * kGsTile, 0, static_cast<cudaStream_t>(stream)>>>(
        map, in, reinterpret_cast<const float2*>(grid), out, C, H, W, h, w, nullptr);
    return check_launch("fovea_grid_sample_fwd (tma)");
  }
  const long long total = static_cast<long long>(B) * h * w;
  grid_sample_fwd_kernel<<<launch_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, reinterpret_cast<const float2*>(grid), out, B, C, H, W, h * w);
  return check_launch("fovea_grid_sample_fwd");
}

extern "C" int fovea_grid_sample_fwd_u8(const uint8_t* in, const float* grid, int B, int C, int H, int W, int h, int w,
                                        float divisor, float* out, fovea_stream_t stream) {
  FOVEA_REQUIRE(in && grid && out, "fovea_grid_sample_fwd_u8: null pointer");
  FOVEA_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && h > 0 && w > 0 && divisor > 0.f, "fovea_grid_sample_fwd_u8: bad arguments");
  const long long total = static_cast<long long>(B) * h * w;
  grid_sample_fwd_u8_kernel<<<launch_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, reinterpret_cast<const float2*>(grid), out, B, C, H, W, h * w, divisor);
  return check_launch("fovea_grid_sample_fwd_u8");
}

extern "C" int fovea_grid_sample_bwd(const float* grad_out, const float* in, const float* grid, int B, int C, int H,
                                     int W, int h, int w, float* grad_in, float* grad_grid, fovea_stream_t stream) {
  FOVEA_REQUIRE(grad_out && in && grid, "fovea_grid_sample_bwd: null pointer");
  FOVEA_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && h > 0 && w > 0, "fovea_grid_sample_bwd: bad sizes");
  FOVEA_REQUIRE(grad_in || grad_grid, "fovea_grid_sample_bwd: nothing to compute");
  const long long total = static_cast<long long>(B) * h * w;
  const int blocks = launch_blocks(total);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const float2* g2 = reinterpret_cast<const float2*>(grid);
  float2* gg = reinterpret_cast<float2*>(grad_grid);
  if (grad_in && grad_grid)
    grid_sample_bwd_kernel<true, true><<<blocks, 256, 0, s>>>(grad_out, in, g2, grad_in, gg, B, C, H, W, h * w);
  else if (grad_in)
    grid_sample_bwd_kernel<true, false><<<blocks, 256, 0, s>>>(grad_out, in, g2, grad_in, gg, B, C, H, W, h * w);
  else
    grid_sample_bwd_kernel<false, true><<<blocks, 256, 0, s>>>(grad_out, in, g2, grad_in, gg, B, C, H, W, h * w);
  return check_launch("fovea_grid_sample_bwd");
}

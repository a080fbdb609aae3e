// Stage 1: saliency -> sampling grid (forward + backward) and the NHWC grid resize.
//
// Reference: models/models.py:594-637 (create_grid), :510-522 (Gaussian filter + P_basis), :819-825 (padding).
// The reference runs three dense (2Rx+1)x(2Ry+1) convolutions over the padded saliency map; the filter is
// rank-1 and P_basis[0] depends only on the column, P_basis[1] only on the row, so
//     den   = Gx . (Gy . xs)           num_x = Gx . (Gy . (P0 xs))          num_y = Gx . (P1 (Gy . xs))
// i.e. one separable row pass producing two maps and one column pass producing three.  One CTA per image keeps
// every intermediate in shared memory; the padded map is never materialised for the fused padding modes.
#include "common.cuh"

namespace fovea {

struct GridParams {
  int B, gh, gw, Rx, Ry, pad_mode;
  int src_h, src_w;  // layout of xs
  int out_h, out_w;
  float scale_y, scale_x;  // gh/out_h, gw/out_w as aten computes them
};

constexpr int kGridThreads = 512;

__device__ __forceinline__ int src_index(int t_padded, int R, int n, int mode) {
  return mode == FOVEA_PAD_NONE ? t_padded : pad_map(t_padded - R, n, mode);
}

// P_basis entries as the reference builds them: double quotient rounded to fp32 (models/models.py:522).
__device__ __forceinline__ float p_basis_value(int t_padded, int R, int n) {
  return static_cast<float>(static_cast<double>(t_padded - R) / (static_cast<double>(n) - 1.0));
}

__global__ void __launch_bounds__(kGridThreads, 1)
grid_fwd_kernel(const float* __restrict__ xs, const float* __restrict__ g1x, const float* __restrict__ g1y,
                float* __restrict__ grid, float* __restrict__ sums, GridParams p) {
  extern __shared__ float smem[];
  const int Kx = 2 * p.Rx + 1, Ky = 2 * p.Ry + 1;
  const int Gh = p.gh + 2 * p.Rx, Gw = p.gw + 2 * p.Ry;
  float* gx = smem;                  // [Kx]
  float* gy = gx + Kx;               // [Ky]
  float* p0 = gy + Ky;               // [Gw]  P_basis[0] along padded columns
  float* p1 = p0 + Gw;               // [Gh]  P_basis[1] along padded rows
  float* S0 = p1 + Gh;               // [src_h][gw]  row-filtered xs
  float* S1 = S0 + p.src_h * p.gw;   // [src_h][gw]  row-filtered P0*xs
  float* raw = S1 + p.src_h * p.gw;  // [2][gh][gw]  clamped grid before the resize

  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const float* xb = xs + static_cast<size_t>(b) * p.src_h * p.src_w;

  for (int i = tid; i < Kx; i += kGridThreads) gx[i] = g1x[i];
  for (int i = tid; i < Ky; i += kGridThreads) gy[i] = g1y[i];
  for (int i = tid; i < Gw; i += kGridThreads) p0[i] = p_basis_value(i, p.Ry, p.gw);
  for (int i = tid; i < Gh; i += kGridThreads) p1[i] = p_basis_value(i, p.Rx, p.gh);
  __syncthreads();

  // ---- row pass: S0[r][j] = sum_b gy[b] xs[r][col(j+b)],  S1 likewise with P0(j+b) folded in
  for (int idx = tid; idx < p.src_h * p.gw; idx += kGridThreads) {
    const int r = idx / p.gw, j = idx - r * p.gw;
    const float* row = xb + static_cast<size_t>(r) * p.src_w;
    float s0 = 0.f, s1 = 0.f;
    for (int t = 0; t < Ky; ++t) {
      const int c = src_index(j + t, p.Ry, p.gw, p.pad_mode);
      if (c >= 0) {
        const float v = __ldg(row + c);
        s0 = fmaf(gy[t], v, s0);
        s1 = fmaf(gy[t], p0[j + t] * v, s1);
      }
    }
    S0[idx] = s0;
    S1[idx] = s1;
  }
  __syncthreads();

  // ---- column pass + quotient + clamp (models/models.py:609-615)
  float* sums_b = sums ? sums + static_cast<size_t>(b) * 3 * p.gh * p.gw : nullptr;
  for (int idx = tid; idx < p.gh * p.gw; idx += kGridThreads) {
    const int i = idx / p.gw, j = idx - i * p.gw;
    float den = 0.f, nx = 0.f, ny = 0.f;
    for (int t = 0; t < Kx; ++t) {
      const int r = src_index(i + t, p.Rx, p.gh, p.pad_mode);
      if (r >= 0) {
        const float a = S0[r * p.gw + j];
        den = fmaf(gx[t], a, den);
        ny = fmaf(gx[t], p1[i + t] * a, ny);
        nx = fmaf(gx[t], S1[r * p.gw + j], nx);
      }
    }
    if (sums_b) {
      sums_b[idx] = den;
      sums_b[p.gh * p.gw + idx] = nx;
      sums_b[2 * p.gh * p.gw + idx] = ny;
    }
    raw[idx] = fminf(fmaxf(__fadd_rn(__fmul_rn(__fdiv_rn(nx, den), 2.f), -1.f), -1.f), 1.f);
    raw[p.gh * p.gw + idx] = fminf(fmaxf(__fadd_rn(__fmul_rn(__fdiv_rn(ny, den), 2.f), -1.f), -1.f), 1.f);
  }
  __syncthreads();

  // ---- nn.Upsample(bilinear) to the task size + NCHW->NHWC (models/models.py:621-637)
  float2* gout = reinterpret_cast<float2*>(grid) + static_cast<size_t>(b) * p.out_h * p.out_w;
  const bool identity = (p.out_h == p.gh) && (p.out_w == p.gw);
  for (int idx = tid; idx < p.out_h * p.out_w; idx += kGridThreads) {
    float2 o;
    if (identity) {
      o.x = raw[idx];
      o.y = raw[p.gh * p.gw + idx];
    } else {
      const int oy = idx / p.out_w, ox = idx - oy * p.out_w;
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      bilinear_src(oy, p.scale_y, p.gh, y0, y1, ly0, ly1);
      bilinear_src(ox, p.scale_x, p.gw, x0, x1, lx0, lx1);
      const float* rx = raw;
      const float* ry = raw + p.gh * p.gw;
      o.x = ly0 * (lx0 * rx[y0 * p.gw + x0] + lx1 * rx[y0 * p.gw + x1]) +
            ly1 * (lx0 * rx[y1 * p.gw + x0] + lx1 * rx[y1 * p.gw + x1]);
      o.y = ly0 * (lx0 * ry[y0 * p.gw + x0] + lx1 * ry[y0 * p.gw + x1]) +
            ly1 * (lx0 * ry[y1 * p.gw + x0] + lx1 * ry[y1 * p.gw + x1]);
    }
    gout[idx] = o;
  }
}

// Backward w.r.t. xs.  grad flows: resize^T -> clamp mask -> quotient rule -> column pass^T -> row pass^T -> pad^T.
__global__ void __launch_bounds__(kGridThreads, 1)
grid_bwd_kernel(const float* __restrict__ grad_grid, const float* __restrict__ sums,
                const float* __restrict__ g1x, const float* __restrict__ g1y, float* __restrict__ grad_xs,
                GridParams p) {
  extern __shared__ float smem[];
  const int Kx = 2 * p.Rx + 1, Ky = 2 * p.Ry + 1;
  const int Gh = p.gh + 2 * p.Rx, Gw = p.gw + 2 * p.Ry;
  const int n = p.gh * p.gw;
  float* gx = smem;
  float* gy = gx + Kx;
  float* p0 = gy + Ky;
  float* p1 = p0 + Gw;
  float* Dd = p1 + Gh;                // [gh][gw] dL/d den
  float* Dx = Dd + n;                 // dL/d num_x   (first used as dL/d raw_x)
  float* Dy = Dx + n;                 // dL/d num_y   (first used as dL/d raw_y)
  float* dS0 = Dy + n;                // [src_h][gw]
  float* dS1 = dS0 + p.src_h * p.gw;  // [src_h][gw]
  float* acc = dS1 + p.src_h * p.gw;  // [src_h][src_w], fused padding modes only

  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const bool fused = p.pad_mode != FOVEA_PAD_NONE;

  for (int i = tid; i < Kx; i += kGridThreads) gx[i] = g1x[i];
  for (int i = tid; i < Ky; i += kGridThreads) gy[i] = g1y[i];
  for (int i = tid; i < Gw; i += kGridThreads) p0[i] = p_basis_value(i, p.Ry, p.gw);
  for (int i = tid; i < Gh; i += kGridThreads) p1[i] = p_basis_value(i, p.Rx, p.gh);
  for (int i = tid; i < 2 * n; i += kGridThreads) Dx[i] = 0.f;  // Dx and Dy are contiguous
  for (int i = tid; i < 2 * p.src_h * p.gw; i += kGridThreads) dS0[i] = 0.f;
  if (fused)
    for (int i = tid; i < p.src_h * p.src_w; i += kGridThreads) acc[i] = 0.f;
  __syncthreads();

  // ---- resize^T: scatter dL/d grid[oy][ox] onto the raw lattice
  const float2* gg = reinterpret_cast<const float2*>(grad_grid) + static_cast<size_t>(b) * p.out_h * p.out_w;
  const bool identity = (p.out_h == p.gh) && (p.out_w == p.gw);
  for (int idx = tid; idx < p.out_h * p.out_w; idx += kGridThreads) {
    const float2 g = gg[idx];
    if (identity) {
      Dx[idx] = g.x;
      Dy[idx] = g.y;
    } else {
      const int oy = idx / p.out_w, ox = idx - oy * p.out_w;
      int y0, y1, x0, x1;
      float ly0, ly1, lx0, lx1;
      bilinear_src(oy, p.scale_y, p.gh, y0, y1, ly0, ly1);
      bilinear_src(ox, p.scale_x, p.gw, x0, x1, lx0, lx1);
      atomicAdd(&Dx[y0 * p.gw + x0], ly0 * lx0 * g.x);
      atomicAdd(&Dx[y0 * p.gw + x1], ly0 * lx1 * g.x);
      atomicAdd(&Dx[y1 * p.gw + x0], ly1 * lx0 * g.x);
      atomicAdd(&Dx[y1 * p.gw + x1], ly1 * lx1 * g.x);
      atomicAdd(&Dy[y0 * p.gw + x0], ly0 * lx0 * g.y);
      atomicAdd(&Dy[y0 * p.gw + x1], ly0 * lx1 * g.y);
      atomicAdd(&Dy[y1 * p.gw + x0], ly1 * lx0 * g.y);
      atomicAdd(&Dy[y1 * p.gw + x1], ly1 * lx1 * g.y);
    }
  }
  __syncthreads();

  // ---- clamp mask (inclusive, as torch.clamp backward) + quotient rule
  const float* sb = sums + static_cast<size_t>(b) * 3 * n;
  for (int idx = tid; idx < n; idx += kGridThreads) {
    const float den = sb[idx], nx = sb[n + idx], ny = sb[2 * n + idx];
    const float vx = __fadd_rn(__fmul_rn(__fdiv_rn(nx, den), 2.f), -1.f);
    const float vy = __fadd_rn(__fmul_rn(__fdiv_rn(ny, den), 2.f), -1.f);
    const float gxr = (vx >= -1.f && vx <= 1.f) ? Dx[idx] : 0.f;
    const float gyr = (vy >= -1.f && vy <= 1.f) ? Dy[idx] : 0.f;
    const float inv = 1.f / den;
    const float dnx = 2.f * gxr * inv;
    const float dny = 2.f * gyr * inv;
    Dx[idx] = dnx;
    Dy[idx] = dny;
    Dd[idx] = -(dnx * nx + dny * ny) * inv;
  }
  __syncthreads();

  // ---- column pass^T: for every padded row t, gather over output rows i = t-a
  for (int idx = tid; idx < Gh * p.gw; idx += kGridThreads) {
    const int t = idx / p.gw, j = idx - t * p.gw;
    const int r = src_index(t, p.Rx, p.gh, p.pad_mode);
    if (r < 0) continue;
    const int i_lo = max(0, t - (Kx - 1)), i_hi = min(p.gh - 1, t);
    float a0 = 0.f, ay = 0.f, a1 = 0.f;
    for (int i = i_lo; i <= i_hi; ++i) {
      const float g = gx[t - i];
      a0 = fmaf(g, Dd[i * p.gw + j], a0);
      ay = fmaf(g, Dy[i * p.gw + j], ay);
      a1 = fmaf(g, Dx[i * p.gw + j], a1);
    }
    a0 = fmaf(p1[t], ay, a0);
    if (fused) {
      atomicAdd(&dS0[r * p.gw + j], a0);
      atomicAdd(&dS1[r * p.gw + j], a1);
    } else {
      dS0[r * p.gw + j] = a0;
      dS1[r * p.gw + j] = a1;
    }
  }
  __syncthreads();

  // ---- row pass^T: for every padded column u, gather over output columns j = u-b
  float* out_b = grad_xs + static_cast<size_t>(b) * p.src_h * p.src_w;
  for (int idx = tid; idx < p.src_h * Gw; idx += kGridThreads) {
    const int r = idx / Gw, u = idx - r * Gw;
    const int c = src_index(u, p.Ry, p.gw, p.pad_mode);
    if (c < 0) continue;
    const int j_lo = max(0, u - (Ky - 1)), j_hi = min(p.gw - 1, u);
    float a0 = 0.f, a1 = 0.f;
    for (int j = j_lo; j <= j_hi; ++j) {
      const float g = gy[u - j];
      a0 = fmaf(g, dS0[r * p.gw + j], a0);
      a1 = fmaf(g, dS1[r * p.gw + j], a1);
    }
    const float v = fmaf(p0[u], a1, a0);
    if (fused)
      atomicAdd(&acc[r * p.src_w + c], v);
    else
      out_b[r * p.src_w + c] = v;
  }
  if (fused) {
    __syncthreads();
    for (int i = tid; i < p.src_h * p.src_w; i += kGridThreads) out_b[i] = acc[i];
  }
}

__global__ void grid_resize_kernel(const float2* __restrict__ in, float2* __restrict__ out, int B, int ih, int iw,
                                   int oh, int ow, float sy, float sx) {
  const int total = B * oh * ow;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int b = idx / (oh * ow), rem = idx - b * oh * ow;
    const int oy = rem / ow, ox = rem - oy * ow;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bilinear_src(oy, sy, ih, y0, y1, ly0, ly1);
    bilinear_src(ox, sx, iw, x0, x1, lx0, lx1);
    const float2* ib = in + static_cast<size_t>(b) * ih * iw;
    const float2 a = ib[y0 * iw + x0], bb = ib[y0 * iw + x1], c = ib[y1 * iw + x0], d = ib[y1 * iw + x1];
    float2 o;
    o.x = ly0 * (lx0 * a.x + lx1 * bb.x) + ly1 * (lx0 * c.x + lx1 * d.x);
    o.y = ly0 * (lx0 * a.y + lx1 * bb.y) + ly1 * (lx0 * c.y + lx1 * d.y);
    out[idx] = o;
  }
}

__global__ void grid_resize_bwd_kernel(const float2* __restrict__ gout, float* __restrict__ gin, int B, int ih, int iw,
                                       int oh, int ow, float sy, float sx) {
  const int total = B * oh * ow;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int b = idx / (oh * ow), rem = idx - b * oh * ow;
    const int oy = rem / ow, ox = rem - oy * ow;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bilinear_src(oy, sy, ih, y0, y1, ly0, ly1);
    bilinear_src(ox, sx, iw, x0, x1, lx0, lx1);
    const float2 g = gout[idx];
    float* gb = gin + static_cast<size_t>(b) * ih * iw * 2;
    atomicAdd(&gb[(y0 * iw + x0) * 2 + 0], ly0 * lx0 * g.x);
    atomicAdd(&gb[(y0 * iw + x0) * 2 + 1], ly0 * lx0 * g.y);
    atomicAdd(&gb[(y0 * iw + x1) * 2 + 0], ly0 * lx1 * g.x);
    atomicAdd(&gb[(y0 * iw + x1) * 2 + 1], ly0 * lx1 * g.y);
    atomicAdd(&gb[(y1 * iw + x0) * 2 + 0], ly1 * lx0 * g.x);
    atomicAdd(&gb[(y1 * iw + x0) * 2 + 1], ly1 * lx0 * g.y);
    atomicAdd(&gb[(y1 * iw + x1) * 2 + 0], ly1 * lx1 * g.x);
    atomicAdd(&gb[(y1 * iw + x1) * 2 + 1], ly1 * lx1 * g.y);
  }
}

static int fill_params(GridParams& p, int B, int gh, int gw, int Rx, int Ry, int pad_mode, int out_h, int out_w) {
  FOVEA_REQUIRE(B > 0 && gh > 1 && gw > 1 && Rx >= 0 && Ry >= 0 && out_h > 0 && out_w > 0,
                "fovea_grid: bad sizes B=%d gh=%d gw=%d Rx=%d Ry=%d out=%dx%d", B, gh, gw, Rx, Ry, out_h, out_w);
  FOVEA_REQUIRE(pad_mode >= FOVEA_PAD_NONE && pad_mode <= FOVEA_PAD_ZERO, "fovea_grid: bad pad_mode %d", pad_mode);
  FOVEA_REQUIRE(pad_mode != FOVEA_PAD_REFLECT || (Rx < gh && Ry < gw),
                "fovea_grid: reflect padding needs R < size (Rx=%d gh=%d Ry=%d gw=%d)", Rx, gh, Ry, gw);
  p.B = B; p.gh = gh; p.gw = gw; p.Rx = Rx; p.Ry = Ry; p.pad_mode = pad_mode;
  p.src_h = pad_mode == FOVEA_PAD_NONE ? gh + 2 * Rx : gh;
  p.src_w = pad_mode == FOVEA_PAD_NONE ? gw + 2 * Ry : gw;
  p.out_h = out_h; p.out_w = out_w;
  p.scale_y = static_cast<float>(gh) / static_cast<float>(out_h);
  p.scale_x = static_cast<float>(gw) / static_cast<float>(out_w);
  return FOVEA_OK;
}

}  // namespace fovea

using namespace fovea;

extern "C" int fovea_grid_fwd(const float* xs, int B, int gh, int gw, int Rx, int Ry, int pad_mode, const float* g1x,
                              const float* g1y, int out_h, int out_w, float* grid, float* sums,
                              fovea_stream_t stream) {
  FOVEA_REQUIRE(xs && g1x && g1y && grid, "fovea_grid_fwd: null pointer");
  GridParams p;
  if (int rc = fill_params(p, B, gh, gw, Rx, Ry, pad_mode, out_h, out_w)) return rc;
  const size_t smem = sizeof(float) * (static_cast<size_t>(2 * Rx + 1) + (2 * Ry + 1) + (gw + 2 * Ry) + (gh + 2 * Rx) +
                                       2 * static_cast<size_t>(p.src_h) * gw + 2 * static_cast<size_t>(gh) * gw);
  if (smem > 227 * 1024) {
    set_error("fovea_grid_fwd: %zu B of shared memory needed (> 227 KB) for gh=%d gw=%d Rx=%d Ry=%d", smem, gh, gw, Rx,
              Ry);
    return FOVEA_ERR_CAPACITY;
  }
  FOVEA_CUDA(cudaFuncSetAttribute(grid_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  grid_fwd_kernel<<<B, kGridThreads, smem, static_cast<cudaStream_t>(stream)>>>(xs, g1x, g1y, grid, sums, p);
  return check_launch("fovea_grid_fwd");
}

extern "C" int fovea_grid_bwd(const float* grad_grid, const float* sums, int B, int gh, int gw, int Rx, int Ry,
                              int pad_mode, const float* g1x, const float* g1y, int out_h, int out_w, float* grad_xs,
                              fovea_stream_t stream) {
  FOVEA_REQUIRE(grad_grid && sums && g1x && g1y && grad_xs, "fovea_grid_bwd: null pointer");
  GridParams p;
  if (int rc = fill_params(p, B, gh, gw, Rx, Ry, pad_mode, out_h, out_w)) return rc;
  size_t words = static_cast<size_t>(2 * Rx + 1) + (2 * Ry + 1) + (gw + 2 * Ry) + (gh + 2 * Rx) +
                 3 * static_cast<size_t>(gh) * gw + 2 * static_cast<size_t>(p.src_h) * gw;
  if (pad_mode != FOVEA_PAD_NONE) words += static_cast<size_t>(p.src_h) * p.src_w;
  const size_t smem = words * sizeof(float);
  if (smem > 227 * 1024) {
    set_error("fovea_grid_bwd: %zu B of shared memory needed (> 227 KB)", smem);
    return FOVEA_ERR_CAPACITY;
  }
  FOVEA_CUDA(cudaFuncSetAttribute(grid_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  grid_bwd_kernel<<<B, kGridThreads, smem, static_cast<cudaStream_t>(stream)>>>(grad_grid, sums, g1x, g1y, grad_xs, p);
  return check_launch("fovea_grid_bwd");
}

extern "C" int fovea_grid_resize(const float* in, int B, int ih, int iw, int oh, int ow, float* out,
                                 fovea_stream_t stream) {
  FOVEA_REQUIRE(in && out && B > 0 && ih > 0 && iw > 0 && oh > 0 && ow > 0, "fovea_grid_resize: bad arguments");
  const int total = B * oh * ow;
  const int blocks = min(ceil_div(total, 256), kNumSMs * 8);
  grid_resize_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out), B, ih, iw, oh, ow,
      static_cast<float>(ih) / static_cast<float>(oh), static_cast<float>(iw) / static_cast<float>(ow));
  return check_launch("fovea_grid_resize");
}

extern "C" int fovea_grid_resize_bwd(const float* grad_out, int B, int ih, int iw, int oh, int ow, float* grad_in,
                                     fovea_stream_t stream) {
  FOVEA_REQUIRE(grad_out && grad_in && B > 0 && ih > 0 && iw > 0 && oh > 0 && ow > 0,
                "fovea_grid_resize_bwd: bad arguments");
  FOVEA_CUDA(cudaMemsetAsync(grad_in, 0, sizeof(float) * 2 * static_cast<size_t>(B) * ih * iw,
                             static_cast<cudaStream_t>(stream)));
  const int total = B * oh * ow;
  const int blocks = min(ceil_div(total, 256), kNumSMs * 8);
  grid_resize_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(grad_out), grad_in, B, ih, iw, oh, ow,
      static_cast<float>(ih) / static_cast<float>(oh), static_cast<float>(iw) / static_cast<float>(ow));
  return check_launch("fovea_grid_resize_bwd");
}

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

constexpr int kGridThreads = 512;   // backward kernel
constexpr int kFwdThreads = 1024;   // forward kernel
constexpr int kTile = 4;            // outputs per thread and pass (register tile along the filtered axis)

__device__ __forceinline__ int src_index(int t_padded, int R, int n, int mode) {
  return mode == FOVEA_PAD_NONE ? t_padded : pad_map(t_padded - R, n, mode);
}

// P_basis entries as the reference builds them: double quotient rounded to fp32 (models/models.py:522).
__device__ __forceinline__ float p_basis_value(int t_padded, int R, int n) {
  return static_cast<float>(static_cast<double>(t_padded - R) / (static_cast<double>(n) - 1.0));
}

// Shared-memory plan of the forward kernel (floats unless noted), in this order:
//   gxz [Kx + 2*(kTile-1)]   column-pass filter, zero-padded by kTile-1 on both sides (no bounds tests in the loops)
//   gyz [Ky + 2*(kTile-1)]   row-pass filter, likewise
//   p0  [Gw]  p1 [Gh]        P_basis along padded columns / rows
//   rmap [Gh] (int)          padded row -> row of S0/S1 (src_h = the all-zero row for zero padding)
//   cmap [Gw] (int)          padded col -> source column (-1 = zero tap)          (fused padding only)
//   S0, S1 [(src_h+1)][gw]   row-filtered xs and P0*xs (+ one zero row)
//   U   max([src_h][Gw] staged padded rows (fused padding), [2][gh][gw] clamped raw grid)   -- reused
struct FwdSmem {
  float *gxz, *gyz, *p0, *p1, *S0, *S1, *U;
  int *rmap, *cmap;
};

__host__ __device__ inline size_t fwd_smem_floats(int gh, int gw, int Rx, int Ry, int src_h, bool staged) {
  const int Kx = 2 * Rx + 1, Ky = 2 * Ry + 1, Gh = gh + 2 * Rx, Gw = gw + 2 * Ry;
  size_t u = 2 * static_cast<size_t>(gh) * gw;
  if (staged && static_cast<size_t>(src_h) * Gw > u) u = static_cast<size_t>(src_h) * Gw;
  return static_cast<size_t>(Kx + Ky + 4 * (kTile - 1)) + 2 * (Gw + Gh) + 2 * static_cast<size_t>(src_h + 1) * gw + u;
}

// kStaged: the (fused-padding) source rows are expanded to padded rows in shared memory once, so the row pass reads
// them without index arithmetic; for FOVEA_PAD_NONE the padded map is read from global memory (L1) directly.
template <bool kStaged>
__global__ void __launch_bounds__(kFwdThreads, 1)
grid_fwd_kernel(const float* __restrict__ xs, const float* __restrict__ g1x, const float* __restrict__ g1y,
                float* __restrict__ grid, float* __restrict__ sums, GridParams p) {
  extern __shared__ float smem[];
  const int Kx = 2 * p.Rx + 1, Ky = 2 * p.Ry + 1;
  const int Gh = p.gh + 2 * p.Rx, Gw = p.gw + 2 * p.Ry;
  constexpr int Z = kTile - 1;
  FwdSmem m;
  m.gxz = smem;
  m.gyz = m.gxz + Kx + 2 * Z;
  m.p0 = m.gyz + Ky + 2 * Z;
  m.p1 = m.p0 + Gw;
  m.rmap = reinterpret_cast<int*>(m.p1 + Gh);
  m.cmap = m.rmap + Gh;
  m.S0 = reinterpret_cast<float*>(m.cmap + Gw);
  m.S1 = m.S0 + (p.src_h + 1) * p.gw;
  m.U = m.S1 + (p.src_h + 1) * p.gw;

  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const float* xb = xs + static_cast<size_t>(b) * p.src_h * p.src_w;

  for (int i = tid; i < Kx + 2 * Z; i += kFwdThreads) m.gxz[i] = (i >= Z && i < Z + Kx) ? g1x[i - Z] : 0.f;
  for (int i = tid; i < Ky + 2 * Z; i += kFwdThreads) m.gyz[i] = (i >= Z && i < Z + Ky) ? g1y[i - Z] : 0.f;
  for (int i = tid; i < Gw; i += kFwdThreads) {
    m.p0[i] = p_basis_value(i, p.Ry, p.gw);
    m.cmap[i] = src_index(i, p.Ry, p.gw, p.pad_mode);
  }
  for (int i = tid; i < Gh; i += kFwdThreads) {
    m.p1[i] = p_basis_value(i, p.Rx, p.gh);
    const int r = src_index(i, p.Rx, p.gh, p.pad_mode);
    m.rmap[i] = r < 0 ? p.src_h : r;
  }
  for (int i = tid; i < p.gw; i += kFwdThreads) m.S0[p.src_h * p.gw + i] = m.S1[p.src_h * p.gw + i] = 0.f;
  __syncthreads();
  if (kStaged) {  // padded rows: xp[r][u] = xs[r][cmap[u]] (0 for a zero tap)
    for (int idx = tid; idx < p.src_h * Gw; idx += kFwdThreads) {
      const int r = idx / Gw, u = idx - r * Gw;
      const int c = m.cmap[u];
      m.U[idx] = c >= 0 ? __ldg(xb + r * p.src_w + c) : 0.f;
    }
    __syncthreads();
  }

  // ---- row pass: S0[r][j] = sum_t gy[t] xp[r][j+t],  S1 likewise with P0(j+t) folded in.  One thread produces
  // kTile consecutive j of one row: every padded element is loaded once and feeds 2*kTile FMAs; the filter taps slide
  // through a register window.  Each output still accumulates its taps in ascending order (as the untiled loop did).
  const int jt = ceil_div(p.gw, kTile);
  for (int item = tid; item < p.src_h * jt; item += kFwdThreads) {
    const int r = item / jt, j0 = (item - r * jt) * kTile;
    const float* row = kStaged ? m.U + r * Gw : xb + static_cast<size_t>(r) * p.src_w;
    float a0[kTile], a1[kTile], g[kTile];
#pragma unroll
    for (int k = 0; k < kTile; ++k) { a0[k] = a1[k] = 0.f; g[k] = 0.f; }
    // output j0+k uses tap t = u - (j0+k); window g[k] = gyz[Z + u - j0 - k]
    const int u_end = min(j0 + kTile - 1 + Ky, Gw);
    for (int u = j0; u < u_end; ++u) {
#pragma unroll
      for (int k = kTile - 1; k > 0; --k) g[k] = g[k - 1];
      g[0] = m.gyz[Z + u - j0];
      const float v = kStaged ? row[u] : __ldg(row + u);
      const float pv = m.p0[u] * v;
#pragma unroll
      for (int k = 0; k < kTile; ++k) {
        a0[k] = fmaf(g[k], v, a0[k]);
        a1[k] = fmaf(g[k], pv, a1[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < kTile; ++k)
      if (j0 + k < p.gw) {
        m.S0[r * p.gw + j0 + k] = a0[k];
        m.S1[r * p.gw + j0 + k] = a1[k];
      }
  }
  __syncthreads();

  // ---- column pass + quotient + clamp (models/models.py:609-615): kTile consecutive i of one column per thread
  float* raw = m.U;  // the staged rows are dead now
  float* sums_b = sums ? sums + static_cast<size_t>(b) * 3 * p.gh * p.gw : nullptr;
  const int it = ceil_div(p.gh, kTile);
  for (int item = tid; item < it * p.gw; item += kFwdThreads) {
    const int ib = item / p.gw, j = item - ib * p.gw, i0 = ib * kTile;
    float den[kTile], nx[kTile], ny[kTile], g[kTile];
#pragma unroll
    for (int k = 0; k < kTile; ++k) { den[k] = nx[k] = ny[k] = 0.f; g[k] = 0.f; }
    const int v_end = min(i0 + kTile - 1 + Kx, Gh);
    for (int v = i0; v < v_end; ++v) {
#pragma unroll
      for (int k = kTile - 1; k > 0; --k) g[k] = g[k - 1];
      g[0] = m.gxz[Z + v - i0];
      const int r = m.rmap[v];
      const float a = m.S0[r * p.gw + j], s1 = m.S1[r * p.gw + j];
      const float pa = m.p1[v] * a;
#pragma unroll
      for (int k = 0; k < kTile; ++k) {
        den[k] = fmaf(g[k], a, den[k]);
        ny[k] = fmaf(g[k], pa, ny[k]);
        nx[k] = fmaf(g[k], s1, nx[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < kTile; ++k) {
      const int i = i0 + k;
      if (i >= p.gh) break;
      const int idx = i * p.gw + j;
      if (sums_b) {
        sums_b[idx] = den[k];
        sums_b[p.gh * p.gw + idx] = nx[k];
        sums_b[2 * p.gh * p.gw + idx] = ny[k];
      }
      raw[idx] = fminf(fmaxf(__fadd_rn(__fmul_rn(__fdiv_rn(nx[k], den[k]), 2.f), -1.f), -1.f), 1.f);
      raw[p.gh * p.gw + idx] = fminf(fmaxf(__fadd_rn(__fmul_rn(__fdiv_rn(ny[k], den[k]), 2.f), -1.f), -1.f), 1.f);
    }
  }
  __syncthreads();

  // ---- nn.Upsample(bilinear) to the task size + NCHW->NHWC (models/models.py:621-637)
  float2* gout = reinterpret_cast<float2*>(grid) + static_cast<size_t>(b) * p.out_h * p.out_w;
  const bool identity = (p.out_h == p.gh) && (p.out_w == p.gw);
  for (int idx = tid; idx < p.out_h * p.out_w; idx += kFwdThreads) {
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
  const bool staged = pad_mode != FOVEA_PAD_NONE;
  const size_t smem = sizeof(float) * fwd_smem_floats(gh, gw, Rx, Ry, p.src_h, staged);
  if (smem > 227 * 1024) {
    set_error("fovea_grid_fwd: %zu B of shared memory needed (> 227 KB) for gh=%d gw=%d Rx=%d Ry=%d", smem, gh, gw, Rx,
              Ry);
    return FOVEA_ERR_CAPACITY;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (staged) {
    FOVEA_CUDA(cudaFuncSetAttribute(grid_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    grid_fwd_kernel<true><<<B, kFwdThreads, smem, s>>>(xs, g1x, g1y, grid, sums, p);
  } else {
    FOVEA_CUDA(cudaFuncSetAttribute(grid_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    grid_fwd_kernel<false><<<B, kFwdThreads, smem, s>>>(xs, g1x, g1y, grid, sums, p);
  }
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

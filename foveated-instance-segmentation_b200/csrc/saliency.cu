// Stage 0 / 2 of the saliency branch: the inputs and the normalisation either side of the (stock) saliency network.
//
//   fovea_saliency_input    models/models.py:684-705  focus map + b_imresize(x, input_size, 'bilinear') + the two
//                           torch.cat -> x_low [B, C+2, HS, WS] in ONE launch (the reference: an int64 index grid
//                           repeated per batch, sqrt/div/pow, F.interpolate, two cats = ~12 launches).  The image
//                           may be fp32 or uint8 (ToTensor's /255 folded in) and may
//                           live in pinned host memory: only the 4 taps per low-res pixel are touched.
//   fovea_saliency_softmax  models/models.py:715-723  nn.Softmax over the gh*gw saliency logits of a frame
//                           (+ backward for training); the reference follows it with a NaN assert (host sync).
#include "common.cuh"

namespace fovea {

__device__ __forceinline__ float load_sample(const float* p, float) { return __ldg(p); }
__device__ __forceinline__ float load_sample(const unsigned char* p, float divisor) {
  return __fdiv_rn(static_cast<float>(__ldg(p)), divisor);
}

// One thread per low-res pixel (b, i, j); lanes cover consecutive j, so the stores of every channel plane are
// coalesced.  Bilinear taps follow aten's upsample_bilinear2d (align_corners=False): source index
// max(scale*(dst+0.5)-0.5, 0), value = h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11).
template <typename T>
__global__ void __launch_bounds__(256)
saliency_input_kernel(const T* __restrict__ img, const float* __restrict__ focus_point, float* __restrict__ out, int B,
                      int C, int H, int W, int HS, int WS, float scale_h, float scale_w, float max_dist,
                      float divisor) {
  const int hw = HS * WS;
  const int total = B * hw;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int b = idx / hw;
    const int pix = idx - b * hw;
    const int i = pix / WS, j = pix - i * WS;
    int y0, y1, x0, x1;
    float h0, h1, w0, w1;
    bilinear_src(i, scale_h, H, y0, y1, h0, h1);
    bilinear_src(j, scale_w, W, x0, x1, w0, w1);
    const size_t plane = static_cast<size_t>(H) * W;
    const T* src = img + static_cast<size_t>(b) * C * plane;
    float* dst = out + static_cast<size_t>(b) * (C + 2) * hw + pix;
    const size_t o00 = static_cast<size_t>(y0) * W + x0, o01 = static_cast<size_t>(y0) * W + x1;
    const size_t o10 = static_cast<size_t>(y1) * W + x0, o11 = static_cast<size_t>(y1) * W + x1;
    for (int c = 0; c < C; ++c) {
      const T* s = src + c * plane;
      const float v00 = load_sample(s + o00, divisor), v01 = load_sample(s + o01, divisor);
      const float v10 = load_sample(s + o10, divisor), v11 = load_sample(s + o11, divisor);
      dst[static_cast<size_t>(c) * hw] = h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11);
    }
    // focus map (:687-694): ((i - hidx)^2 + (j - widx)^2) / (HS^2 + WS^2), evaluated as the reference does
    // (sqrt, divide by the diagonal, square) so that the roundings agree to the last bit or two
    const float hidx = __ldg(focus_point + 2 * b) * static_cast<float>(HS - 1);
    const float widx = __ldg(focus_point + 2 * b + 1) * static_cast<float>(WS - 1);
    const float dh = static_cast<float>(i) - hidx, dw = static_cast<float>(j) - widx;
    const float f = __fdiv_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(dh, dh), __fmul_rn(dw, dw))), max_dist);
    const float f2 = f * f;
    dst[static_cast<size_t>(C) * hw] = f2;
    dst[static_cast<size_t>(C + 1) * hw] = f2;
  }
}

constexpr int kSoftmaxThreads = 256;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float other = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, other) : v + other;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // red may still be read by the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int k = 1; k < kSoftmaxThreads / 32; ++k) r = is_max ? fmaxf(r, red[k]) : r + red[k];
  return r;
}

// One CTA per frame: max, sum of exp(x - max), quotient.  A NaN logit poisons the whole frame, as in torch.
__global__ void __launch_bounds__(kSoftmaxThreads)
saliency_softmax_kernel(const float* __restrict__ logits, float* __restrict__ xs, int n) {
  __shared__ float red[kSoftmaxThreads / 32];
  const float* in = logits + static_cast<size_t>(blockIdx.x) * n;
  float* out = xs + static_cast<size_t>(blockIdx.x) * n;
  float m = -INFINITY;
  bool nan = false;
  for (int k = threadIdx.x; k < n; k += kSoftmaxThreads) {
    const float v = in[k];
    nan |= (v != v);
    m = fmaxf(m, v);
  }
  m = block_reduce(nan ? INFINITY : m, true, red);
  const bool poisoned = __syncthreads_or(nan);
  float s = 0.f;
  for (int k = threadIdx.x; k < n; k += kSoftmaxThreads) s += expf(in[k] - m);
  s = block_reduce(s, false, red);
  for (int k = threadIdx.x; k < n; k += kSoftmaxThreads)
    out[k] = poisoned ? __int_as_float(0x7fc00000) : __fdiv_rn(expf(in[k] - m), s);
}

// grad_logits = xs * (grad_xs - sum(grad_xs * xs))
__global__ void __launch_bounds__(kSoftmaxThreads)
saliency_softmax_bwd_kernel(const float* __restrict__ xs, const float* __restrict__ grad_xs,
                            float* __restrict__ grad_logits, int n) {
  __shared__ float red[kSoftmaxThreads / 32];
  const size_t base = static_cast<size_t>(blockIdx.x) * n;
  float dot = 0.f;
  for (int k = threadIdx.x; k < n; k += kSoftmaxThreads) dot = fmaf(grad_xs[base + k], xs[base + k], dot);
  dot = block_reduce(dot, false, red);
  for (int k = threadIdx.x; k < n; k += kSoftmaxThreads)
    grad_logits[base + k] = xs[base + k] * (grad_xs[base + k] - dot);
}

}  // namespace fovea

using namespace fovea;

extern "C" int fovea_saliency_input(const void* img, int img_u8, float divisor, const float* focus_point, int B, int C,
                                    int H, int W, int HS, int WS, float* out, fovea_stream_t stream) {
  FOVEA_REQUIRE(img && focus_point && out, "fovea_saliency_input: null pointer");
  FOVEA_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && HS > 0 && WS > 0, "fovea_saliency_input: bad sizes");
  FOVEA_REQUIRE(static_cast<long long>(B) * HS * WS < (1ll << 31), "fovea_saliency_input: B*HS*WS too large");
  FOVEA_REQUIRE(!img_u8 || divisor > 0.f, "fovea_saliency_input: divisor must be positive");
  const int total = B * HS * WS;
  const int blocks = min(ceil_div(total, 256), kNumSMs * 8);
  const float sh = static_cast<float>(H) / static_cast<float>(HS), sw = static_cast<float>(W) / static_cast<float>(WS);
  // np.sqrt(HS**2 + WS**2) is a float64 scalar that torch rounds to fp32 for the division (:686, :693)
  const float max_dist = static_cast<float>(sqrt(static_cast<double>(HS) * HS + static_cast<double>(WS) * WS));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (img_u8)
    saliency_input_kernel<unsigned char><<<blocks, 256, 0, s>>>(static_cast<const unsigned char*>(img), focus_point,
                                                                 out, B, C, H, W, HS, WS, sh, sw, max_dist, divisor);
  else
    saliency_input_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(img), focus_point, out, B, C, H, W,
                                                         HS, WS, sh, sw, max_dist, 1.f);
  return check_launch("fovea_saliency_input");
}

extern "C" int fovea_saliency_softmax(const float* logits, int B, int n, float* xs, fovea_stream_t stream) {
  FOVEA_REQUIRE(logits && xs && B > 0 && n > 0, "fovea_saliency_softmax: bad arguments");
  saliency_softmax_kernel<<<B, kSoftmaxThreads, 0, static_cast<cudaStream_t>(stream)>>>(logits, xs, n);
  return check_launch("fovea_saliency_softmax");
}

extern "C" int fovea_saliency_softmax_bwd(const float* xs, const float* grad_xs, int B, int n, float* grad_logits,
                                          fovea_stream_t stream) {
  FOVEA_REQUIRE(xs && grad_xs && grad_logits && B > 0 && n > 0, "fovea_saliency_softmax_bwd: bad arguments");
  saliency_softmax_bwd_kernel<<<B, kSoftmaxThreads, 0, static_cast<cudaStream_t>(stream)>>>(xs, grad_xs, grad_logits,
                                                                                            n);
  return check_launch("fovea_saliency_softmax_bwd");
}

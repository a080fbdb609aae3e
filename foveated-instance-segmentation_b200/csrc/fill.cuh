// Per-pixel source map, per-triangle setup records and launch parameters shared by the stage-3 fill kernels
// (inverse.cu: forward fill; inverse_bwd.cu: its transpose; mask_fill.cu: the pruned arg-max fill).
#pragma once
#include "common.cuh"

namespace fovea {

// `loc` is stored in 16 bits per pixel (the fill kernel's only per-pixel read stream: halving it is worth 6 % of the
// store bandwidth, measured with fovea_probe_store_ceiling): bit 15 clear = triangle id (< 32768), bit 15 set = direct
// table row n (< 32768).  In-kernel the signed form is used: t >= 0, or -(n+1).
__device__ __forceinline__ uint16_t encode_loc(int v) {
  return static_cast<uint16_t>(v >= 0 ? v : (0x8000 | (-v - 1)));
}
__device__ __forceinline__ int decode_loc(unsigned v) { return (v & 0x8000u) ? -static_cast<int>(v & 0x7FFFu) - 1 : static_cast<int>(v); }

// ---- per-triangle setup records ------------------------------------------------------------------------------
// Everything the walkers and the fill need about a triangle, derived once per triangle from (mesh, pts, src) instead
// of once per visit: one 64-byte record = four independent 16-byte loads, no pts/src indirection.
//   e_i(y,x) = A_i*y + B_i*x + C_i  is the orientation-normalised edge function of the edge OPPOSITE vertex i
//   (> 0 inside; e_0/area, e_1/area are the barycentric coordinates of vertices 0 and 1).  All exact int32 for
//   coordinates < 16384.  Pixel (y,x) belongs to the triangle iff e_i >= m_i for i = 0,1,2, where m_i = 0 if the
//   tie rule of mesh.cuh gives an exactly-on-edge pixel to this triangle (or the edge is on the hull), else 1.
//   q0 = (A0, B0, C0, A1)   q1 = (B1, C1, A2, B2)   q2 = (C2, n0 | n1 << 16, n2 | m << 16, area)
//   q3 = (src0 | src1 << 16, src2, 1/area as a double)            area == 0: degenerate, owns nothing
struct TriRec {
  uint4 q0, q1, q2, q3;
};

struct FillParams {
  int C, Cs, h, w, H, W, cap, tcap, zero_residual;
  int mask_u8;  // 1: the fused argmax is written as uint8 (C <= 256) instead of torch.argmax's int64
};


}  // namespace fovea

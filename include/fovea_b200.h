/*
 * fovea_b200.h -- C ABI of libfovea_b200.so: the foveated resampling hot path of FovealSeg
 * (SAI-Lab-NYU/Foveated-Instance-Segmentation) as hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI of its own: its "operator interface" for this path is the Python surface of
 * models/models.py (DeformSegmentationModule.create_grid, the F.grid_sample call sites,
 * fillMissingValues_tensor) and interp2d.py (Interp2D).  Each entry point below names the reference
 * lines it replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes stubs a
 * maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host;
 *   - all float tensors are fp32, contiguous, NCHW unless the comment gives another layout;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing synchronises;
 *   - the caller owns every buffer; the library keeps no global state except the last-error string;
 *   - return value: 0 = FOVEA_OK, negative = FOVEA_ERR_*; no C++ exception crosses the boundary;
 *   - there is NO CPU fallback: every entry point launches CUDA kernels or fails.
 */
#ifndef FOVEA_B200_H_
#define FOVEA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FOVEA_ABI_VERSION 13

enum fovea_status {
  FOVEA_OK = 0,
  FOVEA_ERR_ARG = -1,      /* bad shape / null pointer / unsupported size */
  FOVEA_ERR_CUDA = -2,     /* a CUDA runtime call or launch failed; see fovea_last_error() */
  FOVEA_ERR_CAPACITY = -3  /* a per-image capacity (points, triangles, shared memory) is exceeded */
};

/* how the saliency map handed to fovea_grid_* relates to the padded map xs_hm of the reference
 * (models/models.py:819-825, cfg.TRAIN.def_saliency_pad_mode) */
enum fovea_pad_mode {
  FOVEA_PAD_NONE = 0,        /* input IS xs_hm [B,gh+2Rx,gw+2Ry], as create_grid() receives it */
  FOVEA_PAD_REPLICATION = 1, /* input is xs [B,gh,gw]; nn.ReplicationPad2d fused into the filter taps */
  FOVEA_PAD_REFLECT = 2,     /* F.pad(mode='reflect') fused */
  FOVEA_PAD_ZERO = 3         /* F.pad(mode='constant') fused */
};

typedef void* fovea_stream_t; /* cudaStream_t */

int fovea_abi_version(void);
/* thread-local, NUL-terminated description of the last failure in this thread ("" if none) */
const char* fovea_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Stage 0 -- either side of the (stock) saliency network.
 * ------------------------------------------------------------------------------------------------ */

/* A0, models/models.py:684-705: the saliency network's input in one launch --
 *   out[b, 0:C]  = b_imresize(img, (HS,WS), 'bilinear')      (F.interpolate, align_corners=False: 4 taps / pixel)
 *   out[b, C]    = out[b, C+1] = ((i - hidx)^2 + (j - widx)^2) / (HS^2 + WS^2),  hidx = focus_point[b,0]*(HS-1),
 *                  widx = focus_point[b,1]*(WS-1)            (focus map, concatenated twice as the reference does)
 *   img [B,C,H,W] fp32, or uint8 when img_u8 != 0 (ToTensor's uint8 -> fp32 / divisor folded into the taps,
 *                 DynamicFocus/e_preprocess_scripts/dataset.py:133-137); may be PINNED HOST memory (only 4*HS*WS
 *                 samples per channel are touched)
 *   focus_point [B,2] fp32 (h,w) in [0,1)     out [B,C+2,HS,WS] fp32 */
int fovea_saliency_input(const void* img, int img_u8, float divisor, const float* focus_point, int B, int C, int H,
                         int W, int HS, int WS, float* out, fovea_stream_t stream);

/* A2, models/models.py:715-723: nn.Softmax over the n = gh*gw saliency logits of each frame (logits, xs: [B,n]);
 * a NaN logit makes the whole frame NaN as in torch (the reference's assert, a host sync, becomes the caller's choice).
 * Backward: grad_logits = xs * (grad_xs - sum(grad_xs * xs)). */
int fovea_saliency_softmax(const float* logits, int B, int n, float* xs, fovea_stream_t stream);
int fovea_saliency_softmax_bwd(const float* xs, const float* grad_xs, int B, int n, float* grad_logits,
                               fovea_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 1 -- saliency -> sampling grid.      Replaces models/models.py:594-637 (create_grid, forward
 * part: three dense (2Rx+1)x(2Ry+1) convs + div + clamp + nn.Upsample + NCHW->NHWC) and, for
 * pad_mode != NONE, the padding at models/models.py:819-825.
 *
 * The dense Gaussian `filter.weight` (models/models.py:510-515) is exactly rank-1; the caller passes its
 * two 1-D factors g1x[2Rx+1] (rows) and g1y[2Ry+1] (cols), filter[a][b] == g1x[a]*g1y[b].
 *
 *   xs        [B, src_h, src_w]   src = (gh,gw) for fused padding, (gh+2Rx, gw+2Ry) for FOVEA_PAD_NONE
 *   grid      [B, out_h, out_w, 2]  (x->width, y->height) in [-1,1]; bilinear(align_corners=False)
 *                                   resize of the raw gh x gw grid, identity when out == (gh,gw)
 *   sums      [B, 3, gh, gw]       den, num_x, num_y  -- saved for fovea_grid_bwd (may be NULL)
 * ------------------------------------------------------------------------------------------------ */
int fovea_grid_fwd(const float* xs, int B, int gh, int gw, int Rx, int Ry, int pad_mode,
                   const float* g1x, const float* g1y, int out_h, int out_w,
                   float* grid, float* sums, fovea_stream_t stream);

/* Backward of fovea_grid_fwd w.r.t. xs.  Replaces the autograd chain conv2d_backward(input) ->
 * div/clamp/Upsample backward -> replication_pad2d_backward of models/models.py:594-637, 821.
 * d(filter.weight) is NOT produced: no optimizer consumes it (train_deform_semantic.py:273-288).
 *   grad_grid [B,out_h,out_w,2]   sums [B,3,gh,gw] from the forward   grad_xs [B,src_h,src_w] (overwritten) */
int fovea_grid_bwd(const float* grad_grid, const float* sums, int B, int gh, int gw, int Rx, int Ry,
                   int pad_mode, const float* g1x, const float* g1y, int out_h, int out_w,
                   float* grad_xs, fovea_stream_t stream);

/* Bilinear (align_corners=False) resize of an NHWC 2-channel grid: the second nn.Upsample that derives
 * grid_y from grid (models/models.py:627-631).  in [B,ih,iw,2] -> out [B,oh,ow,2]. */
int fovea_grid_resize(const float* in, int B, int ih, int iw, int oh, int ow, float* out,
                      fovea_stream_t stream);
/* adjoint of fovea_grid_resize: grad_out [B,oh,ow,2] -> grad_in [B,ih,iw,2] (overwritten) */
int fovea_grid_resize_bwd(const float* grad_out, int B, int ih, int iw, int oh, int ow, float* grad_in,
                          fovea_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 2 -- F.grid_sample(input, grid) with the reference's defaults (mode='bilinear',
 * padding_mode='zeros', align_corners=False).   Replaces models/models.py:865, 880, 909 (and :937
 * when the fused stage-3 kernel is not used).
 *   in [B,C,H,W]   grid [B,h,w,2]   out [B,C,h,w]
 * ------------------------------------------------------------------------------------------------ */
int fovea_grid_sample_fwd(const float* in, const float* grid, int B, int C, int H, int W, int h, int w,
                          float* out, fovea_stream_t stream);

/* The same forward gather from a uint8 image, with the data loader's ToTensor() (uint8 -> fp32 / divisor, divisor = 255;
 * DynamicFocus/e_preprocess_scripts/dataset.py:133-137) folded into the tap loads: bit-identical to sampling the converted
 * fp32 image, at a quarter of the PCIe / HBM bytes.  No backward (the image needs no grad).
 *   in [B,C,H,W] uint8 */
int fovea_grid_sample_fwd_u8(const uint8_t* in, const float* grid, int B, int C, int H, int W, int h, int w,
                             float divisor, float* out, fovea_stream_t stream);

/* Backward (aten grid_sampler_2d_backward).  grad_in may be NULL (input does not require grad: the image
 * and label); when given it must be ZERO-FILLED by the caller and receives a warp-aggregated scatter-add.
 * grad_grid [B,h,w,2] may be NULL. */
int fovea_grid_sample_bwd(const float* grad_out, const float* in, const float* grid, int B, int C, int H,
                          int W, int h, int w, float* grad_in, float* grad_grid, fovea_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 3 -- inverse resampling back to full resolution.
 * ------------------------------------------------------------------------------------------------ */

/* A7: the scatter half of create_grid's inverse part, models/models.py:640-651.  winner[b,v,u] = row-major
 * low-res index i*w+j of the node whose truncated target is pixel (v,u), -1 where no node lands.
 * Duplicate targets: the LARGEST index wins (the reference leaves the winner undefined).
 *   grid [B,h,w,2] -> winner [B,H,W] int32 (fully overwritten) */
int fovea_grid_inv_scatter(const float* grid, int B, int h, int w, int H, int W, int32_t* winner,
                           fovea_stream_t stream);

/* A7: the float canvas create_grid returns, models/models.py:640-655: grid_inv[b,v,u] =
 * (j/w*2-1, i/h*2-1) of the winner, NaN where unfilled.  winner [B,H,W] -> grid_inv [B,H,W,2] */
int fovea_grid_inv_canvas(const int32_t* winner, int B, int h, int w, int H, int W, float* grid_inv,
                          fovea_stream_t stream);

/* A8 at the low-res nodes: table[b, i*w+j, c] = F.grid_sample(pred, grid_inv) evaluated at node (i,j)
 * (= the zero-padded 2x2 box mean of pred, with aten's fp32 arithmetic), models/models.py:935-937.
 * Row h*w of every image is NaN (value of an unfilled image corner, models/models.py:202-209, 268); row h*w+1
 * is all zeros (what a NaN pixel becomes under the residual NaN -> 0 rule, models_instance.py:940).
 *   pred [B,C,h,w] -> table [B, h*w+2, Cs] fp32, Cs = channel stride >= C, multiple of 4 (tail zero) */
int fovea_box4_table(const float* pred, int B, int C, int h, int w, int Cs, float* table,
                     fovea_stream_t stream);

/* A9 point selection: getPixelsForInterp of fillMissingValues_tensor, models/models.py:169-211, applied to
 * the NaN pattern `winner < 0`: a filled pixel is an interpolation point iff the 3x3-cross dilation of the
 * invalid mask (at <=512 px directly, else on the nearest-downscaled copy, :183-193) covers it; the four
 * image corners are always points.  Points are emitted in row-major order (torch.where order, :265).
 *   nchan     C of the tensor the reference would dilate (enters max(C,H,W)/512, models.py:183)
 *   grid [B,h,w,2] the sampling grid the winners were scattered from
 *   pts  [B,cap] int32  (row<<16 | col)      src [B,cap] int32 row into `table` (h*w = NaN corner)
 *   npts [B]     int32                       h*w+4 <= cap <= 16384
 * fovea_select_points_nb: the site rule of interp_mode 'nearest' / 'BI' instead (getPixelsForInterp_NB,
 * models/models.py:213-242: cv2.dilate reads the [C,H,W] array as rows=C, cols=H, channels=W, so a filled pixel is a
 * site iff the pixel directly above or below it is unfilled; no forced corners).  With these sites the 'tri' machinery
 * (Delaunay + barycentric fill, NaN outside the hull) is rev_deform_interp='BI': scipy's LinearNDInterpolator over the
 * (class, row, col) voxels (:248-250, 269-272) restricted to a class plane is the 2-D Delaunay interpolant of that plane
 * (DESIGN.md section 4). */
int fovea_select_points(const float* grid, const int32_t* winner, int B, int h, int w, int H, int W, int nchan,
                        int cap, int32_t* pts, int32_t* src, int32_t* npts, fovea_stream_t stream);
int fovea_select_points_nb(const float* grid, const int32_t* winner, int B, int h, int w, int H, int W, int nchan,
                           int cap, int32_t* pts, int32_t* src, int32_t* npts, fovea_stream_t stream);
/* fovea_select_points WITHOUT the dense winner map of fovea_grid_inv_scatter (A7 + A9 in one launch; nothing of 4*H*W
 * bytes per frame is cleared, scattered into or probed): the frame's node targets are sorted in shared memory by
 * (pixel, node), the last node of a run of equal pixels wins the pixel (the scatter's atomicMax), "is this pixel filled"
 * is a binary search.  Same pts / src / npts, bit for bit.  Also written:
 *   targets [B, h*w + 4] int32: (row<<16 | col) of node n if it won its pixel, else -1; entries h*w .. h*w+3 name image
 *                               corners no node landed on (fovea_locate_raster_targets stamps the node pixels from it) */
int fovea_select_points_sparse(const float* grid, int B, int h, int w, int H, int W, int nchan, int cap, int32_t* pts,
                               int32_t* src, int32_t* npts, int32_t* targets, fovea_stream_t stream);

/* Triangle mesh layout shared by the entry points below:
 *   mesh [B,tcap,8] uint16, one 16-byte record per triangle: (v0, v1, v2, 0, n0, n1, n2, 0)
 *     v* = indices into the image's point list (either orientation); n_k = triangle across the edge opposite
 *     vertex k (SciPy's `neighbors` convention, spatial/qhull.pyx), 0xFFFF = convex-hull edge.
 *   ntri [B] int32 = triangles per image;  tcap >= 2*cap. */

/* Delaunay triangulation of each image's points ON THE DEVICE (one CTA per image, mesh in shared memory, exact
 * int64 in-circle predicates).  Replaces the host Qhull call interp2d.py:55 (spatial/qhull.pyx:1679).  The caller
 * may instead fill mesh/ntri from a host triangulation (parity mode: stock SciPy Qhull, what the reference does).
 *   pts must be sorted ascending (row-major) and unique, as fovea_select_points emits them;
 *   max_coord = max(H, W) <= 8192;  cap <= 8190 and small enough for 227 KB of shared memory (~6.6k points),
 *   else FOVEA_ERR_CAPACITY (use the host triangulation);
 *   workspace: fovea_delaunay_workspace_bytes(B, cap) bytes; receives the flip-round count per image (int32). */
int64_t fovea_delaunay_workspace_bytes(int B, int cap);
int fovea_delaunay(const int32_t* pts, const int32_t* npts, int B, int cap, int tcap, int max_coord,
                   uint16_t* mesh, int32_t* ntri, void* workspace, fovea_stream_t stream);

/* fovea_delaunay + fovea_locate_hints in ONE launch: the hints are computed while the mesh is still in the kernel's
 * shared memory (saves re-staging 230 KB per image).  Possible when fovea_delaunay_hints_fused(tcap,H,W) != 0 (the two
 * coarse hint levels must fit the kernel's scratch: canvases up to ~3500^2 at the default tcap); otherwise call the
 * two entry points separately.  hints as fovea_locate_hints; workspace as fovea_delaunay. */
int fovea_delaunay_hints_fused(int tcap, int H, int W);
int fovea_delaunay_with_hints(const int32_t* pts, const int32_t* npts, int B, int cap, int tcap, int H, int W,
                              uint16_t* mesh, int32_t* ntri, int32_t* hints, void* workspace, fovea_stream_t stream);

/* Walk-start hints for fovea_locate_pixels: hints[b, cy, cx] = the triangle containing the centre of the
 * FOVEA_HINT_CELL_W x FOVEA_HINT_CELL_H pixel cell (one thread of the locate kernel walks one cell-wide row run).
 * hints [B, ceil(H/FOVEA_HINT_CELL_H), ceil(W/FOVEA_HINT_CELL_W)] int32
 * workspace: fovea_locate_hints_workspace_bytes(B, H, W) bytes of device memory */
#define FOVEA_HINT_CELL_W 32
#define FOVEA_HINT_CELL_H 8
int64_t fovea_locate_hints_workspace_bytes(int B, int H, int W);
int fovea_locate_hints(const int32_t* pts, const int32_t* npts, const uint16_t* mesh, const int32_t* ntri, int B,
                       int cap, int tcap, int H, int W, int32_t* hints, void* workspace, fovea_stream_t stream);

/* Per-triangle setup records: everything the point location and the fill need about a triangle, derived once per
 * triangle from (mesh, pts, src): the three orientation-normalised edge functions e_i(y,x) = A_i*y + B_i*x + C_i
 * (exact int32), the tie-ownership bit of each edge (top-left rule, csrc/mesh.cuh), the neighbours, |area|, 1/area
 * (float64, as interp2d.py:58 / spatial/qhull.pyx:1210-1264 compute the barycentric transform) and the value-table rows
 * of the three vertices.  Layout: csrc/inverse.cu `struct TriRec` (64 bytes).
 * Tie ownership: a pixel exactly on an edge goes to the triangle that can produce a value when exactly one of the two
 * has a vertex without one (src == nan_row, an image corner no node landed on); else to the top-left rule.
 *   trirec [B, tcap, 16] int32 (64-byte records);  max_coord = max(H, W) <= 16384;  nan_row = h*w (row of NaN). */
#define FOVEA_TRIREC_BYTES 64
int fovea_triangle_setup(const int32_t* pts, const int32_t* src, const uint16_t* mesh, const int32_t* ntri, int B,
                         int cap, int tcap, int max_coord, int nan_row, void* trirec, fovea_stream_t stream);

/* A9 point location, interp2d.py:58 (Delaunay.find_simplex over every pixel, spatial/qhull.pyx:2075-2163), merged
 * with the A7 winners into one per-pixel source map -- a function of the sampling grid only (not of the scores):
 * 16 bits per pixel (this map is the fill kernel's only per-pixel read stream):
 *   bit 15 clear: loc = id of the mesh triangle that owns pixel (y,x).  A pixel exactly on an edge belongs to the
 *                 triangle a top-left fill rule picks (csrc/mesh.cuh), so the map does not depend on walk order;
 *   bit 15 set:   loc & 0x7FFF = n: the pixel received low-res node n directly (winner[b,y,x] = n);  n = h*w: the
 *                 pixel has no value (outside the triangulation / empty mesh) and reads the NaN row of the value table.
 *   winner [B,H,W] int32 from fovea_grid_inv_scatter (all -1 = interpolate every pixel, Interp2D);  trirec from
 *   fovea_triangle_setup;  hints from fovea_locate_hints;  loc [B,H,W] uint16;  H, W <= 16384, tcap <= 32768,
 *   h*w < 32767. */
int fovea_locate_pixels(const int32_t* winner, const void* trirec, const int32_t* ntri, const int32_t* hints, int B,
                        int h, int w, int H, int W, int tcap, uint16_t* loc, fovea_stream_t stream);

/* The same per-pixel source map as fovea_locate_pixels, computed by RASTERISING the mesh from the setup records' exact
 * integer edge functions -- the predicate fovea_locate_pixels tests pixel by pixel, so the two maps are identical --
 * and then stamping the pixels that received a node from the nodes themselves.  Needs no walk-start hints and does not
 * read the winner map except at the 6 400 node targets.
 *   On a canvas the triangulation covers (prefill == 0) and with a workspace: one lane per triangle walks its rows with an
 *   exact integer DDA and marks where each (triangle, row) span STARTS; one sweep per row then carries the last started
 *   id forward with coalesced 16-byte accesses.  Otherwise (or with FOVEA_RAS_MODE=0): every triangle's bounding box is
 *   swept pixel by pixel, the few huge hull triangles row span by row span.
 *   grid, winner : the sampling grid [B,h,w,2] and fovea_grid_inv_scatter's map; both NULL when no pixel carries a node
 *                  (Interp2D on an arbitrary point set)
 *   prefill != 0 : first mark every pixel "no value" -- required when the triangulation does not cover the canvas (no
 *                  forced corners: 'BI' sites, arbitrary point sets);  W must be a multiple of 8
 *   workspace    : fovea_locate_raster_workspace_bytes(B, H, W, tcap) bytes (the queue of tall triangles), or NULL */
int64_t fovea_locate_raster_workspace_bytes(int B, int H, int W, int tcap);
/* The same with the node pixels stamped from fovea_select_points_sparse's `targets` [B, h*w + 4] instead of the sampling
 * grid + the dense winner map (a canvas the triangulation covers: 'tri' sites). */
int fovea_locate_raster_targets(const int32_t* pts, const uint16_t* mesh, const void* trirec, const int32_t* ntri,
                                const int32_t* targets, int B, int h, int w, int H, int W, int cap, int tcap,
                                uint16_t* loc, void* workspace, fovea_stream_t stream);
int fovea_locate_raster(const int32_t* pts, const uint16_t* mesh, const void* trirec, const int32_t* ntri,
                        const float* grid, const int32_t* winner, int B, int h, int w, int H, int W, int cap, int tcap,
                        int prefill, uint16_t* loc, void* workspace, fovea_stream_t stream);

/* A8 + A9 + A10 fused: F.grid_sample(pred, grid_inv) + NaN mask (models/models.py:935-938), the per-sample
 * fillMissingValues_tensor(..., 'tri') = Interp2D barycentric gather (models/models.py:939-940,
 * interp2d.py:65-91), residual NaN -> 0 (models_instance.py:940) and torch.argmax over classes
 * (models/models.py:1044), in ONE pass over the full-resolution canvas: the score tensor is written exactly once.
 *   loc    [B,H,W] from fovea_locate_pixels
 *   trirec [B,tcap,16] from fovea_triangle_setup
 *   table  [B, h*w+2, Cs] from fovea_box4_table   (h*w+2 <= 32768)
 *   scores [B,C,H,W] fp32   (NULL = do not materialise)
 *   mask   [B,H,W]  int64 (torch.argmax's dtype), or uint8 when mask_u8 != 0 (C <= 256)   (NULL = do not compute)
 *          = torch.argmax over classes for FINITE predictions (first maximum wins; a pixel without any value -- NaN in
 *          every channel -- gives 0 like torch).  Only if `pred` itself contains NaN / Inf can a pixel be NaN in SOME
 *          channels; the fused argmax then skips those channels where torch would return the first NaN's index: run
 *          fovea_argmax_classes on the materialised scores for that case.
 *   zero_residual: 1 = NaN -> 0 before writing / argmax */
int fovea_inverse_fill(const uint16_t* loc, const void* trirec, const float* table, int B, int C, int Cs, int h, int w,
                       int H, int W, int tcap, int zero_residual, float* scores, void* mask, int mask_u8,
                       fovea_stream_t stream);

/* Mask mode of stage 3: torch.argmax(pred_sampled, dim=1) (models/models.py:1044) of the tensor fovea_inverse_fill would
 * write, WITHOUT interpolating all C channels -- bit-identical to fovea_inverse_fill(scores = NULL, mask) including ties.
 * Per node: the argmax of its table row (a pixel that received a node carries that row).  Per triangle: only the channels
 * that are not dominated at all three vertices can win a pixel (the interpolation weights are >= 0 and IEEE rounding is
 * monotone); they are evaluated with the arithmetic of fovea_inverse_fill, every other channel is skipped (csrc/mask_fill.cu
 * states the tie / rounding argument).  Writes 8 (int64) or 1 (mask_u8) bytes per pixel.
 *   ntri      [B] triangles per frame (NULL for a fovea_nearest_locate map, which holds no triangle ids)
 *   workspace fovea_inverse_mask_workspace_bytes(B,h,w,tcap) bytes, 16-byte aligned;   C <= 256 */
int64_t fovea_inverse_mask_workspace_bytes(int B, int h, int w, int tcap);
int fovea_inverse_mask(const uint16_t* loc, const void* trirec, const int32_t* ntri, const float* table, int B, int C, int Cs,
                       int h, int w, int H, int W, int tcap, void* workspace, void* mask, int mask_u8,
                       fovea_stream_t stream);

/* Backward of fovea_inverse_fill w.r.t. its value table: the autograd graph the reference builds when it trains through
 * the inverse path (models/models.py:933-940, MODEL.upsample / MODEL.loss_at_high_res) and the gradient Interp2D
 * promises for `values` (interp2d.py:38-47):   grad_table[b, row, c] = sum over pixels p and vertices k with
 * row_k(p) == row of  w_k(p) * grad_scores[b, c, p].   Pixels the forward leaves NaN / zero send nothing.
 * grad_table [B, h*w+2, Cs] is zeroed by the call and accumulated with atomics (the order of the adds, and so the
 * last bits, can differ from run to run, exactly like ATen's grid_sampler backward).  The transpose of
 * fovea_box4_table (table -> pred) is fovea_grid_sample_bwd at the node coordinates. */
int fovea_inverse_fill_bwd(const uint16_t* loc, const void* trirec, const float* grad_scores, int B, int C, int Cs, int h,
                           int w, int H, int W, int tcap, float* grad_table, fovea_stream_t stream);

/* rev_deform_interp = 'nearest' (the mode config/deform.yaml:17 ships): fillMissingValues_tensor(..., 'nearest'),
 * models/models.py:213-250, 259-272 = getPixelsForInterp_NB + scipy NearestNDInterpolator on the host.  Produces the
 * same per-pixel source map as fovea_locate_pixels, every entry a direct table row:
 *   loc[b,y,x] = 0x8000 | n   n = the node the pixel received (winner >= 0), else the node of the NEAREST interpolation
 *                         site (exact integer Euclidean distance; equidistant sites: the leftmost column, then the
 *                         upper site);  n = h*w when the image has no site at all (NaN row).
 * Sites: filled pixels with an unfilled pixel directly above/below (on the nearest-downscaled mask when
 * max(nchan,H,W) > 512, models.py:222-232); no forced corners.  Feed `loc` to fovea_inverse_fill (trirec is not read).
 *   workspace: fovea_nearest_workspace_bytes(B,H,W) bytes;  H, W < 32767. */
int64_t fovea_nearest_workspace_bytes(int B, int H, int W);
int fovea_nearest_locate(const int32_t* winner, int B, int h, int w, int H, int W, int nchan, void* workspace,
                         uint16_t* loc, fovea_stream_t stream);

/* DynamicFocus deformed_unsampler (DynamicFocus/d_model/nn_B0_deformed_sampler.py:115-153):
 * scatter the low-resolution labels at integer pixel targets, then give every other pixel the value of its nearest
 * scattered pixel (the reference: scipy.ndimage.distance_transform_edt(return_indices=True) on the HOST).
 *   fovea_scatter_nodes      : coords [B,2,h,w] int64 (row, column; out-of-canvas targets are dropped) -> winner [B,H,W]
 *                              (largest node index wins a shared pixel, -1 = none), :127-137
 *   fovea_nearest_locate_all : fovea_nearest_locate with EVERY filled pixel a site (exact Euclidean nearest; ties: leftmost
 *                              column, then upper site), :143-149;  same workspace, same `loc`
 *   fovea_node_table         : table[b][n][c] = values[b][c][n] (+ the NaN and zero rows), the plain counterpart of
 *                              fovea_box4_table;  feed loc + table to fovea_inverse_fill (scores mode) */
int fovea_scatter_nodes(const int64_t* coords, int B, int h, int w, int H, int W, int32_t* winner,
                        fovea_stream_t stream);
int fovea_nearest_locate_all(const int32_t* winner, int B, int h, int w, int H, int W, void* workspace, uint16_t* loc,
                             fovea_stream_t stream);
int fovea_node_table(const float* values, int B, int C, int h, int w, int Cs, float* table, fovea_stream_t stream);

/* Diagnostic (bench.py): the store pattern of fovea_inverse_fill with no computation -- same tiling, one 128-bit
 * streaming store per 4 pixels and channel plane.  Its GB/s is the practical write-only ceiling of this layout.
 * side_read (may be NULL): a buffer of >= 2*B*H*W bytes read once, 8 bytes per thread, like the fill kernel's 16-bit
 * `loc` map -- shows what a 1 % read stream mixed into the write stream costs at the DRAM. */
int fovea_probe_store_ceiling(float* scores, const int32_t* side_read, int B, int C, int H, int W,
                              fovea_stream_t stream);

/* torch.argmax(scores, dim=1) as a stand-alone pass (models/models.py:1044): first maximum wins, NaN is
 * treated as the maximum (torch semantics).  scores [B,C,H,W] -> mask [B,H,W] int64 */
int fovea_argmax_classes(const float* scores, int B, int C, int64_t HW, int64_t* mask, fovea_stream_t stream);

/* SURVEY.md section 8f row 4 (C1 decoder tail, models/model_utils.py:277-309): widen the uint8 labels of a
 * reduced-channel fill to the reference's int64 class ids, mask[b,p] = lut[b][labels[b,p]].
 *   labels [B,HW] uint8   lut [B,nl] int64   mask [B,HW] int64   (HW % 4 == 0, nl <= 256; labels >= nl clamp to nl-1) */
int fovea_relabel_mask(const uint8_t* labels, const int64_t* lut, int B, int64_t HW, int nl, int64_t* mask,
                       fovea_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FOVEA_B200_H_ */

"""Torch-facing operators of the foveated resampling path, each a thin call into libfovea_b200.so.

Every function takes/returns CUDA tensors, checks dtype/contiguity/device in Python (the C ABI sees only raw
pointers and sizes) and launches on torch's current stream.  Nothing here computes on the CPU except the optional
host triangulation of `build_inverse_plan(..., triangulation="host")`, which is what the reference itself does
(interp2d.py:53-58 runs Qhull on the host).
"""
from __future__ import annotations

import ctypes as C
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import FoveaError, PAD_MODES


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    """torch's current stream on the current device; `_req` refuses tensors that live on another device."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req_out(t, dtype, name, shape):
    """An output buffer supplied by the caller: written through its raw pointer, so it must BE the expected buffer --
    CUDA, right dtype, exact shape, contiguous (a .contiguous() copy would silently receive the result instead)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise FoveaError(f"{name}: expected a CUDA tensor")
    dtypes = dtype if isinstance(dtype, tuple) else (dtype,)
    if t.dtype not in dtypes:
        raise FoveaError(f"{name}: expected dtype {' or '.join(str(d) for d in dtypes)}, got {t.dtype}")
    if tuple(t.shape) != tuple(shape):
        raise FoveaError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    if not t.is_contiguous():
        raise FoveaError(f"{name}: must be contiguous")
    return t


def _req(t, dtype, name, ndim=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise FoveaError(f"{name}: expected a CUDA tensor (this path has no CPU fallback), got "
                         f"{type(t).__name__}{'' if not isinstance(t, torch.Tensor) else ' on ' + str(t.device)}")
    if t.dtype != dtype:
        raise FoveaError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise FoveaError(f"{name}: expected {ndim} dims, got shape {tuple(t.shape)}")
    if t.device.index != torch.cuda.current_device():    # kernels launch on the CURRENT device's current stream
        raise FoveaError(f"{name}: tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                         f"call inside `with torch.cuda.device({t.device.index})`")
    return t.contiguous()


# ------------------------------------------------------------------------------------------------ stage 0

def saliency_input(img, focus_point, input_size, divisor=255.0):
    """models/models.py:684-705 in one launch: `cat(b_imresize(img, input_size, 'bilinear'), focus, focus)`.

    img: [B,C,H,W] fp32 or uint8 (uint8: ToTensor's `/ divisor` is folded into the taps), on the GPU or in pinned
    host memory; focus_point: [B,2] (h,w) in [0,1).  Returns x_low [B,C+2,HS,WS] fp32.  Not differentiable (neither
    the image nor the gaze needs a gradient in the reference)."""
    if not isinstance(img, torch.Tensor) or img.dtype not in (torch.float32, torch.uint8):
        raise FoveaError("saliency_input: img must be a float32 or uint8 tensor")
    x = _req_gather_source(img, "img", img.dtype)
    dev = x.device if x.is_cuda else focus_point.device
    fp = _req(focus_point.detach().to(torch.float32), torch.float32, "focus_point", 2)
    B, Cc, H, W = x.shape
    if fp.shape[0] != B or fp.shape[1] != 2:
        raise FoveaError(f"saliency_input: focus_point {tuple(fp.shape)} does not match img {tuple(x.shape)}")
    HS, WS = int(input_size[0]), int(input_size[1])
    out = torch.empty(B, Cc + 2, HS, WS, device=dev, dtype=torch.float32)
    _lib.call("fovea_saliency_input", _ptr(x), int(x.dtype == torch.uint8), float(divisor), _ptr(fp), B, Cc, H, W, HS,
              WS, _ptr(out), _stream())
    return out


class _SaliencySoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits):
        z = _req(logits, torch.float32, "logits", 2)
        xs = torch.empty_like(z)
        _lib.call("fovea_saliency_softmax", _ptr(z), z.shape[0], z.shape[1], _ptr(xs), _stream())
        ctx.save_for_backward(xs)
        return xs

    @staticmethod
    def backward(ctx, grad_xs):
        xs, = ctx.saved_tensors
        g = _req(grad_xs, torch.float32, "grad_xs", 2)
        out = torch.empty_like(xs)
        _lib.call("fovea_saliency_softmax_bwd", _ptr(xs), _ptr(g), xs.shape[0], xs.shape[1], _ptr(out), _stream())
        return out


def saliency_softmax(logits):
    """nn.Softmax over each frame's saliency logits, models/models.py:715-723.  logits: [B,...] -> same shape, every
    frame sums to 1; differentiable."""
    shape = logits.shape
    return _SaliencySoftmaxFn.apply(logits.reshape(shape[0], -1)).view(shape)


# ------------------------------------------------------------------------------------------------ stage 1

def separable_factors(filter_weight: torch.Tensor, rtol: float = 1e-5):
    """Split the dense Gaussian `filter.weight[0,0]` ([2Rx+1, 2Ry+1], models/models.py:510-515) into its two 1-D
    factors g1x (rows) and g1y (cols).  Raises if the weight is not rank-1 (e.g. it was trained away from the
    Gaussian): the sm_100a kernel is a separable filter and will not silently approximate."""
    w = filter_weight.detach().reshape(filter_weight.shape[-2], filter_weight.shape[-1]).double().cpu()
    Kx, Ky = w.shape
    Rx, Ry = Kx // 2, Ky // 2
    centre = w[Rx, Ry]
    if centre == 0:
        raise FoveaError("filter.weight centre is zero; cannot factor")
    g1y = w[Rx, :].clone()
    g1x = w[:, Ry] / centre
    err = (torch.outer(g1x, g1y) - w).abs().max() / w.abs().max()
    if err > rtol:
        raise FoveaError(f"filter.weight is not rank-1 (relative residual {err:.2e}); the separable kernel "
                         f"requires the Gaussian built from cfg.MODEL.gaussian_radius")
    return g1x.float(), g1y.float()


class _GridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xs, g1x, g1y, gh, gw, Rx, Ry, pad_mode, out_h, out_w):
        B = xs.shape[0]
        xs_c = _req(xs, torch.float32, "xs").view(B, xs.shape[-2], xs.shape[-1])
        grid = torch.empty(B, out_h, out_w, 2, device=xs.device, dtype=torch.float32)
        sums = torch.empty(B, 3, gh, gw, device=xs.device, dtype=torch.float32)
        _lib.call("fovea_grid_fwd", _ptr(xs_c), B, gh, gw, Rx, Ry, pad_mode, _ptr(g1x), _ptr(g1y), out_h, out_w,
                  _ptr(grid), _ptr(sums), _stream())
        ctx.save_for_backward(sums, g1x, g1y)
        ctx.geom = (gh, gw, Rx, Ry, pad_mode, out_h, out_w, tuple(xs.shape))
        return grid

    @staticmethod
    def backward(ctx, grad_grid):
        sums, g1x, g1y = ctx.saved_tensors
        gh, gw, Rx, Ry, pad_mode, out_h, out_w, shape = ctx.geom
        B = shape[0]
        gg = _req(grad_grid, torch.float32, "grad_grid")
        grad_xs = torch.empty(shape, device=gg.device, dtype=torch.float32)
        _lib.call("fovea_grid_bwd", _ptr(gg), _ptr(sums), B, gh, gw, Rx, Ry, pad_mode, _ptr(g1x), _ptr(g1y), out_h,
                  out_w, _ptr(grad_xs), _stream())
        return (grad_xs,) + (None,) * 9


def saliency_to_grid(xs, g1x, g1y, gh, gw, Rx, Ry, pad_mode, out_size):
    """Stage 1 (models/models.py:594-637 [+ :819-825 when pad_mode != 'none']).

    xs: [B,1,gh,gw] normalised saliency (fused padding) or the padded xs_hm [B,1,gh+2Rx,gw+2Ry] (pad_mode='none').
    Returns grid [B,out_h,out_w,2]; differentiable w.r.t. xs.
    """
    mode = PAD_MODES[pad_mode] if isinstance(pad_mode, str) else int(pad_mode)
    eh, ew = (gh + 2 * Rx, gw + 2 * Ry) if mode == _lib.PAD_NONE else (gh, gw)
    if xs.dim() != 4 or xs.shape[1] != 1 or tuple(xs.shape[-2:]) != (eh, ew):
        raise FoveaError(f"saliency_to_grid: expected xs of shape [B,1,{eh},{ew}], got {tuple(xs.shape)}")
    g1x = _req(g1x, torch.float32, "g1x")
    g1y = _req(g1y, torch.float32, "g1y")
    if g1x.numel() != 2 * Rx + 1 or g1y.numel() != 2 * Ry + 1:
        raise FoveaError("saliency_to_grid: filter factor lengths do not match Rx/Ry")
    return _GridFn.apply(xs, g1x, g1y, gh, gw, Rx, Ry, mode, int(out_size[0]), int(out_size[1]))


class _GridResizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grid, oh, ow):
        g = _req(grid, torch.float32, "grid", 4)
        B, ih, iw, _ = g.shape
        out = torch.empty(B, oh, ow, 2, device=g.device, dtype=torch.float32)
        _lib.call("fovea_grid_resize", _ptr(g), B, ih, iw, oh, ow, _ptr(out), _stream())
        ctx.geom = (B, ih, iw, oh, ow)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        B, ih, iw, oh, ow = ctx.geom
        go = _req(grad_out, torch.float32, "grad_out")
        gi = torch.empty(B, ih, iw, 2, device=go.device, dtype=torch.float32)
        _lib.call("fovea_grid_resize_bwd", _ptr(go), B, ih, iw, oh, ow, _ptr(gi), _stream())
        return gi, None, None


def grid_resize(grid, out_size):
    """nn.Upsample(size, 'bilinear') of an NHWC grid (models/models.py:627-631); identity sizes alias."""
    oh, ow = int(out_size[0]), int(out_size[1])
    if grid.shape[1] == oh and grid.shape[2] == ow:
        return grid
    return _GridResizeFn.apply(grid, oh, ow)


# ------------------------------------------------------------------------------------------------ stage 2

def _req_gather_source(t, name, dtype=torch.float32):
    """The image a gather kernel reads: a CUDA tensor, or a PINNED host tensor (unified virtual addressing makes its
    data_ptr dereferenceable on the device: the kernel then pulls only the 32-byte sectors it touches over PCIe)."""
    if isinstance(t, torch.Tensor) and not t.is_cuda and t.is_pinned():
        if t.dtype != dtype or t.dim() != 4 or not t.is_contiguous():
            raise FoveaError(f"{name}: a pinned host source must be a contiguous 4-D {dtype} tensor")
        return t
    return _req(t, dtype, name, 4)


def grid_sample_u8(inp, grid, divisor=255.0):
    """F.grid_sample(inp.float() / divisor, grid) for a uint8 image [B,C,H,W] (CUDA or pinned host): ToTensor() folded
    into the tap loads; bit-identical to converting first.  Not differentiable w.r.t. the
    image; use grid_sample() on the fp32 image when the grid needs a gradient."""
    x = _req_gather_source(inp, "input", torch.uint8)
    g = _req(grid.detach(), torch.float32, "grid", 4)
    B, Cc, H, W = x.shape
    if g.shape[0] != B or g.shape[3] != 2:
        raise FoveaError(f"grid_sample_u8: grid {tuple(g.shape)} does not match input {tuple(x.shape)}")
    h, w = g.shape[1], g.shape[2]
    out = torch.empty(B, Cc, h, w, device=g.device, dtype=torch.float32)
    _lib.call("fovea_grid_sample_fwd_u8", _ptr(x), _ptr(g), B, Cc, H, W, h, w, float(divisor), _ptr(out), _stream())
    return out


class _GridSampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, grid):
        x = _req_gather_source(inp, "input")
        g = _req(grid, torch.float32, "grid", 4)
        B, Cc, H, W = x.shape
        if g.shape[0] != B or g.shape[3] != 2:
            raise FoveaError(f"grid_sample: grid {tuple(g.shape)} does not match input {tuple(x.shape)}")
        h, w = g.shape[1], g.shape[2]
        out = torch.empty(B, Cc, h, w, device=g.device, dtype=torch.float32)
        _lib.call("fovea_grid_sample_fwd", _ptr(x), _ptr(g), B, Cc, H, W, h, w, _ptr(out), _stream())
        ctx.save_for_backward(x, g)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, g = ctx.saved_tensors
        B, Cc, H, W = x.shape
        h, w = g.shape[1], g.shape[2]
        go = _req(grad_out, torch.float32, "grad_out")
        need_in, need_grid = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if need_in and not x.is_cuda:
            raise FoveaError("grid_sample backward: a pinned host source cannot receive a gradient")
        gi = torch.zeros_like(x) if need_in else None
        gg = torch.empty_like(g) if need_grid else None
        if need_in or need_grid:
            _lib.call("fovea_grid_sample_bwd", _ptr(go), _ptr(x), _ptr(g), B, Cc, H, W, h, w, _ptr(gi), _ptr(gg),
                      _stream())
        return gi, gg


def grid_sample(inp, grid):
    """F.grid_sample(inp, grid) with the reference's defaults (models/models.py:865, 880, 909, 937)."""
    return _GridSampleFn.apply(inp, grid)


# ------------------------------------------------------------------------------------------------ stage 3

def grid_inv_scatter(grid, segSize):
    """models/models.py:640-651 -> winner int32 [B,H,W] (row-major low-res index, -1 unfilled)."""
    g = _req(grid.detach(), torch.float32, "grid", 4)
    B, h, w, _ = g.shape
    H, W = int(segSize[0]), int(segSize[1])
    winner = torch.empty(B, H, W, device=g.device, dtype=torch.int32)
    _lib.call("fovea_grid_inv_scatter", _ptr(g), B, h, w, H, W, _ptr(winner), _stream())
    return winner


def grid_inv_canvas(winner, h, w):
    """models/models.py:652-655 -> grid_inv float [B,H,W,2], NaN where unfilled."""
    win = _req(winner, torch.int32, "winner", 3)
    B, H, W = win.shape
    out = torch.empty(B, H, W, 2, device=win.device, dtype=torch.float32)
    _lib.call("fovea_grid_inv_canvas", _ptr(win), B, h, w, H, W, _ptr(out), _stream())
    return out


@dataclass
class InversePlan:
    """Everything stage 3 needs that depends only on the sampling grid (not on the scores being warped)."""
    winner: torch.Tensor   # [B,H,W] int32 (None for a plan built with dense_winner=False)
    pts: torch.Tensor      # [B,cap] int32 (row<<16|col), row-major sorted
    src: torch.Tensor      # [B,cap] int32 row of the value table
    npts: torch.Tensor     # [B] int32
    mesh: torch.Tensor     # [B,tcap,8] uint16: (v0,v1,v2,0,n0,n1,n2,0) per triangle
    ntri: torch.Tensor     # [B] int32
    hints: torch.Tensor    # [B,ceil(H/8),ceil(W/32)] int32 walk-start triangles
    trirec: torch.Tensor   # [B,tcap,16] int32: 64-byte per-triangle setup records (edge functions, 1/area, rows)
    loc: torch.Tensor      # [B,H,W] uint16 per-pixel source: triangle id, or 0x8000 | table row
    h: int
    w: int
    H: int
    W: int
    cap: int
    tcap: int
    triangulation: str
    rounds: torch.Tensor = None   # [B] int32 flip rounds of the device Delaunay kernel; -1 = did not converge (check_plan)


def _host_delaunay_one(pts_packed: np.ndarray):
    """Stock SciPy Qhull on one image's points, exactly as interp2d.py:55 (default options + Qt)."""
    from scipy.spatial import Delaunay
    rc = np.stack([pts_packed >> 16, pts_packed & 0xFFFF], 1).astype(np.float64)
    tri = Delaunay(rc)
    return tri.simplices.astype(np.int64), tri.neighbors.astype(np.int64)


def _triangulate_host(pts, npts, cap, tcap, pool=None):
    B = pts.shape[0]
    pts_h = pts.cpu().numpy()
    npts_h = npts.cpu().numpy()
    mesh = np.zeros((B, tcap, 8), dtype=np.uint16)
    ntri = np.zeros(B, dtype=np.int32)
    jobs = [pts_h[b, : npts_h[b]] for b in range(B)]
    if pool is None and B > 1:
        with ThreadPoolExecutor(max_workers=min(B, 32)) as ex:
            res = list(ex.map(_host_delaunay_one, jobs))
    elif pool is not None:
        res = list(pool.map(_host_delaunay_one, jobs))
    else:
        res = [_host_delaunay_one(jobs[0])]
    for b, (simp, nb) in enumerate(res):
        T = simp.shape[0]
        if T > tcap:
            raise FoveaError(f"host triangulation produced {T} triangles > tcap={tcap}")
        mesh[b, :T, 0:3] = simp
        mesh[b, :T, 4:7] = np.where(nb < 0, 0xFFFF, nb)
        ntri[b] = T
    dev = pts.device
    return torch.from_numpy(mesh).to(dev), torch.from_numpy(ntri).to(dev)


def build_inverse_plan(grid, segSize, nchan, triangulation="host", pool=None, sites="tri", dense_winner=True) -> InversePlan:
    """A7 scatter + A9 point selection + triangulation + walk hints for a batch of sampling grids.

    triangulation='host'  : stock SciPy Qhull on the host, exactly what the reference does (parity mode);
    triangulation='device': the sm_100a Delaunay kernel (fast mode; differs from Qhull only in how co-circular
                            point sets are split).
    sites='tri': interpolation sites of rev_deform_interp='tri' (getPixelsForInterp, models/models.py:169-211);
    sites='nb' : the sites of 'BI' / 'nearest' (getPixelsForInterp_NB, :213-242; no forced corners: pixels outside the
                 sites' convex hull stay NaN, as scipy's LinearNDInterpolator leaves them).
    dense_winner=False (sites='tri' on the raster path): never materialise the A7 winner map [B,H,W] int32 -- the
                 sites and the node stamps come from the sorted node targets (fovea_select_points_sparse); same plan,
                 bit for bit, with plan.winner = None.  The pipelines use it; the reference-facing mirrors keep the map.
    """
    if sites not in ("tri", "nb"):
        raise FoveaError(f"unknown site rule {sites!r}")
    g = _req(grid.detach(), torch.float32, "grid", 4)
    B, h, w, _ = g.shape
    H, W = int(segSize[0]), int(segSize[1])
    dev = g.device
    cap = h * w + 4
    tcap = 2 * cap
    pts = torch.empty(B, cap, device=dev, dtype=torch.int32)
    src = torch.empty(B, cap, device=dev, dtype=torch.int32)
    npts = torch.empty(B, device=dev, dtype=torch.int32)
    raster = _use_raster(W)
    sparse = not dense_winner and sites == "tri" and raster
    winner = targets = None
    if sparse:
        targets = torch.empty(B, cap, device=dev, dtype=torch.int32)
        _lib.call("fovea_select_points_sparse", _ptr(g), B, h, w, H, W, int(nchan), cap, _ptr(pts), _ptr(src),
                  _ptr(npts), _ptr(targets), _stream())
    else:
        winner = grid_inv_scatter(g, (H, W))
        _lib.call("fovea_select_points" if sites == "tri" else "fovea_select_points_nb", _ptr(g), _ptr(winner), B, h, w,
                  H, W, int(nchan), cap, _ptr(pts), _ptr(src), _ptr(npts), _stream())
    hints = rounds = None
    if triangulation == "host":
        mesh, ntri = _triangulate_host(pts, npts, cap, tcap, pool)
    elif triangulation == "device":
        if not raster and _lib.load().fovea_delaunay_hints_fused(tcap, H, W):
            mesh, ntri, hints, ws = delaunay_device_with_hints(pts, npts, cap, tcap, H, W)
        else:
            mesh, ntri, ws = delaunay_device(pts, npts, cap, tcap, max(H, W))
        rounds = ws[:B]
    else:
        raise FoveaError(f"unknown triangulation mode {triangulation!r}")
    trirec = _triangle_setup(pts, src, mesh, ntri, cap, tcap, max(H, W), h * w)
    if sparse:
        loc = _locate_raster_targets(pts, mesh, trirec, ntri, targets, h, w, H, W, cap, tcap)
    elif raster:   # sites='nb' has no forced corners: the hull does not cover the canvas, start from "no value"
        loc = _locate_raster(pts, mesh, trirec, ntri, g, winner, h, w, cap, tcap, prefill=sites != "tri")
    else:
        if hints is None:
            hints = _hints(pts, npts, mesh, ntri, B, cap, tcap, H, W)
        loc = _locate(winner, trirec, ntri, hints, h, w, tcap)
    return InversePlan(winner, pts, src, npts, mesh, ntri, hints, trirec, loc, h, w, H, W, cap, tcap, triangulation,
                       rounds)


def _use_raster(W):
    """fovea_locate_raster (one warp per triangle) unless FOVEA_LOCATE=walk asks for the scan-line walker
    (fovea_locate_hints + fovea_locate_pixels) or the canvas width is not a multiple of 8 (its 16-byte span stores)."""
    import os
    return os.environ.get("FOVEA_LOCATE", "raster") != "walk" and W % 8 == 0


def _locate_raster_targets(pts, mesh, trirec, ntri, targets, h, w, H, W, cap, tcap):
    """_locate_raster with the node pixels stamped from fovea_select_points_sparse's targets (no winner map)."""
    B = pts.shape[0]
    loc = torch.empty(B, H, W, device=pts.device, dtype=torch.int16)       # uint16 bit patterns
    nbytes = int(_lib.load().fovea_locate_raster_workspace_bytes(B, H, W, tcap))
    ws = torch.empty((nbytes + 3) // 4, device=pts.device, dtype=torch.int32)
    _lib.call("fovea_locate_raster_targets", _ptr(pts), _ptr(mesh), _ptr(trirec), _ptr(ntri), _ptr(targets), B, h, w, H, W,
              cap, tcap, _ptr(loc), _ptr(ws), _stream())
    return loc


def _locate_raster(pts, mesh, trirec, ntri, grid, winner, h, w, cap, tcap, prefill):
    """interp2d.py:58 (find_simplex for every pixel) by rasterising the mesh, merged with the A7 winners (grid may be
    None: no pixel carries a node): the per-pixel source map `loc`."""
    B, H, W = winner.shape
    if grid is not None:
        grid = _req(grid, torch.float32, "grid", 4)                        # (the kernel reads it through the raw pointer)
    loc = torch.empty(B, H, W, device=winner.device, dtype=torch.int16)    # uint16 bit patterns
    ws = None
    if not prefill:                                                        # the span-start bitmap of the marker raster
        nbytes = int(_lib.load().fovea_locate_raster_workspace_bytes(B, H, W, tcap))
        ws = torch.empty((nbytes + 3) // 4, device=winner.device, dtype=torch.int32)
    _lib.call("fovea_locate_raster", _ptr(pts), _ptr(mesh), _ptr(trirec), _ptr(ntri), _ptr(grid),
              _ptr(winner if grid is not None else None), B, h, w, H, W, cap, tcap, 1 if prefill else 0, _ptr(loc),
              _ptr(ws), _stream())
    return loc


def check_plan(plan: InversePlan):
    """Raise FoveaError if the device Delaunay kernel reported a frame whose flip loop hit its safety bound (it then
    emits NO mesh for that frame rather than a non-Delaunay one).  Reads a [B] int32 tensor back: one host sync -- the
    reference-facing mirrors call it after the fill has been enqueued; pipelines call it when they drain."""
    if plan.rounds is not None:
        bad = torch.nonzero(plan.rounds < 0).flatten().tolist()
        if bad:
            raise FoveaError(f"device Delaunay did not converge for frame(s) {bad}; use triangulation='host'")
    return plan


def delaunay_device(pts, npts, cap, tcap, max_coord):
    """interp2d.py:55 on the device: Delaunay-triangulate every image's sorted, unique packed points.
    Returns (mesh [B,tcap,8] uint16, ntri [B] int32, flip_rounds [B] int32)."""
    pts = _req(pts, torch.int32, "pts", 2)
    npts = _req(npts, torch.int32, "npts", 1)
    B, dev = pts.shape[0], pts.device
    mesh = torch.empty(B, tcap, 8, device=dev, dtype=torch.uint16)
    ntri = torch.empty(B, device=dev, dtype=torch.int32)
    nbytes = int(_lib.load().fovea_delaunay_workspace_bytes(B, cap))
    ws = torch.zeros((nbytes + 3) // 4, device=dev, dtype=torch.int32)  # [B] flip rounds, [B,8] counters, scratch
    _lib.call("fovea_delaunay", _ptr(pts), _ptr(npts), B, cap, tcap, int(max_coord), _ptr(mesh), _ptr(ntri),
              _ptr(ws), _stream())
    return mesh, ntri, ws


def delaunay_device_with_hints(pts, npts, cap, tcap, H, W):
    """delaunay_device + the walk-start hints in one launch (the mesh is still in the kernel's shared memory)."""
    B, dev = pts.shape[0], pts.device
    mesh = torch.empty(B, tcap, 8, device=dev, dtype=torch.uint16)
    ntri = torch.empty(B, device=dev, dtype=torch.int32)
    hints = torch.empty(B, -(-H // _lib.HINT_CELL_H), -(-W // _lib.HINT_CELL_W), device=dev, dtype=torch.int32)
    nbytes = int(_lib.load().fovea_delaunay_workspace_bytes(B, cap))
    ws = torch.zeros((nbytes + 3) // 4, device=dev, dtype=torch.int32)
    _lib.call("fovea_delaunay_with_hints", _ptr(pts), _ptr(npts), B, cap, tcap, H, W, _ptr(mesh), _ptr(ntri),
              _ptr(hints), _ptr(ws), _stream())
    return mesh, ntri, hints, ws


def _hints(pts, npts, mesh, ntri, B, cap, tcap, H, W):
    dev = pts.device
    ch, cw = -(-H // _lib.HINT_CELL_H), -(-W // _lib.HINT_CELL_W)
    hints = torch.empty(B, ch, cw, device=dev, dtype=torch.int32)
    hws = torch.empty(int(_lib.load().fovea_locate_hints_workspace_bytes(B, H, W)) // 4 + 1, device=dev,
                      dtype=torch.int32)
    _lib.call("fovea_locate_hints", _ptr(pts), _ptr(npts), _ptr(mesh), _ptr(ntri), B, cap, tcap, H, W, _ptr(hints),
              _ptr(hws), _stream())
    return hints


def _triangle_setup(pts, src, mesh, ntri, cap, tcap, max_coord, nan_row):
    """Per-triangle setup records (edge functions, tie bits, 1/area, value rows) for the walkers and the fill."""
    B = pts.shape[0]
    trirec = torch.empty(B, tcap, 16, device=pts.device, dtype=torch.int32)
    _lib.call("fovea_triangle_setup", _ptr(pts), _ptr(src), _ptr(mesh), _ptr(ntri), B, cap, tcap, int(max_coord),
              int(nan_row), _ptr(trirec), _stream())
    return trirec


def _locate(winner, trirec, ntri, hints, h, w, tcap):
    """interp2d.py:58 (find_simplex for every pixel) merged with the A7 winners: the per-pixel source map `loc`."""
    B, H, W = winner.shape
    if W % 4:
        raise FoveaError(f"canvas width {W} must be a multiple of 4 (128-bit accesses)")
    loc = torch.empty(B, H, W, device=winner.device, dtype=torch.int16)    # uint16 bit patterns
    _lib.call("fovea_locate_pixels", _ptr(winner), _ptr(trirec), _ptr(ntri), _ptr(hints), B, h, w, H, W, tcap,
              _ptr(loc), _stream())
    return loc


def nearest_locate(winner, h, w, nchan):
    """rev_deform_interp='nearest' (models/models.py:213-250, 259-272): per-pixel source map from the winner map alone --
    every unfilled pixel points at the table row of its nearest interpolation site (exact integer distances)."""
    win = _req(winner, torch.int32, "winner", 3)
    B, H, W = win.shape
    nbytes = int(_lib.load().fovea_nearest_workspace_bytes(B, H, W))
    ws = torch.empty((nbytes + 3) // 4, device=win.device, dtype=torch.int32)
    loc = torch.empty(B, H, W, device=win.device, dtype=torch.int16)    # uint16 bit patterns
    _lib.call("fovea_nearest_locate", _ptr(win), B, int(h), int(w), H, W, int(nchan), _ptr(ws), _ptr(loc), _stream())
    return loc


def scatter_nodes(coords, segSize):
    """DynamicFocus deformed_unsampler's scatter (nn_B0_deformed_sampler.py:127-137): coords [B,2,h,w] int64 (row, column)
    -> winner int32 [B,H,W]; the largest node index wins a shared pixel, targets outside the canvas are dropped."""
    c = _req(coords, torch.int64, "coords", 4)
    B, two, h, w = c.shape
    if two != 2:
        raise FoveaError(f"scatter_nodes: expected coords [B,2,h,w], got {tuple(c.shape)}")
    H, W = int(segSize[0]), int(segSize[1])
    winner = torch.empty(B, H, W, device=c.device, dtype=torch.int32)
    _lib.call("fovea_scatter_nodes", _ptr(c), B, h, w, H, W, _ptr(winner), _stream())
    return winner


def nearest_locate_all(winner, h, w):
    """Exact nearest-filled-pixel labelling (a Euclidean distance transform with indices, nn_B0_deformed_sampler.py:143-149):
    like nearest_locate, but EVERY filled pixel is a site."""
    win = _req(winner, torch.int32, "winner", 3)
    B, H, W = win.shape
    nbytes = int(_lib.load().fovea_nearest_workspace_bytes(B, H, W))
    ws = torch.empty((nbytes + 3) // 4, device=win.device, dtype=torch.int32)
    loc = torch.empty(B, H, W, device=win.device, dtype=torch.int16)
    _lib.call("fovea_nearest_locate_all", _ptr(win), B, int(h), int(w), H, W, _ptr(ws), _ptr(loc), _stream())
    return loc


def node_table(values, Cs=None):
    """[B,C,h,w] -> channel-contiguous value table [B, h*w+2, Cs] of the nodes themselves (no 2x2 box mean)."""
    v = _req(values.detach(), torch.float32, "values", 4)
    B, Cc, h, w = v.shape
    Cs = Cs or (Cc + 3) // 4 * 4
    table = torch.empty(B, h * w + 2, Cs, device=v.device, dtype=torch.float32)
    _lib.call("fovea_node_table", _ptr(v), B, Cc, h, w, Cs, _ptr(table), _stream())
    return table


def build_nearest_plan(grid, segSize, nchan) -> InversePlan:
    """A7 scatter + nearest-site labelling: the 'nearest' counterpart of build_inverse_plan (no triangulation)."""
    g = _req(grid.detach(), torch.float32, "grid", 4)
    B, h, w, _ = g.shape
    H, W = int(segSize[0]), int(segSize[1])
    winner = grid_inv_scatter(g, (H, W))
    loc = nearest_locate(winner, h, w, nchan)
    dummy = torch.zeros(1, 1, 16, device=g.device, dtype=torch.int32)   # no triangle ids in `loc`: never read
    return InversePlan(winner, None, None, None, None, None, None, dummy, loc, h, w, H, W, h * w + 4, 1, "nearest")


def plan_from_mesh(pts, src, npts, mesh, ntri, H, W, table_rows) -> InversePlan:
    """Plan for interpolating EVERY pixel of an H x W canvas from an arbitrary site set (Interp2D, interp2d.py:37-91):
    no pixel carries a node (winner = -1 everywhere); `table_rows` is the index of the NaN row of the value table."""
    B, cap = pts.shape
    tcap = mesh.shape[1]
    winner = torch.full((B, H, W), -1, device=pts.device, dtype=torch.int32)
    trirec = _triangle_setup(pts, src, mesh, ntri, cap, tcap, max(H, W), table_rows)
    hints = None
    if _use_raster(W):
        loc = _locate_raster(pts, mesh, trirec, ntri, None, winner, table_rows, 1, cap, tcap, prefill=True)
    else:
        hints = _hints(pts, npts, mesh, ntri, B, cap, tcap, H, W)
        loc = _locate(winner, trirec, ntri, hints, table_rows, 1, tcap)
    return InversePlan(winner, pts, src, npts, mesh, ntri, hints, trirec, loc, table_rows, 1, H, W, cap, tcap, "given")


def inverse_fill_table(plan: InversePlan, table, C, zero_residual=False, scores=None, mask=None):
    """fovea_inverse_fill on a caller-built value table [B, plan.h*plan.w + 2, Cs].  Differentiable w.r.t. `table` when it
    requires grad and scores are produced (interp2d.py:38-47: gradients flow to `values`)."""
    if scores is not None and torch.is_grad_enabled() and table.requires_grad:
        return _FillFn.apply(table, plan, int(C), bool(zero_residual), scores, mask), mask
    _fill(plan, table.detach(), int(C), zero_residual, scores, mask)
    return scores, mask


_FULL_MASK_FILL = False   # tests / A-B runs: True forces the all-channel fused argmax in mask mode


def _fill(plan, table, C, zero_residual, scores, mask):
    mask_u8 = 0
    B = plan.loc.shape[0]
    if scores is not None:
        _req_out(scores, torch.float32, "inverse_fill: scores", (B, C, plan.H, plan.W))
    if mask is not None:   # int64 = torch.argmax's dtype
        _req_out(mask, (torch.int64, torch.uint8), "inverse_fill: mask", (B, plan.H, plan.W))
        mask_u8 = 1 if mask.dtype == torch.uint8 else 0
    if table.dim() != 3 or table.shape[0] != B or table.shape[1] != plan.h * plan.w + 2 or table.shape[2] < C:
        raise FoveaError(f"inverse_fill: value table {tuple(table.shape)} does not match the plan "
                         f"([{B}, {plan.h * plan.w + 2}, >= {C}])")
    table = _req(table, torch.float32, "inverse_fill: table", 3)
    if scores is None and mask is not None and C <= 256 and not _FULL_MASK_FILL:
        # mask mode: the pruned arg-max fill (csrc/mask_fill.cu) -- bit-identical to the fused argmax below
        nbytes = int(_lib.load().fovea_inverse_mask_workspace_bytes(B, plan.h, plan.w, plan.tcap))
        ws = torch.empty(nbytes, device=table.device, dtype=torch.uint8)
        _lib.call("fovea_inverse_mask", _ptr(plan.loc), _ptr(plan.trirec), _ptr(plan.ntri), _ptr(table), B, C,
                  table.shape[2], plan.h, plan.w, plan.H, plan.W, plan.tcap, _ptr(ws), _ptr(mask), mask_u8, _stream())
        return
    _lib.call("fovea_inverse_fill", _ptr(plan.loc), _ptr(plan.trirec), _ptr(table), plan.loc.shape[0], C,
              table.shape[2], plan.h, plan.w, plan.H, plan.W, plan.tcap, 1 if zero_residual else 0, _ptr(scores),
              _ptr(mask), mask_u8, _stream())


class _FillFn(torch.autograd.Function):
    """fovea_inverse_fill with the gradient of the scores w.r.t. the value table (fovea_inverse_fill_bwd): the autograd
    edge the reference gets from torch.gather / mul / sum in interp2d.py:76-89 and from F.grid_sample(pred, grid_inv)
    at the pixels that received a node (models/models.py:937).  The argmax mask (if any) is written as a side effect."""

    @staticmethod
    def forward(ctx, table, plan, C, zero_residual, scores, mask):
        _fill(plan, table.detach(), C, zero_residual, scores, mask)
        ctx.plan, ctx.C, ctx.tshape = plan, C, tuple(table.shape)
        ctx.mark_dirty(scores)
        return scores

    @staticmethod
    def backward(ctx, grad_scores):
        plan, C = ctx.plan, ctx.C
        g = _req(grad_scores, torch.float32, "grad_scores", 4)
        B, rows, Cs = ctx.tshape
        gtable = torch.empty(B, rows, Cs, device=g.device, dtype=torch.float32)
        _lib.call("fovea_inverse_fill_bwd", _ptr(plan.loc), _ptr(plan.trirec), _ptr(g), B, C, Cs, plan.h, plan.w, plan.H,
                  plan.W, plan.tcap, _ptr(gtable), _stream())
        return gtable, None, None, None, None, None


def _node_grid(B, h, w, device):
    """The coordinates the reference stores in grid_inv for node (i,j) (models/models.py:652-653): x = j/w*2-1, y = i/h*2-1,
    fp32 op for op -- fovea_box4_table samples `pred` exactly there."""
    gx = torch.arange(w, device=device, dtype=torch.float32) / float(w) * 2.0 - 1.0
    gy = torch.arange(h, device=device, dtype=torch.float32) / float(h) * 2.0 - 1.0
    g = torch.stack([gx[None, :].expand(h, w), gy[:, None].expand(h, w)], dim=-1)
    return g[None].expand(B, h, w, 2).contiguous()


class _Box4TableFn(torch.autograd.Function):
    """fovea_box4_table with its transpose: the table is F.grid_sample(pred, node grid) laid out channel-last, so the
    gradient w.r.t. pred is fovea_grid_sample_bwd's scatter-add at the node coordinates."""

    @staticmethod
    def forward(ctx, pred, Cs):
        table = box4_table(pred, Cs)
        ctx.pshape = tuple(pred.shape)
        return table

    @staticmethod
    def backward(ctx, gtable):
        B, Cc, h, w = ctx.pshape
        go = gtable[:, :h * w, :Cc].permute(0, 2, 1).reshape(B, Cc, h, w).contiguous()
        grid = _node_grid(B, h, w, go.device)
        gi = torch.zeros(B, Cc, h, w, device=go.device, dtype=torch.float32)
        # grad_input only: the input values themselves are not read (they only enter d/d grid)
        _lib.call("fovea_grid_sample_bwd", _ptr(go), _ptr(gi), _ptr(grid), B, Cc, h, w, h, w, _ptr(gi), None, _stream())
        return gi, None


def box4_table(pred, Cs=None):
    """A8 at the nodes: [B, h*w+2, Cs] value table (row h*w NaN, row h*w+1 zeros), models/models.py:935-937."""
    p = _req(pred.detach(), torch.float32, "pred", 4)
    B, Cc, h, w = p.shape
    Cs = Cs or (Cc + 7) // 8 * 8        # multiple of 8: the fill kernel reads rows with 256-bit loads
    table = torch.empty(B, h * w + 2, Cs, device=p.device, dtype=torch.float32)
    _lib.call("fovea_box4_table", _ptr(p), B, Cc, h, w, Cs, _ptr(table), _stream())
    return table


def inverse_fill(plan: InversePlan, pred, want_scores=True, want_mask=False, zero_residual=True, out=None,
                 mask_out=None):
    """A8 + A9 (+A10): full-resolution scores [B,C,H,W] and/or argmax mask [B,H,W] int64 from pred [B,C,h,w].
    Differentiable w.r.t. `pred` (scores only) when it requires grad: models/models.py:933-940."""
    p = _req(pred, torch.float32, "pred", 4)
    B, Cc, h, w = p.shape
    if (h, w) != (plan.h, plan.w) or B != plan.loc.shape[0]:
        raise FoveaError(f"inverse_fill: pred {tuple(p.shape)} does not match the plan ({B}x{plan.h}x{plan.w})")
    scores = None
    if want_scores:
        scores = out if out is not None else torch.empty(B, Cc, plan.H, plan.W, device=p.device, dtype=torch.float32)
        _req_out(scores, torch.float32, "inverse_fill: out", (B, Cc, plan.H, plan.W))
    mask = None
    if want_mask:
        mask = mask_out if mask_out is not None else torch.empty(B, plan.H, plan.W, device=p.device, dtype=torch.int64)
        _req_out(mask, (torch.int64, torch.uint8), "inverse_fill: mask_out", (B, plan.H, plan.W))
    if want_scores and torch.is_grad_enabled() and p.requires_grad:
        table = _Box4TableFn.apply(p, None)
        return _FillFn.apply(table, plan, Cc, bool(zero_residual), scores, mask), mask
    _fill(plan, box4_table(p), Cc, zero_residual, scores, mask)
    return scores, mask


def c1_tail_pred(cls_pred, x):
    """The C1 decoder's tail as the reference materialises it (models/model_utils.py:298-309): class logits broadcast
    over the map, the last one multiplied by the mask logit x = sigmoid(.) - 0.5.  -> pred [B,K,h,w]."""
    B, K = cls_pred.shape
    pred = cls_pred[:, :, None, None].expand(B, K, x.shape[-2], x.shape[-1]).clone()
    pred[:, -1:] = cls_pred[:, -1:, None, None] * x
    return pred


def inverse_mask_c1(plan: InversePlan, cls_pred, x, mask_out=None):
    """Instance mask of the C1 decoder without materialising its [B,K,h,w] prediction or the [B,K,H,W] scores
    (SURVEY.md section 8f row 4): `argmax(inverse_fill(plan, c1_tail_pred(cls_pred, x)))`.

    cls_pred [B,K] class logits, x [B,1,h,w] mask logit.  Only ONE of the K channels varies over the map; the first
    K-1 are per-frame constants c_k, and the whole inverse path is linear with non-negative weights, so their
    full-resolution scores are c_k * s(p) with one common s(p) > 0: their arg-max is the same class k* at every pixel
    (first maximum, as torch.argmax).  Stage 3 therefore runs on three channels -- a sentinel, k*, and the varying
    channel -- writing 1 byte per pixel, and fovea_relabel_mask widens {0,1,2} to the reference's int64 ids {0, k*,
    K-1} (0 = a pixel the reference leaves NaN / sets to zero in every channel: torch.argmax gives class 0 there).
    Stage-3 traffic drops from 4*K*H*W to 9*H*W bytes.  Differs from the general path only at exact ties."""
    cp = _req(cls_pred.detach(), torch.float32, "cls_pred", 2)
    xm = _req(x.detach(), torch.float32, "x", 4)
    B, K = cp.shape
    if K < 2 or xm.shape[0] != B or xm.shape[1] != 1 or tuple(xm.shape[-2:]) != (plan.h, plan.w):
        raise FoveaError(f"inverse_mask_c1: cls_pred {tuple(cp.shape)} / x {tuple(xm.shape)} do not match the plan "
                         f"({plan.loc.shape[0]}x{plan.h}x{plan.w})")
    kstar = torch.argmax(cp[:, :K - 1], dim=1)
    pred3 = torch.empty(B, 3, plan.h, plan.w, device=cp.device, dtype=torch.float32)
    pred3[:, 0] = -1e30                                            # sentinel: never the maximum of a valid pixel
    pred3[:, 1] = cp.gather(1, kstar[:, None])[:, :, None]
    pred3[:, 2] = cp[:, K - 1, None, None] * xm[:, 0]
    table = box4_table(pred3, Cs=4)
    labels = torch.empty(B, plan.H, plan.W, device=cp.device, dtype=torch.uint8)
    _fill(plan, table, 3, True, None, labels)
    lut = torch.stack([torch.zeros_like(kstar), kstar, torch.full_like(kstar, K - 1)], dim=1).contiguous()
    mask = mask_out if mask_out is not None else torch.empty(B, plan.H, plan.W, device=cp.device, dtype=torch.int64)
    if mask.dtype != torch.int64 or not mask.is_cuda or not mask.is_contiguous():
        raise FoveaError("inverse_mask_c1: mask_out must be a contiguous CUDA int64 tensor")
    _lib.call("fovea_relabel_mask", _ptr(labels), _ptr(lut), B, plan.H * plan.W, 3, _ptr(mask), _stream())
    return mask


def probe_store_ceiling(scores, side_read=None):
    """Diagnostic: overwrite `scores` [B,C,H,W] with the store pattern of fovea_inverse_fill and no computation;
    `side_read` (any int32 buffer of >= B*H*W/2 elements) adds the fill kernel's 2-byte-per-pixel `loc` read stream."""
    s = _req(scores, torch.float32, "scores", 4)
    B, Cc, H, W = s.shape
    if side_read is not None:
        side_read = _req(side_read, torch.int32, "side_read", 3)     # only its first 2*B*H*W bytes are read
    _lib.call("fovea_probe_store_ceiling", _ptr(s), _ptr(side_read), B, Cc, H, W, _stream())


def argmax_classes(scores):
    """torch.argmax(scores, dim=1) (models/models.py:1044) as one streaming pass."""
    s = _req(scores, torch.float32, "scores", 4)
    B, Cc, H, W = s.shape
    mask = torch.empty(B, H, W, device=s.device, dtype=torch.int64)
    _lib.call("fovea_argmax_classes", _ptr(s), B, Cc, H * W, _ptr(mask), _stream())
    return mask

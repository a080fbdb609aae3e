"""CPU oracle for the foveated resampling path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement (torch CPU ops + NumPy + stock SciPy) of the reference's algorithm for SURVEY.md
section 8 rows A0..A10.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this file; the product package (`foveated-instance-segmentation_b200/`)
never does, and fails loudly when its CUDA library is missing.

Parity status: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: `tests/golden/make_golden.py` imports the
unmodified `/root/reference/models/models.py` + `interp2d.py` through `tests/golden/ref_shim.py` and stores
their outputs on seeded inputs in `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function
below against them.  The Delaunay triangulation (`delaunay()` below, stock SciPy Qhull) is additionally pinned
against the reference's OWN vendored Qhull 2019.1, compiled from `/root/reference/spatial/qhull_src` into
`oracle/_ref/libqhull_ref.so` (oracle/Makefile, oracle/qhull_ref.py): identical simplices, orientation, order
and neighbours on every point set of the path (`tests/test_oracle_qhull_ref.py`).  What stays unpinned: the
fork's Cython `find_simplex(return_c=True)` (`spatial/qhull.pyx:2075-2163`) cannot be built; stock SciPy's
`find_simplex` + `transform` stand in for it (same algorithm and eps).

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# A3: Gaussian filter + P_basis                                   models/models.py:140-157, 510-522
# ----------------------------------------------------------------------------------------------


def make_gaussian(size: int, fwhm: float) -> np.ndarray:
    """models/models.py:140-157 -- un-normalised square Gaussian, centre size//2, float64."""
    ax = np.arange(0, size, 1, float)
    c = size // 2
    return np.exp(-4.0 * np.log(2.0) * ((ax[None, :] - c) ** 2 + (ax[:, None] - c) ** 2) / fwhm ** 2)


def gaussian_filter_weight(Rx: int, Ry: int, fwhm: float) -> torch.Tensor:
    """models/models.py:510-515 -- fp32 weight [2Rx+1, 2Ry+1] of the 1->1 channel `filter` conv."""
    g = torch.FloatTensor(make_gaussian(2 * Rx + 1, fwhm))
    g = F.interpolate(g[None, None], (2 * Rx + 1, 2 * Ry + 1), mode="bilinear")  # dataset.py:30-31
    return g[0, 0].contiguous()


def p_basis(gh: int, gw: int, Rx: int, Ry: int) -> torch.Tensor:
    """models/models.py:517-522 -- P[0,i,j]=(j-Ry)/(gw-1), P[1,i,j]=(i-Rx)/(gh-1), fp32 [2,gh+2Rx,gw+2Ry].

    The reference fills a float32 tensor element by element from Python floats (double arithmetic,
    rounded once on store); reproduced by computing in float64 and casting.
    """
    i = np.arange(gh + 2 * Rx, dtype=np.float64)[:, None]
    j = np.arange(gw + 2 * Ry, dtype=np.float64)[None, :]
    P = np.empty((2, gh + 2 * Rx, gw + 2 * Ry), dtype=np.float64)
    P[0] = (j - Ry) / (gw - 1.0) + 0.0 * i
    P[1] = (i - Rx) / (gh - 1.0) + 0.0 * j
    return torch.from_numpy(P.astype(np.float32))


# ----------------------------------------------------------------------------------------------
# A0: gaze focus map + low-res saliency-net input                        models/models.py:684-705
# ----------------------------------------------------------------------------------------------


def focus_map(focus_point: torch.Tensor, HS: int, WS: int) -> torch.Tensor:
    """models/models.py:684-694 (+ torch_tools.py:65-69) -> [B,1,HS,WS] fp32."""
    max_dist = np.sqrt(HS ** 2 + WS ** 2)
    hidx = focus_point[:, 0] * (HS - 1)
    widx = focus_point[:, 1] * (WS - 1)
    ii = torch.arange(HS)[:, None].repeat(1, WS)
    jj = torch.arange(WS)[None, :].repeat(HS, 1)
    dist = torch.sqrt((ii[None] - hidx[:, None, None]) ** 2 + (jj[None] - widx[:, None, None]) ** 2)
    return (dist / max_dist).unsqueeze(1) ** 2


def saliency_input(x: torch.Tensor, focus_point: torch.Tensor, sal_size) -> torch.Tensor:
    """models/models.py:701-705 -- bilinear x_low (dataset.py:30-31) + the focus map twice -> [B,5,HS,WS]."""
    x_low = F.interpolate(x, tuple(sal_size), mode="bilinear")
    fm = focus_map(focus_point, sal_size[0], sal_size[1]).to(x_low.dtype)
    return torch.cat((x_low, fm, fm), dim=1)


# ----------------------------------------------------------------------------------------------
# A2: saliency normalisation + padding                               models/models.py:715-723, 819-825
# ----------------------------------------------------------------------------------------------


def saliency_normalise(xs_logits: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
    """models/models.py:715-723 -- Upsample((gh,gw),'bilinear') -> softmax over gh*gw -> [B,1,gh,gw]."""
    xs = F.interpolate(xs_logits, (gh, gw), mode="bilinear")
    xs = torch.softmax(xs.reshape(-1, gh * gw), dim=1)
    return xs.view(-1, 1, gh, gw)


def pad_saliency(xs: torch.Tensor, Rx: int, Ry: int, mode: str = "replication") -> torch.Tensor:
    """models/models.py:819-825 -- (left,right,top,bottom) = (Ry,Ry,Rx,Rx)."""
    pad = (Ry, Ry, Rx, Rx)
    if mode == "replication":
        return F.pad(xs, pad, mode="replicate")
    if mode == "reflect":
        return F.pad(xs, pad, mode="reflect")
    if mode == "zero":
        return F.pad(xs, pad, mode="constant")
    raise ValueError(mode)


# ----------------------------------------------------------------------------------------------
# A4 + A7: create_grid                                                    models/models.py:594-657
# ----------------------------------------------------------------------------------------------


def create_grid(xs_hm, filt_w, P, gh, gw, task_size, task_size_eval=(), rate=1, segSize=None, x_inv=None,
                tie="max"):
    """models/models.py:594-657.  `filt_w` [Kx,Ky] fp32, `P` [2,G,G] fp32, xs_hm [B,1,gh+2Rx,gw+2Ry].

    Returns (grid[B,h,w,2], grid_y) or, when segSize and x_inv are given, (grid, grid_inv[B,Hs,Ws,2] with NaN).
    """
    B = xs_hm.shape[0]
    w4 = filt_w[None, None]
    Pb = P[None].expand(B, -1, -1, -1)
    den = F.conv2d(xs_hm, w4)                                                   # :602
    num = F.conv2d((Pb * torch.cat((xs_hm, xs_hm), 1)).reshape(-1, 1, *xs_hm.shape[-2:]), w4)  # :603-604
    num = num.view(-1, 2, gh, gw)
    gx = torch.clamp(num[:, 0:1] / den * 2 - 1, min=-1, max=1)                  # :609-615
    gy = torch.clamp(num[:, 1:2] / den * 2 - 1, min=-1, max=1)
    grid = torch.cat((gx, gy), 1)                                              # :619
    infer_size = tuple(task_size_eval) if len(task_size_eval) != 0 else tuple(task_size)
    if len(task_size_eval) != 0 and segSize is not None:                       # :621-625
        grid = F.interpolate(grid, infer_size, mode="bilinear")
    else:
        grid = F.interpolate(grid, tuple(task_size), mode="bilinear")
    if segSize is None:                                                        # :627-631
        grid_y = F.interpolate(grid, tuple(int(s) // rate for s in task_size), mode="bilinear")
    else:
        grid_y = F.interpolate(grid, infer_size, mode="bilinear")
    grid = grid.permute(0, 2, 3, 1)                                            # :633-637
    grid_y = grid_y.permute(0, 2, 3, 1)
    if segSize is not None and x_inv is not None:
        return grid, grid_inverse(grid, segSize, tie=tie)
    return grid, grid_y


def grid_inverse(grid: torch.Tensor, segSize, tie: str = "max") -> torch.Tensor:
    """models/models.py:640-655 -- scatter low-res (col,row) indices into a NaN canvas at truncated targets.

    Duplicate targets (several low-res nodes truncating to one pixel): the reference's `index_put_` leaves the
    winner UNDEFINED (nondeterministic on CUDA; on CPU it follows TensorIterator's internal traversal order,
    which is neither first nor last -- see tests/test_oracle_golden.py).  `tie="torch"` replays the reference's
    own ops (bit-identical to the golden on this torch build); `tie="max"` is the deterministic rule the CUDA
    path adopts: the largest row-major low-res index wins.  Both agree wherever there is no collision.
    """
    B, h, w, _ = grid.shape
    Hs, Ws = int(segSize[0]), int(segSize[1])
    if tie == "max":
        win = grid_inverse_winner(grid, segSize)
        filled = win >= 0
        xc = (win % w).float()
        yc = torch.div(win, w, rounding_mode="floor").float()
        nan = torch.full_like(xc, float("nan"))
        inv0 = torch.where(filled, xc / w * 2 - 1, nan)
        inv1 = torch.where(filled, yc / h * 2 - 1, nan)
        return torch.stack((inv0, inv1), -1)
    g = grid.permute(3, 0, 1, 2)
    inv = torch.full((2, B, Hs, Ws), float("nan"), dtype=grid.dtype)
    u = (((g[0] + 1) / 2) * (Ws - 1)).int().long().view(B, -1)
    v = (((g[1] + 1) / 2) * (Hs - 1)).int().long().view(B, -1)
    xc = torch.arange(w).unsqueeze(0).expand(h, w).reshape(-1).unsqueeze(0).expand(B, -1).float()
    yc = torch.arange(h).unsqueeze(-1).expand(h, w).reshape(-1).unsqueeze(0).expand(B, -1).float()
    bidx = torch.arange(B).unsqueeze(-1)
    inv[0][bidx, v, u] = xc
    inv[1][bidx, v, u] = yc
    inv[0] = inv[0] / w * 2 - 1
    inv[1] = inv[1] / h * 2 - 1
    return inv.permute(1, 2, 3, 0)


def grid_inverse_winner(grid: torch.Tensor, segSize) -> torch.Tensor:
    """Integer restatement of `grid_inverse`: winner low-res linear index i*w+j per target pixel, -1 = unfilled.

    Same truncation arithmetic as models/models.py:644-645, tie rule 'largest index wins' (explicit max
    instead of relying on index_put order).  int64 [B,Hs,Ws].
    """
    B, h, w, _ = grid.shape
    Hs, Ws = int(segSize[0]), int(segSize[1])
    u = (((grid[..., 0] + 1) / 2) * (Ws - 1)).int().long().view(B, -1)
    v = (((grid[..., 1] + 1) / 2) * (Hs - 1)).int().long().view(B, -1)
    win = torch.full((B, Hs * Ws), -1, dtype=torch.int64)
    src = torch.arange(h * w).unsqueeze(0).expand(B, -1)
    win.scatter_reduce_(1, v * Ws + u, src, reduce="amax", include_self=True)
    return win.view(B, Hs, Ws)


# ----------------------------------------------------------------------------------------------
# A5 / A6: grid_sample                                        models/models.py:865, 880, 909, 937
# ----------------------------------------------------------------------------------------------


def grid_sample(inp: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """The reference calls F.grid_sample with defaults (bilinear, zeros, align_corners=False)."""
    return F.grid_sample(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


def grid_sample_explicit(inp: np.ndarray, grid: np.ndarray) -> np.ndarray:
    """Independent float64 NumPy restatement of the same op used to cross-check."""
    B, C, H, W = inp.shape
    _, h, w, _ = grid.shape
    inp = inp.astype(np.float64)
    ix = ((grid[..., 0].astype(np.float64) + 1.0) * W - 1.0) / 2.0
    iy = ((grid[..., 1].astype(np.float64) + 1.0) * H - 1.0) / 2.0
    x0 = np.floor(ix).astype(np.int64)
    y0 = np.floor(iy).astype(np.int64)
    tx = ix - x0
    ty = iy - y0
    out = np.zeros((B, C, h, w), dtype=np.float64)
    bb = np.arange(B)[:, None, None]
    for dy, dx, wt in ((0, 0, (1 - tx) * (1 - ty)), (0, 1, tx * (1 - ty)), (1, 0, (1 - tx) * ty), (1, 1, tx * ty)):
        yy = y0 + dy
        xx = x0 + dx
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = inp[bb, :, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]  # [B,h,w,C]
        out += np.moveaxis(v * (wt * ok)[..., None], -1, 1)
    return out


# ----------------------------------------------------------------------------------------------
# A8: inverse sampling with NaN mask                                       models/models.py:935-938
# ----------------------------------------------------------------------------------------------


def inverse_sample(pred: torch.Tensor, grid_inv: torch.Tensor) -> torch.Tensor:
    """models/models.py:935-938 (= models_instance.py:883-886)."""
    grid_inv = grid_inv.clone()
    unfilled = torch.isnan(grid_inv[:, :, :, 0])
    grid_inv[torch.isnan(grid_inv)] = 0
    out = grid_sample(pred, grid_inv.float())
    out[unfilled.unsqueeze(1).expand(out.shape)] = float("nan")
    return out


# ----------------------------------------------------------------------------------------------
# A9: fillMissingValues_tensor('tri') + Interp2D             models/models.py:159-286, interp2d.py:14-91
# ----------------------------------------------------------------------------------------------

_CROSS = torch.tensor([[0.0, 1.0, 0.0], [1.0, 1.0, 1.0], [0.0, 1.0, 0.0]])  # cv2 MORPH_ELLIPSE (3,3)


def pixels_for_interp(t: torch.Tensor):
    """models/models.py:169-211 (getPixelsForInterp) -> (mask_for_interp[C,H,W] bool, invalid[C,H,W] bool)."""
    invalid = torch.isnan(t)
    C = invalid.shape[0]
    kernel = _CROSS[None, None].expand(1, C, 3, 3).float()
    if max(invalid.shape) > 512:                                               # :183-193
        dr = max(invalid.shape) / 512
        shape_ori = (invalid.shape[-2], int(invalid.shape[-1]))
        shape_scaled = (int(invalid.shape[-2] / dr), int(invalid.shape[-1] / dr))
        scaled = F.interpolate(invalid.float().unsqueeze(0), shape_scaled, mode="nearest")
        dil_s = torch.clamp(F.conv2d(scaled, kernel, padding=(1, 1)), 0, 1)
        dilated = F.interpolate(dil_s.float(), shape_ori, mode="nearest").squeeze(0)
    else:                                                                      # :195-197
        dilated = torch.clamp(F.conv2d(invalid.float().unsqueeze(0), kernel, padding=(1, 1)), 0, 1).squeeze(0)
    m = dilated * (~invalid).float()                                           # :200
    for (r, c) in ((0, 0), (0, -1), (-1, 0), (-1, -1)):                        # :202-209
        m[:, r, c] = 1.0
    return m.bool(), invalid


def delaunay(points_rc: np.ndarray):
    """interp2d.py:55 -- Qhull Delaunay with SciPy's default options (`Qbb Qc Qz Q12` + `Qt`)."""
    from scipy.spatial import Delaunay
    return Delaunay(np.asarray(points_rc, dtype=np.float64))


def find_simplex_with_c(tri, xi: np.ndarray):
    """spatial/qhull.pyx:2075-2163 with return_c=True: simplex id and barycentric c (float64).

    c[:2] = T[:2] . (x - T[2]);  c[2] = 1 - c[0] - c[1]   (qhull.pyx:1210-1264); T = tri.transform.
    """
    xi = np.asarray(xi, dtype=np.float64)
    isimplex = tri.find_simplex(xi)
    T = tri.transform[np.maximum(isimplex, 0)]
    c01 = np.einsum("nij,nj->ni", T[:, :2, :], xi - T[:, 2, :])
    return isimplex, np.concatenate([c01, 1.0 - c01.sum(1, keepdims=True)], axis=1)


def interp2d_forward(points: torch.Tensor, values: torch.Tensor, h: int, w: int, tri=None) -> torch.Tensor:
    """interp2d.py:37-91 -- points [N,2] (row,col) long, values [N,vdim] -> [vdim,h,w]."""
    if tri is None:
        tri = delaunay(points.cpu().numpy())
    rr, cc = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")             # interp2d.py:27-30
    coord = np.stack([rr, cc], -1).reshape(-1, 2)
    isimplex, c = find_simplex_with_c(tri, coord)
    isimplex = isimplex.copy()
    isimplex[isimplex == -1] = 0                                               # interp2d.py:61-63
    wts = torch.from_numpy(c).float()                                          # :65
    verts = torch.from_numpy(tri.simplices[isimplex]).long()                   # :74-76
    out = torch.zeros(h * w, values.shape[1], dtype=values.dtype)
    # :80-89  out = sum_k (values[verts[:,k]] * wts[:,k]) -- stack+mul+sum(dim=0) adds k=0,1,2 in order
    out = values[verts[:, 0]] * wts[:, 0:1]
    out = out + values[verts[:, 1]] * wts[:, 1:2]
    out = out + values[verts[:, 2]] * wts[:, 2:3]
    return out.reshape(h, w, -1).permute(2, 0, 1)


def fill_missing_values_tensor(t: torch.Tensor, copy: bool = False) -> torch.Tensor:
    """models/models.py:159-286 with interp_mode='tri'. t [C,H,W]; in place unless copy."""
    if copy:
        t = t.clone()
    mask, invalid = pixels_for_interp(t)
    if invalid.float().sum() == 0:                                             # :254-255
        return t
    pts = torch.stack(torch.where(mask[0]), 1)                                 # :265-267
    vals = t.clone()[mask].view(mask.shape[0], -1).permute(1, 0)               # :268
    interp = interp2d_forward(pts, vals, t.shape[-2], t.shape[-1])
    t[invalid] = interp[torch.where(invalid)].clone()                          # :280
    return t


def _dilate_cv2_chw(mask_u8: np.ndarray) -> np.ndarray:
    """cv2.dilate(img[C,H,W], cross3x3, BORDER_CONSTANT 0) as getPixelsForInterp_NB calls it (models/models.py:229, 236):
    OpenCV reads a 3-D array as (rows, cols, channels) = (C, H, W), so the cross spans the CLASS and ROW axes -- never
    the column axis.  Restated in NumPy (cv2 is used when importable; tests assert the two agree)."""
    try:
        import cv2
        k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3))
        if mask_u8.shape[2] <= 512:                                             # CV_CN_MAX
            return cv2.dilate(np.ascontiguousarray(mask_u8), k, borderType=cv2.BORDER_CONSTANT, borderValue=int(0))
    except ImportError:
        pass
    o = mask_u8.copy()
    o[1:] |= mask_u8[:-1]; o[:-1] |= mask_u8[1:]
    o[:, 1:] |= mask_u8[:, :-1]; o[:, :-1] |= mask_u8[:, 1:]
    return o


def pixels_for_interp_nb(t: torch.Tensor):
    """models/models.py:213-242 (getPixelsForInterp_NB) -> (mask_for_interp[C,H,W] bool ndarray, invalid ndarray)."""
    invalid = np.isnan(t.cpu().numpy())
    if max(invalid.shape) > 512:                                               # :222-232
        dr = max(invalid.shape) / 512
        shape_ori = (invalid.shape[-2], int(invalid.shape[-1]))
        shape_scaled = (int(invalid.shape[-2] / dr), int(invalid.shape[-1] / dr))
        scaled = F.interpolate(torch.tensor(invalid.astype("float")).unsqueeze(0), shape_scaled, mode="nearest").squeeze(0)
        dil_s = _dilate_cv2_chw(scaled.numpy().astype("bool").astype("uint8"))
        dil = F.interpolate(torch.tensor(dil_s.astype("float")).unsqueeze(0), shape_ori, mode="nearest").squeeze(0)
        dilated = dil.numpy().astype("uint8")
    else:                                                                      # :235-237
        dilated = _dilate_cv2_chw(invalid.astype("uint8"))
    return (dilated * ~invalid).astype("bool"), invalid                        # :240-242


def fill_missing_values_nearest(t: torch.Tensor, copy: bool = False, return_sites: bool = False):
    """models/models.py:159-286 with interp_mode='nearest': NearestNDInterpolator over the 3-D (class,row,col) voxels."""
    import scipy.interpolate
    if copy:
        t = t.clone()
    mask, invalid = pixels_for_interp_nb(t)
    points = np.argwhere(mask)                                                 # :259
    values = t[torch.from_numpy(mask)].cpu().numpy()                           # :261
    interp = scipy.interpolate.NearestNDInterpolator(points, values)           # :245-247, :269
    t[torch.from_numpy(invalid)] = torch.tensor(interp(np.argwhere(invalid))).float()   # :272
    return (t, mask) if return_sites else t


def fill_missing_values_bi(t: torch.Tensor, copy: bool = False) -> torch.Tensor:
    """models/models.py:159-286 with interp_mode='BI': scipy LinearNDInterpolator over the 3-D (class,row,col) voxels of
    getPixelsForInterp_NB (:248-250, :259-272); queries outside the convex hull stay NaN (its fill_value)."""
    import scipy.interpolate
    if copy:
        t = t.clone()
    mask, invalid = pixels_for_interp_nb(t)
    points = np.argwhere(mask)                                                 # :259
    values = t[torch.from_numpy(mask)].cpu().numpy()                           # :261
    interp = scipy.interpolate.LinearNDInterpolator(points, values)            # :248-250, :269
    t[torch.from_numpy(invalid)] = torch.tensor(interp(np.argwhere(invalid))).float()   # :272
    return t


def inverse_path(pred: torch.Tensor, grid: torch.Tensor, segSize, zero_residual: bool = True,
                 tie: str = "max", interp_mode: str = "tri") -> torch.Tensor:
    """A7 -> A8 -> A9 per sample (models/models.py:933-940; models_instance.py:883-893, 940)."""
    gi = grid_inverse(grid, segSize, tie=tie)
    ps = inverse_sample(pred, gi)
    fill = {"nearest": fill_missing_values_nearest, "BI": fill_missing_values_bi, "tri": fill_missing_values_tensor}
    for n in range(ps.shape[0]):
        ps[n] = fill[interp_mode](ps[n])
    if zero_residual:
        ps[torch.isnan(ps)] = 0                                                # models_instance.py:940
    return ps


def instance_mask(pred_sampled: torch.Tensor) -> torch.Tensor:
    """A10: models/models.py:1044."""
    return torch.argmax(pred_sampled, dim=1)


# ----------------------------------------------------------------------------------------------
# Synthetic inputs -- shared by tests and bench so both sides see identical data
# ----------------------------------------------------------------------------------------------


# ----------------------------------------------------------------------------------------------
# DynamicFocus deformed_unsampler          DynamicFocus/d_model/nn_B0_deformed_sampler.py:83-153
# ----------------------------------------------------------------------------------------------


def int_round_scale_grid(grid: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """nn_B0_deformed_sampler.py:83-102 (`int_rount_scale_grid`): [-1,1] -> clipped, truncated int64 canvas coordinates."""
    g = 0.5 * (grid + 1.0)
    g[:, 0] *= H - 1
    g[:, 1] *= W - 1
    g[:, 0] = torch.clip(g[:, 0], 0, H - 1)
    g[:, 1] = torch.clip(g[:, 1], 0, W - 1)
    return g.to(torch.int64)


def deformed_unsampler(labels: torch.Tensor, coords: torch.Tensor, H: int, W: int, return_sites: bool = False):
    """nn_B0_deformed_sampler.py:115-153: scatter labels [B,K,HS,WS] at coords [B,2,HS,WS] (row, col), then every
    unscattered pixel copies its nearest scattered pixel (scipy distance_transform_edt, return_indices).  Nodes that
    share a pixel: the last write wins (torch CPU index_put order = largest node index).
    return_sites: also the node index each pixel received [B,H,W] and the EDT distances (for tie analysis in tests)."""
    from scipy.ndimage import distance_transform_edt
    B, K, HS, WS = labels.shape
    out = np.zeros((B, K, H, W), np.float32)
    owner = np.full((B, H, W), -1, np.int64)
    dist = np.zeros((B, H, W), np.float64)
    lab = labels.numpy().reshape(B, K, HS * WS)
    for b in range(B):
        r, c = coords[b, 0].numpy().ravel(), coords[b, 1].numpy().ravel()
        node = np.full((H, W), -1, np.int64)
        node[r, c] = np.arange(HS * WS)                          # :127-137, sequential: the last duplicate wins
        d, idx = distance_transform_edt(node < 0, return_indices=True)        # :143
        owner[b] = node[idx[0], idx[1]]                          # :147-149 (filled pixels index themselves)
        dist[b] = d
        out[b] = lab[b][:, owner[b]]
    res = torch.from_numpy(out)
    return (res, owner, dist) if return_sites else res


def synthetic_saliency(B: int, gh: int = 80, gw: int = 80, seed: int = 0):
    """xs = softmax(3*N(0,1) + 6*exp(-d^2/(2*8^2))) centred at a random gaze; returns (xs[B,1,gh,gw], gaze[B,2])."""
    g = torch.Generator().manual_seed(seed)
    gaze = torch.rand(B, 2, generator=g) * 0.98
    ii = torch.arange(gh, dtype=torch.float32)[None, :, None]
    jj = torch.arange(gw, dtype=torch.float32)[None, None, :]
    d2 = (ii - gaze[:, 0, None, None] * (gh - 1)) ** 2 + (jj - gaze[:, 1, None, None] * (gw - 1)) ** 2
    logits = 3.0 * torch.randn(B, gh, gw, generator=g) + 6.0 * torch.exp(-d2 / (2 * 8.0 ** 2))
    xs = torch.softmax(logits.view(B, -1), 1).view(B, 1, gh, gw)
    return xs, gaze


def synthetic_pred(B: int, C: int = 51, h: int = 80, w: int = 80, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed + 1000)
    return torch.randn(B, C, h, w, generator=g)


def reference_hot_path(x, xs, pred, Rx, Ry, fwhm, segSize, pad_mode="replication"):
    """The whole resample path as the reference runs it on CPU (timed by bench.py's cpu_baseline leg).

    x [B,3,H,W] image, xs [B,1,gh,gw] normalised saliency, pred [B,C,h,w] decoder scores.
    Returns (x_sampled, pred_sampled, mask).
    """
    gh, gw = xs.shape[-2:]
    filt = gaussian_filter_weight(Rx, Ry, fwhm)
    P = p_basis(gh, gw, Rx, Ry)
    xs_hm = pad_saliency(xs, Rx, Ry, pad_mode)
    grid, _ = create_grid(xs_hm, filt, P, gh, gw, pred.shape[-2:])
    x_sampled = grid_sample(x, grid)
    ps = inverse_path(pred, grid, segSize)
    return x_sampled, ps, instance_mask(ps)

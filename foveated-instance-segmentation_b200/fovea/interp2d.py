"""Mirror of the reference's `interp2d.py` (Interp2D) on the sm_100a kernels.

Reference: interp2d.py:14-91.  Same constructor and call signature:

    Interp2D(h, w, add_corner=False)(points[N,2] long (row,col), values[N,vdim] float32 on the GPU) -> [vdim,h,w]

What the reference does on the host (Qhull Delaunay + find_simplex over all h*w pixels, interp2d.py:53-63) and with
a [3,h*w,vdim] gather (interp2d.py:76-89) runs here as: triangulation -> exact point location + barycentric gather
fused into one write of the output.  `triangulation`:

  "host"   (default) Qhull on the host with the reference's options, exactly what interp2d.py:55 does: the mesh -- and so
           every output value -- is the reference's, including how Qhull's `Qt` splits co-circular lattice cells;
  "device" the sm_100a Delaunay kernel (no host round trip).  A valid Delaunay triangulation of the same points; it
           differs from Qhull's only inside co-circular cells, where the choice is arbitrary in the reference too
           (DESIGN.md section 2).  Opt-in; if the point set exceeds the kernel's capacity a FoveaError is raised --
           nothing falls back silently.

Differences, all documented in DESIGN.md:
  * pixels outside the triangulation get NaN (the reference maps them to simplex 0 with whatever barycentrics its last
    walk left behind, interp2d.py:61-63 -- undefined); with the four corners present (the only way the reference's own
    call site uses it, models/models.py:202-209) no pixel is outside;
  * duplicate points are merged (first value wins); Qhull would drop them as coplanar.
Gradients flow to `values` (interp2d.py:38-47) through fovea_inverse_fill_bwd.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from ._lib import FoveaError


def interp2d_scores(points, values, h, w, triangulation="host", zero_residual=False):
    """Core of Interp2D.forward: points [N,2] (row,col) integer, values [N,vdim] -> [vdim,h,w]."""
    if not values.is_cuda:
        raise FoveaError("Interp2D: values must be a CUDA tensor (there is no CPU fallback)")
    dev = values.device
    pts = points.to(dev).long()
    N, vdim = values.shape
    if pts.shape != (N, 2):
        raise FoveaError(f"Interp2D: points {tuple(points.shape)} do not match values {tuple(values.shape)}")
    if N < 3:
        raise FoveaError("Interp2D: need at least 3 points")
    if pts.min() < 0 or pts[:, 0].max() >= 65536 or pts[:, 1].max() >= 65536:
        raise FoveaError("Interp2D: point coordinates must lie in [0, 65536)")
    # row-major sort + de-duplication (the device triangulation needs sorted unique sites)
    key = pts[:, 0] * 65536 + pts[:, 1]
    skey, order = torch.sort(key, stable=True)
    keep = torch.ones_like(skey, dtype=torch.bool)
    keep[1:] = skey[1:] != skey[:-1]
    skey, order = skey[keep], order[keep]
    n = int(skey.numel())
    cap = max(n, 4)
    tcap = 2 * cap
    packed = torch.zeros(1, cap, device=dev, dtype=torch.int32)
    packed[0, :n] = ((skey // 65536) * 65536 + (skey % 65536)).to(torch.int32)
    npts = torch.tensor([n], device=dev, dtype=torch.int32)
    src = torch.arange(cap, device=dev, dtype=torch.int32).view(1, cap)
    Cs = (vdim + 3) // 4 * 4
    # value table: rows 0..n-1 = values in sorted order, row `cap` = NaN, row `cap+1` = 0 (layout of fovea_box4_table)
    table = torch.zeros(1, cap + 2, Cs, device=dev, dtype=torch.float32)
    table[0, cap, :] = float("nan")
    table[0, :n, :vdim] = values.float()[order]                        # (CopySlices: differentiable w.r.t. `values`)
    max_coord = max(int(h), int(w), int(pts.max().item()) + 1)
    rounds = None
    if triangulation == "device":
        mesh, ntri, ws = ops.delaunay_device(packed, npts, cap, tcap, max_coord)   # FoveaError if over capacity
        rounds = ws[:1]
    elif triangulation == "host":
        mesh, ntri = ops._triangulate_host(packed, npts, cap, tcap)
    else:
        raise FoveaError(f"unknown triangulation mode {triangulation!r}")
    plan = ops.plan_from_mesh(packed, src, npts, mesh, ntri, int(h), int(w), table_rows=cap)
    plan.rounds = rounds
    out = torch.empty(1, vdim, int(h), int(w), device=dev, dtype=torch.float32)
    out, _ = ops.inverse_fill_table(plan, table, vdim, zero_residual=zero_residual, scores=out)
    ops.check_plan(plan)
    return out[0]


class Interp2D(nn.Module):
    """Drop-in for interp2d.py:14-91."""

    def __init__(self, h, w, add_corner=False, triangulation="host"):
        super().__init__()
        self.h, self.w, self.add_corner, self.triangulation = int(h), int(w), add_corner, triangulation

    def forward(self, points, values):
        if self.add_corner:  # interp2d.py:48-51
            corners = torch.tensor([[0, 0], [0, self.w - 1], [self.h - 1, 0], [self.h - 1, self.w - 1]],
                                   device=values.device, dtype=torch.long)
            points = torch.cat([points.to(values.device).long(), corners], dim=0)
            values = torch.cat([values, torch.zeros(4, values.shape[1], device=values.device, dtype=values.dtype)], 0)
        return interp2d_scores(points, values, self.h, self.w, self.triangulation)

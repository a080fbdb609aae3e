"""Host-side mirror of DynamicFocus/d_model/nn_B0_deformed_sampler.py's inverse ("unsampler") half -- same names,
arguments and return values -- on the sm_100a kernels.

    int_rount_scale_grid(grid_Bx2xHSxWS, canvas_H, canvas_W)                       nn_B0_deformed_sampler.py:83-102
    deformed_unsampler(label_sample_BxKxHSxWS, grid_Bx2xHSxWS, canvas_H, canvas_W)  nn_B0_deformed_sampler.py:115-153

The reference scatters the labels with index_put, copies the canvas to the HOST and runs
scipy.ndimage.distance_transform_edt(return_indices=True) per image, then a Python loop over the K channels.  Here:
fovea_scatter_nodes -> fovea_nearest_locate_all (exact integer Euclidean nearest filled pixel) -> fovea_node_table ->
fovea_inverse_fill, all on the device.  Where several filled pixels are equally near, SciPy's choice is an artefact of its
scan order; ours is the leftmost column, then the upper pixel (tests compare everything else bit for bit).  Where several
nodes land on one pixel the reference's index_put keeps an unspecified one (the last on CPU); here the largest node index.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import FoveaError


def int_rount_scale_grid(grid_Bx2xHSxWS: torch.Tensor, canvas_H, canvas_W):
    """[-1,1] grid (channel 0 -> rows, 1 -> columns) to integer canvas coordinates (truncation), as the reference."""
    g = 0.5 * (grid_Bx2xHSxWS + 1.0)
    g[:, 0, :, :] *= canvas_H - 1
    g[:, 1, :, :] *= canvas_W - 1
    g[:, 0, :, :] = torch.clip(g[:, 0, :, :], 0, canvas_H - 1)
    g[:, 1, :, :] = torch.clip(g[:, 1, :, :], 0, canvas_W - 1)
    return g.to(dtype=torch.int64)


def deformed_unsampler(label_sample_BxKxHSxWS: torch.Tensor, grid_Bx2xHSxWS: torch.Tensor, canvas_H, canvas_W):
    """Labels sampled at the (integer) grid positions -> full canvas [B,K,canvas_H,canvas_W]: scattered pixels keep their
    label, every other pixel takes the label of the nearest scattered pixel.  Runs on grid's device (must be CUDA)."""
    if not grid_Bx2xHSxWS.is_cuda:
        raise FoveaError("deformed_unsampler: the grid must live on a CUDA device (this path has no CPU fallback)")
    dev = grid_Bx2xHSxWS.device
    labels = label_sample_BxKxHSxWS.to(device=dev, dtype=torch.float32)
    coords = grid_Bx2xHSxWS.to(dtype=torch.int64)
    B, K, HS, WS = labels.shape
    winner = ops.scatter_nodes(coords, (canvas_H, canvas_W))
    loc = ops.nearest_locate_all(winner, HS, WS)
    dummy = torch.zeros(1, 1, 16, device=dev, dtype=torch.int32)            # `loc` holds no triangle ids: never read
    plan = ops.InversePlan(winner, None, None, None, None, None, None, dummy, loc, HS, WS, int(canvas_H), int(canvas_W),
                           HS * WS + 4, 1, "nearest")
    out = torch.empty(B, K, int(canvas_H), int(canvas_W), device=dev, dtype=torch.float32)
    ops.inverse_fill_table(plan, ops.node_table(labels), K, zero_residual=False, scores=out)
    return out

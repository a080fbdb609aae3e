"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference through ref_shim.py, runs the reference's own `create_grid`, `F.grid_sample`
call pattern, NaN-mask inverse sampling and `fillMissingValues_tensor('tri')` (with the reference's own
`interp2d.Interp2D`) on seeded synthetic inputs, and stores inputs + outputs as small .npz fixtures.
The GPU box has no /root/reference: tests only read the .npz files.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402

rm, ri = ref_shim.install_all()
from oracle.reference_port import synthetic_saliency, synthetic_pred  # noqa: E402  (input generators only)


def build(cfg):
    torch.manual_seed(0)
    return rm.DeformSegmentationModule(None, None, None, None, None, cfg)


def pad(m, xs, mode):
    # models/models.py:819-825 verbatim call pattern
    px, py = m.padding_size_x, m.padding_size_y
    if mode == "replication":
        return nn.ReplicationPad2d((py, py, px, px))(xs)
    if mode == "reflect":
        return F.pad(xs, (py, py, px, px), mode="reflect")
    return F.pad(xs, (py, py, px, px), mode="constant")


def case_grid(name, sal, task, R, segSize, B, seed, pad_mode="replication", task_eval=(), rate=1):
    cfg = ref_shim.make_cfg(sal=sal, task=task, R=R, pad=pad_mode, task_eval=task_eval, rate=rate)
    m = build(cfg)
    xs, gaze = synthetic_saliency(B, m.grid_size_x, m.grid_size_y, seed=seed)
    xs_hm = pad(m, xs, pad_mode)
    with torch.no_grad():
        grid, grid_y = m.create_grid(xs_hm)
        grid2, grid_inv = m.create_grid(xs_hm, segSize=segSize, x_inv=1 - xs_hm)
    out = dict(xs=xs.numpy(), gaze=gaze.numpy(), filt=m.filter.weight.detach()[0, 0].numpy(),
               P_basis=m.P_basis.numpy(), grid=grid.numpy(), grid_y=grid_y.numpy(), grid_infer=grid2.numpy(),
               grid_inv=grid_inv.numpy(), segSize=np.array(segSize), sal=np.array(sal), task=np.array(task),
               task_eval=np.array(task_eval, dtype=np.int64), rate=np.array(rate), R=np.array(R),
               Rx=np.array(m.padding_size_x), Ry=np.array(m.padding_size_y), pad_mode=np.array(pad_mode))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items() if hasattr(v, "shape") and v.ndim > 1})
    return m, xs, grid, grid_inv


def case_inverse(name, m, grid, grid_inv, C, seed, image_hw):
    """models/models.py:935-940 call pattern: grid_sample(pred, grid_inv) -> NaN mask -> per-sample tri fill."""
    B, h, w, _ = grid.shape
    pred = synthetic_pred(B, C, h, w, seed=seed)
    gi = grid_inv.clone()
    unfilled = torch.isnan(gi[:, :, :, 0])
    gi[torch.isnan(gi)] = 0
    ps = F.grid_sample(pred, gi.float())
    ps[unfilled.unsqueeze(1).expand(ps.shape)] = float("nan")
    ps_nan = ps.clone()
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        for n in range(B):
            ps[n] = rm.fillMissingValues_tensor(ps[n], interp_mode="tri")
    # grid_sample of an image with the same grid (A5), through the reference's own call
    g = torch.Generator().manual_seed(seed + 7)
    x = torch.rand(B, 3, image_hw[0], image_hw[1], generator=g)
    x_sampled = F.grid_sample(x, grid)
    out = dict(pred=pred.numpy(), grid=grid.numpy(), pred_sampled_nan=ps_nan.numpy(), pred_sampled=ps.numpy(),
               x_seed=np.array(seed + 7), x_hw=np.array(image_hw), x_sampled=x_sampled.numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, ps.shape, "nan left:", int(torch.isnan(ps).sum()))


def case_interp2d(name, h, w, N, vdim, seed):
    """interp2d.py: the reference's own Interp2D class on random integer points + the 4 corners."""
    g = torch.Generator().manual_seed(seed)
    lin = torch.randperm(h * w, generator=g)[:N]
    corners = torch.tensor([0, w - 1, (h - 1) * w, h * w - 1])
    lin = torch.unique(torch.cat([lin, corners]))  # sorted = row-major, as torch.where would give
    pts = torch.stack([lin // w, lin % w], 1)
    vals = torch.randn(pts.shape[0], vdim, generator=g)
    out = ri.Interp2D(h, w)(pts, vals)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), points=pts.numpy(), values=vals.numpy(), out=out.numpy(),
                        hw=np.array([h, w]))
    print(name, out.shape)


if __name__ == "__main__":
    torch.set_num_threads(8)
    # BASELINE geometry: saliency 80x80, task 80x80, R=45, replication pad
    m, xs, grid, gi = case_grid("grid_80_R45", (80, 80), (80, 80), 45, (128, 128), B=2, seed=1)
    case_inverse("inverse_80_to_128", m, grid, gi, C=5, seed=1, image_hw=(256, 256))
    # > 512 px: exercises the nearest-downscaled dilation branch (models.py:183-193), non-integer ratio
    m, xs, grid, gi = case_grid("grid_80_R45_seg520", (80, 80), (80, 80), 45, (520, 536), B=1, seed=2)
    case_inverse("inverse_80_to_520", m, grid, gi, C=2, seed=2, image_hw=(520, 536))
    # anisotropic saliency (gaussian_ap = 2 -> Ry = 2*Rx, bilinear-resized Gaussian), task != saliency size,
    # label grid at rate 2, reflect / zero padding
    case_grid("grid_40x80_R12_reflect", (40, 80), (64, 96), 12, (96, 160), B=2, seed=3, pad_mode="reflect", rate=2)
    case_grid("grid_40x80_R12_zero", (40, 80), (64, 96), 12, (96, 160), B=2, seed=4, pad_mode="zero", rate=2)
    case_grid("grid_32_R10_eval", (32, 32), (32, 32), 10, (64, 64), B=3, seed=5, task_eval=(48, 48))
    case_interp2d("interp2d_64x48", 64, 48, 300, 4, seed=6)

"""Edge cases of stage 3 the domain offers: a canvas smaller than the sampling lattice (most nodes
collide), a uniform saliency (no fovea), an extremely peaked one (a large solid block of filled pixels), a single frame,
constant / all-zero predictions.  Every case: host-mesh scores and masks against the oracle, device mesh valid and confined
to co-circular differences, raster map == walker map, pruned mask == argmax of the scores."""
import numpy as np
import pytest
import torch

from oracle import reference_port as rp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.fail("these tests need a CUDA device (run on the B200 box with -m gpu)")
    from fovea import ops as _ops
    _ops._lib.load()
    return _ops


def _grid_from(xs):
    filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
    return rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))[0]


def _saliency(kind, B, seed):
    xs, _ = rp.synthetic_saliency(B, seed=seed)
    if kind == "uniform":
        xs = torch.full_like(xs, 1.0 / 6400)
    elif kind == "peaked":
        xs = torch.softmax(40.0 * torch.log(xs.view(B, -1)) / torch.log(xs.view(B, -1)).abs().max(), 1).view_as(xs)
    return xs


CASES = [("normal", 64, 64, 2), ("normal", 72, 56, 1), ("uniform", 256, 256, 2), ("peaked", 256, 256, 2), ("peaked", 1024, 1024, 1)]


@pytest.mark.parametrize("kind,H,W,B", CASES)
def test_stage3_edge_cases(ops, monkeypatch, kind, H, W, B):
    from test_device_mesh_parity_gpu import _cocircular_flags
    from test_parity_gpu import _check_masks, _edge_exempt
    C = 5
    grid = _grid_from(_saliency(kind, B, seed=H + W))
    pred = rp.synthetic_pred(B, C, seed=H)
    plan_h = ops.build_inverse_plan(grid.cuda(), (H, W), nchan=C, triangulation="host")
    sh, mh = ops.inverse_fill(plan_h, pred.cuda(), want_scores=True, want_mask=True)
    want = rp.inverse_path(pred, grid, (H, W))
    exempt = _edge_exempt(sh.cpu(), want, plan_h)
    assert (((sh.cpu() - want).abs() <= 1e-5 * want.abs().max()).all(1) | exempt).all()
    _check_masks(mh.cpu(), want, exempt)
    # device mesh: converged, same sites, differs from Qhull's only inside co-circular cells
    plan_d = ops.check_plan(ops.build_inverse_plan(grid.cuda(), (H, W), nchan=C, triangulation="device"))
    assert torch.equal(plan_d.npts, plan_h.npts) and torch.equal(plan_d.ntri.cpu(), plan_h.ntri.cpu())
    sd, md = ops.inverse_fill(plan_d, pred.cuda(), want_scores=True, want_mask=True)
    for b in range(B):
        flag, _ = _cocircular_flags(plan_h, b)
        differs = ((sd[b] - sh[b]).abs() > 1e-5 * sh.abs().max()).any(0)
        loc = plan_h.loc[b].view(torch.int16).long() & 0xFFFF
        assert ((loc[differs] & 0x8000) == 0).all() and flag[loc[differs]].all()
    # mask mode == argmax of the scores; raster map == walker map (where the canvas width allows the raster path)
    _, mp = ops.inverse_fill(plan_d, pred.cuda(), want_scores=False, want_mask=True)
    assert torch.equal(mp, torch.argmax(sd, dim=1))
    if W % 8 == 0:
        monkeypatch.setenv("FOVEA_LOCATE", "walk")
        plan_w = ops.build_inverse_plan(grid.cuda(), (H, W), nchan=C, triangulation="device")
        sw, _ = ops.inverse_fill(plan_w, pred.cuda(), want_scores=True)
        assert torch.equal(sw, sd)
    print(f"{kind} {H}x{W}: sites {plan_h.npts.tolist()}, winners {int((plan_h.winner >= 0).sum())} of {B * 6400} nodes, "
          f"{int(exempt.sum())} edge-exempt pixels")


@pytest.mark.parametrize("value", [0.0, 1.5])
def test_constant_predictions(ops, value):
    grid = _grid_from(_saliency("normal", 2, seed=8)).cuda()
    plan = ops.build_inverse_plan(grid, (256, 256), nchan=7, triangulation="device")
    pred = torch.full((2, 7, 80, 80), value, device="cuda")
    scores, mask = ops.inverse_fill(plan, pred, want_scores=True, want_mask=True)
    _, pruned = ops.inverse_fill(plan, pred, want_scores=False, want_mask=True)
    assert torch.equal(mask, torch.zeros_like(mask)) and torch.equal(pruned, mask)      # every channel ties: the first wins
    assert torch.isfinite(scores).all()

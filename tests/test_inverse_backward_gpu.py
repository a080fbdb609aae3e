"""Stage-3 backward (fovea_inverse_fill_bwd) and rev_deform_interp='BI', against the CPU oracle.

Backward: the reference differentiates through F.grid_sample(pred, grid_inv) + NaN mask + fillMissingValues_tensor
(models/models.py:933-940) when MODEL.loss_at_high_res is set, and Interp2D promises gradients w.r.t. `values`
(interp2d.py:38-47).  The oracle restates both with differentiable torch CPU ops, so torch.autograd on the oracle is the
reference gradient.  Tolerance: 1e-5 of the largest gradient entry (sums of up to ~10^4 fp32 terms per node; the order
of the atomic adds is not fixed).
"""
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


def _small_case(g, H, W, C, seed, R=8):
    xs, _ = rp.synthetic_saliency(2, g, g, seed=seed)
    filt, P = rp.gaussian_filter_weight(R, R, R), rp.p_basis(g, g, R, R)
    grid, _ = rp.create_grid(rp.pad_saliency(xs, R, R), filt, P, g, g, (g, g))
    return grid, rp.synthetic_pred(2, C, g, g, seed=seed)


def test_interp2d_gradient_wrt_values_matches_oracle_autograd(ops, golden_dir):
    import os
    from fovea.interp2d import Interp2D
    g = dict(np.load(os.path.join(golden_dir, "interp2d_64x48.npz")))
    h, w = (int(v) for v in g["hw"])
    pts, vals = torch.from_numpy(g["points"]), torch.from_numpy(g["values"])
    gen = torch.Generator().manual_seed(0)
    gout = torch.randn(vals.shape[1], h, w, generator=gen)
    v_ref = vals.clone().requires_grad_(True)
    out_ref = rp.interp2d_forward(pts, v_ref, h, w)
    # outside the hull the reference is undefined (it maps the pixel to simplex 0, interp2d.py:61-63): no gradient there
    rr, cc = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    inside = torch.from_numpy(rp.delaunay(pts.numpy()).find_simplex(np.stack([rr, cc], -1).reshape(-1, 2)) >= 0).view(h, w)
    (torch.nan_to_num(out_ref) * gout * inside).sum().backward()
    v = vals.clone().cuda().requires_grad_(True)
    out = Interp2D(h, w, triangulation="host")(pts.cuda(), v)
    assert out.requires_grad
    (torch.nan_to_num(out) * (gout * inside).cuda()).sum().backward()
    err = (v.grad.cpu() - v_ref.grad).abs().max().item()
    scale = v_ref.grad.abs().max().item()
    print(f"Interp2D d/d values: max err {err:.3e} of scale {scale:.3e}")
    assert err <= 1e-5 * scale


@pytest.mark.parametrize("g,H,W,C", [(16, 64, 64, 5), (24, 128, 96, 3)])
def test_inverse_fill_gradient_wrt_pred_matches_oracle_autograd(ops, g, H, W, C):
    grid, pred = _small_case(g, H, W, C, seed=g)
    gout = torch.randn(2, C, H, W, generator=torch.Generator().manual_seed(1))
    p_ref = pred.clone().requires_grad_(True)
    want = rp.inverse_path(p_ref, grid, (H, W), zero_residual=True)
    (want * gout).sum().backward()
    plan = ops.build_inverse_plan(grid.cuda(), (H, W), nchan=C, triangulation="host")
    p = pred.clone().cuda().requires_grad_(True)
    scores, _ = ops.inverse_fill(plan, p, want_scores=True, zero_residual=True)
    assert scores.requires_grad
    # forward parity first (pixels on an edge next to an unfilled corner are path-dependent in the reference: exclude
    # them from the gradient comparison by zeroing their upstream gradient on both sides)
    same = ((scores.detach().cpu() - want.detach()).abs() <= 1e-5 * want.detach().abs().max()).all(1, keepdim=True)
    assert same.float().mean().item() > 0.995
    p_ref.grad = None
    want = rp.inverse_path(p_ref, grid, (H, W), zero_residual=True)
    (want * gout * same).sum().backward()
    (scores * (gout * same).cuda()).sum().backward()
    err = (p.grad.cpu() - p_ref.grad).abs().max().item()
    scale = p_ref.grad.abs().max().item()
    print(f"d/d pred ({g}x{g} -> {H}x{W}): max err {err:.3e} of scale {scale:.3e}")
    assert err <= 1e-5 * scale
    # gradient of a full-size frame: finite, and linear in the upstream gradient
    p2 = pred.clone().cuda().requires_grad_(True)
    s2, _ = ops.inverse_fill(plan, p2, want_scores=True)
    (s2 * (2.0 * gout * same).cuda()).sum().backward()
    assert torch.allclose(p2.grad, 2.0 * p.grad, rtol=1e-4, atol=1e-4 * scale)


def test_inverse_fill_backward_at_full_size(ops):
    """1024^2, C=51, device mesh: the transpose identity <fill(pred), G> == <pred, fill^T(G)> (size-independent property:
    the backward kernel is the exact adjoint of the forward)."""
    B, C, H, W = 2, 51, 1024, 1024
    xs, _ = rp.synthetic_saliency(B, seed=9)
    filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
    grid, _ = rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))
    plan = ops.check_plan(ops.build_inverse_plan(grid.cuda(), (H, W), nchan=C, triangulation="device"))
    pred = rp.synthetic_pred(B, C, seed=9).cuda().requires_grad_(True)
    G = torch.randn(B, C, H, W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    scores, _ = ops.inverse_fill(plan, pred, want_scores=True)
    lhs = (scores.double() * G.double()).sum()
    scores.backward(G)
    rhs = (pred.detach().double() * pred.grad.double()).sum()
    rel = abs(lhs.item() - rhs.item()) / abs(lhs.item())
    print(f"adjoint identity at 1024^2: <A p, G> = {lhs.item():.6e}, <p, A^T G> = {rhs.item():.6e}, rel diff {rel:.2e}")
    assert rel < 1e-4


def test_module_loss_at_high_res_trains_through_the_inverse_path():
    """MODEL.loss_at_high_res (models/models.py:945-947, 962-965): the loss is taken on the inverse-upsampled scores and
    its gradient reaches decoder, encoder and (through the edge loss) the saliency network."""
    from fovea.models import CompressNet, DeformSegmentationModule
    from fovea.saliency_network import fov_simple
    from test_module_gpu import make_cfg
    from tiny_nets import TinyDecoder, TinyEncoder, synthetic_batch
    cfg = make_cfg(False)
    cfg.MODEL.loss_at_high_res = True
    torch.manual_seed(0)
    sal, comp, enc, dec = fov_simple(cfg), CompressNet(cfg), TinyEncoder(), TinyDecoder()
    m = DeformSegmentationModule(enc, dec, sal, comp, None, cfg, triangulation="device").cuda().train()
    feed = {k: v.cuda() for k, v in synthetic_batch(3, 192, 192, 5).items()}
    loss, acc, edge = m(feed, rank=1, cur_iter=-1)
    assert torch.isfinite(loss)
    loss.backward()
    g_dec = torch.cat([p.grad.flatten() for p in dec.parameters()])
    g_enc = torch.cat([p.grad.flatten() for p in enc.parameters()])
    assert torch.isfinite(g_dec).all() and g_dec.abs().sum() > 0
    assert torch.isfinite(g_enc).all() and g_enc.abs().sum() > 0


@pytest.mark.parametrize("g,H,W,C", [(16, 64, 64, 3), (24, 96, 80, 2)])
def test_bi_mode_matches_reference_linear_nd(ops, g, H, W, C):
    """rev_deform_interp='BI' (models/models.py:248-250, 259-272): scipy's LinearNDInterpolator over the 3-D
    (class,row,col) voxels of getPixelsForInterp_NB.  Restricted to a class plane that is the 2-D Delaunay interpolant of
    the plane's sites: same NaN pattern (outside the sites' hull), same values except inside co-circular cells."""
    from test_device_mesh_parity_gpu import _cocircular_flags
    grid, pred = _small_case(g, H, W, C, seed=g + 1)
    want = rp.inverse_path(pred, grid, (H, W), zero_residual=False, interp_mode="BI")
    plan = ops.build_inverse_plan(grid.cuda(), (H, W), nchan=C, triangulation="host", sites="nb")
    got, _ = ops.inverse_fill(plan, pred.cuda(), want_scores=True, zero_residual=False)
    got = got.cpu()
    # sites: the NB rule, in torch.where order
    ps_nan = rp.inverse_sample(pred, rp.grid_inverse(grid, (H, W), tie="max"))
    for b in range(2):
        sites, _ = rp.pixels_for_interp_nb(ps_nan[b])
        rc = np.argwhere(sites[0])
        n = int(plan.npts[b])
        mine = plan.pts[b, :n].cpu().numpy()
        assert np.array_equal(np.stack([mine >> 16, mine & 0xFFFF], 1), rc)
    assert torch.equal(torch.isnan(got), torch.isnan(want)), "NaN pattern (pixels outside the hull of the sites) differs"
    differs = (torch.nan_to_num(got - want).abs() > 1e-5 * torch.nan_to_num(want).abs().max()).any(1)
    frac = differs.float().mean().item()
    for b in range(2):
        flag, _ = _cocircular_flags(plan, b)
        loc = plan.loc[b].view(torch.int16).long().cpu() & 0xFFFF
        tri_of = loc[differs[b]]
        assert ((tri_of & 0x8000) == 0).all()
        assert flag.cpu()[tri_of].all(), "a pixel differs from LinearNDInterpolator outside a co-circular cell"
    print(f"'BI' {g}x{g} -> {H}x{W}: {frac:.4%} of pixels differ from the 3-D LinearNDInterpolator, all in co-circular cells")
    assert frac < 0.05

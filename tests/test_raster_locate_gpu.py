"""fovea_locate_raster (one warp per triangle, closed-form row spans) must produce the per-pixel source map of the scan-line
walker it replaces (fovea_locate_hints + fovea_locate_pixels): both evaluate the same exact integer predicate
e_i(y,x) >= m_i of the setup records (find_simplex of interp2d.py:58 with a fixed tie rule).  The only pixels where the
maps may differ are image corners no node landed on (a mesh VERTEX without a value: every incident triangle gives NaN
there, the raster kernel writes the explicit "no value" code) -- the interpolated scores must be identical everywhere."""
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


def _grid(B, seed):
    xs, _ = rp.synthetic_saliency(B, seed=seed)
    filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
    return rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))[0].cuda()


def _both(monkeypatch, build):
    monkeypatch.setenv("FOVEA_LOCATE", "walk")
    walk = build()
    monkeypatch.setenv("FOVEA_LOCATE", "raster")
    ras = build()
    assert walk.hints is not None and ras.hints is None          # the raster path needs no walk-start hints
    return walk, ras


@pytest.mark.parametrize("H,W,tri,sites", [(256, 320, "device", "tri"), (1024, 1024, "device", "tri"),
                                           (2048, 2048, "device", "tri"), (520, 392, "host", "tri"),
                                           (256, 256, "host", "nb"), (1024, 1024, "device", "nb")])
def test_raster_map_equals_walker_map(ops, monkeypatch, H, W, tri, sites):
    B, C = 2, 6
    grid = _grid(B, seed=H + W)
    walk, ras = _both(monkeypatch, lambda: ops.build_inverse_plan(grid, (H, W), nchan=51, triangulation=tri, sites=sites))
    for b in range(B):
        T = int(walk.ntri[b])
        assert T == int(ras.ntri[b]) and torch.equal(walk.mesh[b, :T].view(torch.int16), ras.mesh[b, :T].view(torch.int16))
    lw, lr = walk.loc.view(torch.int16).long() & 0xFFFF, ras.loc.view(torch.int16).long() & 0xFFFF
    diff = lw != lr
    if diff.any():                                                # only unfilled image corners may differ
        b, ys, xs_ = torch.where(diff)
        corner = ((ys == 0) | (ys == H - 1)) & ((xs_ == 0) | (xs_ == W - 1))
        assert corner.all() and (walk.winner[b, ys, xs_] < 0).all(), f"{int(diff.sum())} pixels of the map differ"
        assert int(diff.sum()) <= 4 * B
    pred = rp.synthetic_pred(B, C, seed=1).cuda()
    for zr in (False, True):
        sw, mw = ops.inverse_fill(walk, pred, want_scores=True, want_mask=True, zero_residual=zr)
        sr, mr = ops.inverse_fill(ras, pred, want_scores=True, want_mask=True, zero_residual=zr)
        assert torch.equal(torch.isnan(sw), torch.isnan(sr))
        assert torch.equal(torch.nan_to_num(sw), torch.nan_to_num(sr)) and torch.equal(mw, mr)


def test_raster_on_arbitrary_point_sets(ops, monkeypatch):
    """Interp2D on points without the corners: pixels outside the hull must come out NaN on both paths."""
    from fovea.interp2d import interp2d_scores
    rng = np.random.default_rng(3)
    pts = torch.from_numpy(np.unique(rng.integers(5, [60, 90], size=(400, 2)), axis=0)).cuda()
    vals = torch.randn(len(pts), 5, generator=torch.Generator().manual_seed(3)).cuda()
    outs = []
    for mode in ("walk", "raster"):
        monkeypatch.setenv("FOVEA_LOCATE", mode)
        outs.append(interp2d_scores(pts, vals, 64, 96, triangulation="device"))
    assert torch.isnan(outs[0]).any()
    assert torch.equal(torch.isnan(outs[0]), torch.isnan(outs[1]))
    assert torch.equal(torch.nan_to_num(outs[0]), torch.nan_to_num(outs[1]))


def test_width_not_multiple_of_8_uses_the_walker(ops):
    grid = _grid(1, seed=2)
    plan = ops.build_inverse_plan(grid, (100, 100), nchan=4, triangulation="device")
    assert plan.hints is not None
    s, _ = ops.inverse_fill(plan, rp.synthetic_pred(1, 4, seed=2).cuda(), want_scores=True)
    assert torch.isfinite(s).all()


@pytest.mark.parametrize("H,W,tri", [(256, 320, "device"), (1024, 1024, "device"), (4096, 4096, "device"), (520, 392, "host")])
def test_raster_variants_write_the_same_map(ops, monkeypatch, H, W, tri):
    """The three rasterisers of fovea_locate_raster -- span-start markers + row sweep (default), the per-pixel sweep and the
    closed-form row spans -- evaluate one predicate and must agree bit for bit (the marker path additionally relies on
    the spans of a row partitioning it, and on skipping spans made of vertex pixels only)."""
    grid = _grid(2, seed=3 * H + W).float().contiguous()          # (_locate_raster takes the raw pointer)
    plan = ops.build_inverse_plan(grid, (H, W), nchan=51, triangulation=tri)
    maps = {}
    for mode in ("64", "0", "8"):
        monkeypatch.setenv("FOVEA_RAS_MODE", mode)
        maps[mode] = ops._locate_raster(plan.pts, plan.mesh, plan.trirec, plan.ntri, grid, plan.winner, plan.h, plan.w,
                                        plan.cap, plan.tcap, False).clone()
    assert torch.equal(maps["64"], plan.loc)
    assert torch.equal(maps["64"], maps["0"]) and torch.equal(maps["8"], maps["0"])


def test_marker_raster_leaves_unmeshed_frames_unset(ops, monkeypatch):
    """A frame whose Delaunay run reported no mesh (ntri = 0) must come out "no value" everywhere (except its stamped nodes)."""
    grid = _grid(2, seed=11).float().contiguous()
    plan = ops.build_inverse_plan(grid, (256, 256), nchan=51, triangulation="device")
    ntri = plan.ntri.clone(); ntri[1] = 0
    loc = ops._locate_raster(plan.pts, plan.mesh, plan.trirec, ntri, grid, plan.winner, plan.h, plan.w, plan.cap, plan.tcap, False)
    l = loc.view(torch.int16).long() & 0xFFFF
    assert torch.equal(l[0], plan.loc.view(torch.int16).long()[0] & 0xFFFF)
    none = 0x8000 | (plan.h * plan.w)
    unset = plan.winner[1] < 0
    assert (l[1][unset] == none).all() and (l[1][~unset] != none).all()

"""The oracle's triangulation stand-in (stock SciPy Qhull) against the REFERENCE'S OWN vendored Qhull 2019.1
(oracle/_ref/libqhull_ref.so, built by oracle/Makefile from /root/reference/spatial/qhull_src): identical
simplices, orientation, order and neighbours on every point set the path produces -- so Delaunay tie-breaking among
co-circular lattice points is pinned to the reference's native code, not just to a newer Qhull."""
import os

import numpy as np
import pytest
import torch

from oracle import qhull_ref
from oracle import reference_port as rp

pytestmark = pytest.mark.skipif(not qhull_ref.available(), reason="oracle/_ref/libqhull_ref.so not built (make -C oracle)")


def _points_from_nan_tensor(t):
    mask, _ = rp.pixels_for_interp(t)
    return torch.stack(torch.where(mask[0]), 1).numpy().astype(np.float64)


def _same(p):
    simp, nb = qhull_ref.delaunay(p)
    tri = rp.delaunay(p)
    assert np.array_equal(simp, tri.simplices)
    assert np.array_equal(nb, tri.neighbors)
    return len(simp)


def test_reference_qhull_is_the_vendored_version():
    assert qhull_ref.version().startswith("2019.1")


@pytest.mark.parametrize("name", ["inverse_80_to_128", "inverse_80_to_520"])
def test_scipy_standin_equals_reference_qhull_on_golden_point_sets(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    for b in range(g["pred_sampled_nan"].shape[0]):
        assert _same(_points_from_nan_tensor(torch.from_numpy(g["pred_sampled_nan"][b]))) > 0


@pytest.mark.parametrize("H,W,seed", [(256, 320, 2), (1024, 1024, 3)])
def test_scipy_standin_equals_reference_qhull_on_foveated_point_sets(H, W, seed):
    """Full BASELINE geometry (80x80 nodes -> H x W canvas): the A7 winners + A9 selection of the oracle."""
    xs, _ = rp.synthetic_saliency(1, seed=seed)
    filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
    grid, _ = rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))
    gi = rp.grid_inverse(grid, (H, W), tie="max")
    canvas = torch.where(torch.isnan(gi[0, :, :, 0]), torch.nan, 1.0)[None]  # [1,H,W]: NaN where unfilled
    p = _points_from_nan_tensor(canvas)
    assert len(p) > 1000
    assert _same(p) > len(p)


def test_interp2d_smoke_points_of_the_reference():
    """interp2d.py:94-102 (the reference's own __main__ smoke): 10 random points + corners on a 64x48 canvas."""
    rng = np.random.RandomState(0)
    p = np.unique(np.concatenate([rng.randint(0, [64, 48], (10, 2)), [[0, 0], [0, 47], [63, 0], [63, 47]]]), axis=0)
    _same(p.astype(np.float64))


def _incircle_rc(P, a, b, c, d):
    ax, ay = P[a, 1] - P[d, 1], P[a, 0] - P[d, 0]
    bx, by = P[b, 1] - P[d, 1], P[b, 0] - P[d, 0]
    cx, cy = P[c, 1] - P[d, 1], P[c, 0] - P[d, 0]
    return ((ax * ax + ay * ay) * (bx * cy - by * cx) - (bx * bx + by * by) * (ax * cy - ay * cx)
            + (cx * cx + cy * cy) * (ax * by - ay * bx))


@pytest.mark.parametrize("H,W,seed", [(256, 320, 2), (1024, 1024, 3)])
def test_qt_splits_cocircular_cells_as_a_fan_from_the_last_inserted_vertex(H, W, seed):
    """WHY the device Delaunay kernel cannot reproduce Qhull inside co-circular cells: the
    reference's Qhull merges each set of co-circular sites into one facet and `Qt` re-triangulates it as a FAN around
    the vertex with the largest vertex id = the point its incremental hull happened to insert last -- a property of
    Qhull's global processing order (furthest-point selection over outside sets), not of the local geometry.  Asserted:
    every such cell is a fan, its apex is the largest-id vertex, every vertex of the cell lies exactly on one circle;
    and the apex is NOT predictable from the cell alone (no local rule picks it more often than chance allows)."""
    xs, _ = rp.synthetic_saliency(1, seed=seed)
    filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
    grid, _ = rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))
    gi = rp.grid_inverse(grid, (H, W), tie="max")
    canvas = torch.where(torch.isnan(gi[0, :, :, 0]), torch.nan, 1.0)[None]
    p = _points_from_nan_tensor(canvas)
    simp, owner, vid = qhull_ref.delaunay_probe(p)
    assert (vid >= 0).all()
    Pi = p.astype(np.int64)
    cells = {}
    for t in np.flatnonzero(owner != -1):
        cells.setdefault(int(owner[t]), []).append(t)
    assert len(cells) > 50
    hits = {"max_index": 0, "min_index": 0, "max_col": 0, "min_lift": 0}
    for ts in cells.values():
        verts = sorted(set(simp[ts].ravel().tolist()))
        assert len(ts) == len(verts) - 2                       # a triangulated polygon
        common = set(simp[ts[0]].tolist())
        for t in ts[1:]:
            common &= set(simp[t].tolist())
        apex = max(verts, key=lambda q: vid[q])
        assert apex in common, "cell is not a fan around its largest-id vertex"
        a, b, c = simp[ts[0]]
        for d in verts:                                        # every vertex on the circle of the first triangle
            assert d in (a, b, c) or _incircle_rc(Pi, a, b, c, d) == 0
        hits["max_index"] += apex == max(verts)
        hits["min_index"] += apex == min(verts)
        hits["max_col"] += apex == max(verts, key=lambda q: (Pi[q, 1], Pi[q, 0]))
        hits["min_lift"] += apex == min(verts, key=lambda q: (Pi[q] ** 2).sum())
    n = len(cells)
    # a 4-vertex cell gives a blind guess 1/4: none of the local rules is right even half of the time
    assert all(v < 0.5 * n for v in hits.values()), (hits, n)

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

"""DynamicFocus deformed_unsampler on the sm_100a kernels against the output of the unmodified
reference (tests/golden/unsampler_*.npz) and the CPU oracle.

Bar: integer work -- every pixel must receive the label of A nearest scattered pixel (exact squared distances, checked
against SciPy's EDT), scattered pixels their own label (largest node index where nodes collide, = the reference's CPU
behaviour), and every difference from the reference's output must be a tie between equidistant scattered pixels, where
SciPy's pick is an artefact of its scan order.
"""
import os

import numpy as np
import pytest
import torch

from oracle import reference_port as rp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def df():
    if not torch.cuda.is_available():
        pytest.fail("these tests need a CUDA device (run on the B200 box with -m gpu)")
    from fovea import dynamic_focus, ops
    ops._lib.load()
    return dynamic_focus


def _check(df, labels, coords, H, W, want=None):
    from fovea import ops
    B, K, HS, WS = labels.shape
    got = df.deformed_unsampler(labels.cuda(), coords.cuda(), H, W).cpu()
    ref, owner, dist = rp.deformed_unsampler(labels, coords, H, W, return_sites=True)
    if want is not None:
        assert np.array_equal(ref.numpy(), want)                       # the oracle reproduces the reference exactly
    # which node did every pixel receive?  (the product's own source map)
    winner = ops.scatter_nodes(coords.cuda(), (H, W))
    loc = ops.nearest_locate_all(winner, HS, WS).cpu().numpy().astype(np.int64) & 0x7FFF
    lab = labels.numpy().reshape(B, K, HS * WS)
    yy, xx = np.mgrid[0:H, 0:W]
    n_ties = 0
    for b in range(B):
        r, c = coords[b, 0].numpy().ravel(), coords[b, 1].numpy().ravel()
        node = loc[b]
        assert node.min() >= 0 and node.max() < HS * WS
        # value = the label of that node, bit for bit
        assert np.array_equal(got[b].numpy(), lab[b][:, node])
        # that node's pixel is a NEAREST scattered pixel: exact integer squared distance == EDT distance squared
        d2 = (r[node] - yy) ** 2 + (c[node] - xx) ** 2
        assert np.array_equal(d2, np.rint(dist[b] ** 2).astype(np.int64))
        # scattered pixels keep their own (largest-index) node, as the reference
        filled = owner[b][r, c]
        assert np.array_equal(node[r, c], filled)
        # every disagreement with the reference is therefore a tie between equidistant scattered pixels
        diff = node != owner[b]
        n_ties += int(diff.sum())
        assert np.array_equal(d2[diff], (r[owner[b][diff]] - yy[diff]) ** 2 + (c[owner[b][diff]] - xx[diff]) ** 2)
    agree = (got == ref).all(dim=1).float().mean().item()
    assert agree > 0.97, agree
    return n_ties, agree


@pytest.mark.parametrize("name", ["unsampler_24_to_96x128", "unsampler_40x64_to_520"])
def test_deformed_unsampler_matches_reference(df, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    H, W = int(g["H"]), int(g["W"])
    coords = df.int_rount_scale_grid(torch.from_numpy(g["grid"]).cuda(), H, W).cpu()
    assert np.array_equal(coords.numpy(), g["coords"])
    ties, agree = _check(df, torch.from_numpy(g["labels"]), coords, H, W, want=g["out"])
    print(f"{name}: {ties} tie pixels, {agree:.4f} of the pixels bit-identical to the reference")


def test_deformed_unsampler_full_size_properties(df):
    """1024^2 canvas, 80x80 lattice (the path's own sizes), K = 5: same exactness checks against the oracle."""
    gen = torch.Generator().manual_seed(5)
    B, K, HS, WS, H, W = 2, 5, 80, 80, 1024, 1024
    xs, _ = rp.synthetic_saliency(B, seed=5)
    filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
    grid, _ = rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))      # [B,80,80,2] (x, y)
    g2 = grid.permute(0, 3, 1, 2)[:, [1, 0]].contiguous()                                  # -> (rows, columns)
    coords = rp.int_round_scale_grid(g2.clone(), H, W)
    labels = torch.randn(B, K, HS, WS, generator=gen)
    _check(df, labels, coords, H, W)


def test_deformed_unsampler_sparse_lattice_tiled_search(df):
    """A 20 x 24 lattice on a 400 x 520 canvas (>= 400 pixels per node): fovea_nearest_locate_all runs its tiled search."""
    gen = torch.Generator().manual_seed(9)
    B, K, HS, WS, H, W = 2, 2, 20, 24, 400, 520
    coords = torch.stack([torch.randint(0, H, (B, HS, WS), generator=gen), torch.randint(0, W, (B, HS, WS), generator=gen)], 1)
    coords[0, :, :4, :] = coords[0, :, :1, :1]                       # many nodes on one pixel; a cluster of neighbours
    coords[1, 0] = (coords[1, 0] // 8).clamp(max=H - 1)              # everything in the top eighth: a far-away rest
    labels = torch.randn(B, K, HS, WS, generator=gen)
    assert H * W >= 400 * HS * WS
    _check(df, labels, coords, H, W)


def test_deformed_unsampler_edge_cases(df):
    # a single scattered pixel: the whole canvas takes its label; targets outside the canvas are dropped
    labels = torch.tensor([[[[3.0, 7.0]]]])
    coords = torch.tensor([[[[5, 400]], [[9, 2]]]])                    # node 0 -> (5,9); node 1 -> row 400: outside
    out = df.deformed_unsampler(labels.cuda(), coords.cuda(), 16, 24)
    assert out.shape == (1, 1, 16, 24) and bool((out == 3.0).all())
    from fovea import FoveaError
    with pytest.raises(FoveaError):
        df.deformed_unsampler(labels, coords, 16, 24)                  # CPU grid: no fallback

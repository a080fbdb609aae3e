"""Mask mode of stage 3 (fovea_inverse_mask: per-node argmax + per-triangle dominance pruning) must be BIT-IDENTICAL to
torch.argmax of the materialised scores (models/models.py:1044) -- including exact ties, which torch resolves to the first
maximum -- and to the all-channel fused argmax of fovea_inverse_fill it replaces."""
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


def _preds(kind, B, C, seed):
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(B, C, 80, 80, generator=g)
    if kind == "quantised":            # values on a coarse lattice: exact ties between channels at many vertices
        p = torch.round(p * 2) / 2
    elif kind == "few_levels":         # three levels only: almost every vertex comparison is a tie
        p = torch.randint(-1, 2, (B, C, 80, 80), generator=g).float()
    elif kind == "c1":                 # the reference's decoder tail: C-1 per-frame constants, one varying channel
        cls = torch.randn(B, C, generator=g)
        p = cls[:, :, None, None].expand(B, C, 80, 80).clone()
        p[:, -1:] = cls[:, -1:, None, None] * (torch.sigmoid(3 * torch.randn(B, 1, 80, 80, generator=g)) - 0.5)
    elif kind == "wide_range":         # magnitudes from 1e-6 to 1e6 in one frame
        p = p * torch.exp(14 * (torch.rand(B, C, 1, 1, generator=g) - 0.5))
    elif kind == "constant":           # all channels equal everywhere: the first channel must win every pixel
        p = torch.ones(B, C, 80, 80)
    elif kind == "smooth":             # spatially smooth logits: long runs of one winner, boundaries inside triangles
        p = torch.nn.functional.interpolate(torch.randn(B, C, 6, 6, generator=g), size=(80, 80), mode="bicubic")
    return p.cuda()


@pytest.mark.parametrize("kind", ["normal", "quantised", "few_levels", "c1", "wide_range", "constant", "smooth"])
@pytest.mark.parametrize("H,W,tri", [(256, 320, "device"), (1024, 1024, "device"), (520, 392, "host")])
def test_pruned_mask_equals_argmax_of_scores(ops, kind, H, W, tri):
    B, C = 2, 51
    plan = ops.build_inverse_plan(_grid(B, seed=H), (H, W), nchan=C, triangulation=tri)
    pred = _preds(kind, B, C, seed=W)
    scores, fused = ops.inverse_fill(plan, pred, want_scores=True, want_mask=True)       # all-channel fused argmax
    _, pruned = ops.inverse_fill(plan, pred, want_scores=False, want_mask=True)          # fovea_inverse_mask
    assert pruned.dtype == torch.int64
    want = torch.argmax(scores, dim=1)
    assert torch.equal(fused, want)
    n = int((pruned != want).sum())
    assert n == 0, f"{n} pixels differ from torch.argmax ({kind}, {H}x{W}, {tri} mesh)"
    _, pruned8 = ops.inverse_fill(plan, pred, want_scores=False, want_mask=True,
                                  mask_out=torch.empty(B, H, W, device="cuda", dtype=torch.uint8))
    assert torch.equal(pruned8.long(), want)


@pytest.mark.parametrize("C", [1, 2, 3, 7, 200])
def test_pruned_mask_channel_counts(ops, C):
    B, H, W = 2, 256, 256
    plan = ops.build_inverse_plan(_grid(B, seed=C), (H, W), nchan=C, triangulation="device")
    pred = _preds("quantised", B, C, seed=C)
    scores, _ = ops.inverse_fill(plan, pred, want_scores=True)
    _, pruned = ops.inverse_fill(plan, pred, want_scores=False, want_mask=True)
    assert torch.equal(pruned, torch.argmax(scores, dim=1))


def test_pruned_mask_unfilled_corners_and_nearest_plans(ops):
    """zero_residual=False: pixels in triangles with an unfilled-corner vertex are NaN in every channel -> class 0 (torch);
    'nearest' plans hold no triangles at all (every pixel is a direct table row)."""
    B, C, H, W = 2, 9, 256, 256
    grid = _grid(B, seed=3)
    pred = _preds("normal", B, C, seed=3)
    plan = ops.build_inverse_plan(grid, (H, W), nchan=C, triangulation="device")
    for zr in (False, True):
        scores, _ = ops.inverse_fill(plan, pred, want_scores=True, zero_residual=zr)
        _, pruned = ops.inverse_fill(plan, pred, want_scores=False, want_mask=True, zero_residual=zr)
        assert torch.equal(pruned, torch.argmax(scores, dim=1))
    nplan = ops.build_nearest_plan(grid, (H, W), nchan=C)
    scores, _ = ops.inverse_fill(nplan, pred, want_scores=True)
    _, pruned = ops.inverse_fill(nplan, pred, want_scores=False, want_mask=True)
    assert torch.equal(pruned, torch.argmax(scores, dim=1))


def test_pruned_mask_agrees_with_all_channel_kernel_at_bench_size(ops):
    """Size-independent property at BASELINE configs[1] geometry (a slice of the batch): the two mask kernels are the same
    function, so their outputs must be identical on the bench inputs too."""
    B, C, H, W = 4, 51, 1024, 1024
    plan = ops.build_inverse_plan(_grid(B, seed=11), (H, W), nchan=C, triangulation="device")
    pred = _preds("normal", B, C, seed=11)
    _, pruned = ops.inverse_fill(plan, pred, want_scores=False, want_mask=True)
    ops._FULL_MASK_FILL = True
    try:
        _, full = ops.inverse_fill(plan, pred, want_scores=False, want_mask=True)
    finally:
        ops._FULL_MASK_FILL = False
    assert torch.equal(pruned, full)

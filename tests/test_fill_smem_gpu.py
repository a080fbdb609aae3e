"""FOVEA_FILL_SMEM=1|2 (the scores kernel with the tile's table rows staged in shared memory, and with TMA tile stores on top,
csrc/inverse_smem.cu) must be bit-identical to the default fill kernel: same arithmetic, different path of the operands."""
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


@pytest.mark.parametrize("H,W,C,tri", [(256, 320, 7, "device"), (1024, 1024, 51, "device"), (520, 392, 5, "host"),
                                       (2048, 2048, 3, "device"), (128, 128, 51, "device")])
@pytest.mark.parametrize("zero_residual", [True, False])
def test_smem_fill_is_bit_identical(ops, monkeypatch, H, W, C, tri, zero_residual):
    B = 2
    xs, _ = rp.synthetic_saliency(B, seed=H + C)
    filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
    grid = rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))[0].cuda()
    pred = rp.synthetic_pred(B, C, seed=W).cuda()
    plan = ops.build_inverse_plan(grid, (H, W), nchan=C, triangulation=tri)
    monkeypatch.setenv("FOVEA_FILL_SMEM", "0")
    a, _ = ops.inverse_fill(plan, pred, want_scores=True, zero_residual=zero_residual)
    for mode in ("1", "2"):      # 1: table rows staged in shared memory; 2: + TMA tile stores
        monkeypatch.setenv("FOVEA_FILL_SMEM", mode)
        b = torch.full_like(a, 7.0)
        ops.inverse_fill(plan, pred, want_scores=True, zero_residual=zero_residual, out=b)
        assert torch.equal(torch.isnan(a), torch.isnan(b)), mode
        assert torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)), mode

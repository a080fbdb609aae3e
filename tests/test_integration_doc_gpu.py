"""The ctypes stub printed in INTEGRATION.md (Level 2) is executed as written and must reproduce fovea.ops."""
import os
import re

import pytest
import torch

from oracle import reference_port as rp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_level2_stub_runs_and_matches_ops():
    from fovea import ops
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "fovea_abi_version()" in b)
    ns = {}
    exec(stub.replace("<repo>", ROOT), ns)
    B, C, H, W, g, R = 2, 7, 192, 256, 80, 45
    xs, _ = rp.synthetic_saliency(B, seed=5)
    pred = rp.synthetic_pred(B, C, seed=5).cuda()
    x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(5)).cuda()
    g1x, g1y = (t.cuda() for t in ops.separable_factors(rp.gaussian_filter_weight(R, R, R)))
    grid, sums = ns["create_grid_fwd"](xs.cuda(), g1x, g1y, R)
    want_grid = ops.saliency_to_grid(xs.cuda(), g1x, g1y, g, g, R, R, "replication", (g, g))
    assert torch.equal(grid, want_grid)
    assert torch.equal(ns["grid_sample"](x, grid), ops.grid_sample(x, grid))
    scores, mask = ns["inverse_upsample"](pred, grid, H, W)
    plan = ops.build_inverse_plan(grid, (H, W), nchan=C, triangulation="device")
    want_scores, want_mask = ops.inverse_fill(plan, pred, want_scores=True, want_mask=True)
    torch.cuda.synchronize()
    assert torch.equal(mask, want_mask)
    assert torch.equal(scores, want_scores)
    # the stage-0 stub further down the document, executed in the same namespace
    exec(next(b for b in blocks if "def saliency_input" in b), ns)
    fp = torch.rand(B, 2, generator=torch.Generator().manual_seed(6)).cuda()
    assert torch.equal(ns["saliency_input"](x, fp), ops.saliency_input(x, fp, (80, 80)))

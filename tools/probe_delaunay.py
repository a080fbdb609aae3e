"""Phase breakdown of the device Delaunay kernel from its workspace counters (GPU box)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
from fovea import ops
from bench import make_inputs, WORKLOADS, Path
cfg = dict(WORKLOADS["b64_1024"]); B, C, H, W, g, R = (cfg[k] for k in "BCHWgR")
dev = torch.device("cuda", 0)
cfg["B"] = B
x, xs, pred = make_inputs(dict(cfg, H=64, W=64), 0, device=dev)
path = Path(cfg, dev, "device")
grid = ops.saliency_to_grid(xs, path.g1x, path.g1y, g, g, R, R, "replication", (g, g))
plan = ops.build_inverse_plan(grid, (H, W), nchan=C, triangulation="device")
mesh, ntri, ws = ops.delaunay_device(plan.pts, plan.npts, plan.cap, plan.tcap, max(H, W))
torch.cuda.synchronize()
ws = ws.cpu()
rounds = ws[:B]; dbg = ws[B:B + 8 * B].view(B, 8)
import numpy as np
d = dbg.numpy().astype(np.int64) * 16 / 1965.0  # us
print("npts  mean %.0f max %d" % (plan.npts.float().mean().item(), plan.npts.max().item()))
print("flip rounds: mean %.1f max %d" % (rounds.float().mean().item(), rounds.max().item()))
print("pocket rounds L/R: mean %.1f / %.1f" % (dbg[:, 4].float().mean().item(), dbg[:, 5].float().mean().item()))
for name, col_a, col_b in (("rows", None, 1), ("strips", 1, 2), ("pockets", 2, 3), ("flips", 3, 0)):
    a = d[:, col_a] if col_a is not None else 0
    seg = d[:, col_b] - a
    print(f"{name:8s} mean {seg.mean():8.1f} us   max {seg.max():8.1f} us")
print("total    mean %8.1f us   max %8.1f us" % (d[:, 0].mean(), d[:, 0].max()))

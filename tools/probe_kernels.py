"""A few launches of every stage-3 kernel on the bench batch (64 x 1024^2, C = 51): the target of `ncu -k regex:<kernel>`
captures (tools/round_measure.sh) -- forward fill (scores), mask mode (pruned), backward."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
import bench
from fovea import ops
cfg = dict(bench.WORKLOADS["b64_1024"]); dev = torch.device("cuda", 0)
x, xs, pred = bench.make_inputs(cfg, 0, device=dev)
path = bench.Path(cfg, dev, "device")
for _ in range(4):
    path.step(x, xs, pred)                                   # scores mode: the headline kernels
    path.step(x, xs, pred, want_scores=False, want_mask=True)  # mask mode: node_argmax, triangle_candidates, inverse_mask
grid = ops.saliency_to_grid(xs, path.g1x, path.g1y, 80, 80, 45, 45, "replication", (80, 80))
plan = ops.build_inverse_plan(grid, (cfg["H"], cfg["W"]), nchan=cfg["C"], triangulation="device")
p = pred.clone().requires_grad_(True)
gup = torch.randn_like(path.scores[:, :1]).expand_as(path.scores)     # (a 13.7 GB upstream gradient, read once per launch)
table = ops.box4_table(pred)
gt = torch.empty_like(table)
for i in range(5):
    if i == 1:
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
    ops._lib.call("fovea_inverse_fill_bwd", ops._ptr(plan.loc), ops._ptr(plan.trirec), ops._ptr(path.scores), cfg["B"], cfg["C"],
                  table.shape[2], plan.h, plan.w, plan.H, plan.W, plan.tcap, ops._ptr(gt), ops._stream())
b.record(); torch.cuda.synchronize()
print(f"fovea_inverse_fill_bwd: {a.elapsed_time(b) / 4:.2f} ms per launch (reads {path.scores.numel() * 4 / 1e9:.1f} GB)")
s, _ = ops.inverse_fill(plan, p, want_scores=True, out=path.scores)
s.backward(path.scores)
torch.cuda.synchronize()
print("ok")

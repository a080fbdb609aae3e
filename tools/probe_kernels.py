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
for _ in range(4):
    s, _ = ops.inverse_fill(plan, p, want_scores=True, out=path.scores)
    s.backward(path.scores)                                  # fovea_inverse_fill_bwd on a 13.7 GB upstream gradient
    p.grad = None
torch.cuda.synchronize()
print("ok")

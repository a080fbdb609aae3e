"""Timing of the 'nearest' reverse-deform mode against the 'tri' mode on the bench workload (GPU box)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
from fovea import ops
from bench import make_inputs, WORKLOADS, Path
cfg = dict(WORKLOADS["b64_1024"]); B, C, H, W, g, R = (cfg[k] for k in "BCHWgR")
dev = torch.device("cuda", 0)
x, xs, pred = make_inputs(dict(cfg, H=64, W=64), 0, device=dev)
path = Path(cfg, dev, "device")
grid = ops.saliency_to_grid(xs, path.g1x, path.g1y, g, g, R, R, "replication", (g, g))
table = ops.box4_table(pred)
def timeit(name, fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): out = fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name:44s} {a.elapsed_time(b)/n:8.3f} ms")
    return out
pn = timeit("build_nearest_plan (scatter+columns+rows)", lambda: ops.build_nearest_plan(grid, (H, W), C))
pt = timeit("build_inverse_plan 'tri' (device Delaunay)", lambda: ops.build_inverse_plan(grid, (H, W), nchan=C, triangulation="device"))
timeit("inverse_fill on the nearest plan", lambda: ops._fill(pn, table, C, True, path.scores, None))
timeit("inverse_fill on the tri plan", lambda: ops._fill(pt, table, C, True, path.scores, None))

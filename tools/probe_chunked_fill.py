"""Does running locate_pixels -> inverse_fill per chunk of frames keep `loc` in L2 for the fill? (GPU box)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
from fovea import ops, _lib
from fovea.ops import _ptr, _stream
from bench import make_inputs, WORKLOADS, Path
cfg = dict(WORKLOADS["b64_1024"]); B, C, H, W, g, R = (cfg[k] for k in "BCHWgR")
dev = torch.device("cuda", 0)
x, xs, pred = make_inputs(cfg, 0, device=dev)
path = Path(cfg, dev, "device")
grid = ops.saliency_to_grid(xs, path.g1x, path.g1y, g, g, R, R, "replication", (g, g))
plan = ops.build_inverse_plan(grid, (H, W), nchan=C, triangulation="device")
table = ops.box4_table(pred)
scores = path.scores

def run(chunk):
    for b0 in range(0, B, chunk):
        b1 = b0 + chunk
        _lib.call("fovea_locate_pixels", _ptr(plan.winner[b0:b1]), _ptr(plan.pts[b0:b1]), _ptr(plan.npts[b0:b1]),
                  _ptr(plan.mesh[b0:b1]), _ptr(plan.ntri[b0:b1]), _ptr(plan.hints[b0:b1]), chunk, g, g, H, W, plan.cap,
                  plan.tcap, _ptr(plan.loc[b0:b1]), _stream())
        _lib.call("fovea_inverse_fill", _ptr(plan.loc[b0:b1]), _ptr(plan.pts[b0:b1]), _ptr(plan.src[b0:b1]),
                  _ptr(plan.mesh[b0:b1]), _ptr(table[b0:b1]), chunk, C, table.shape[2], g, g, H, W, plan.cap, plan.tcap,
                  1, _ptr(scores[b0:b1]), None, _stream())

for chunk in (64, 32, 16, 8, 4, 2):
    run(chunk); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(chunk); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"chunk {chunk:3d}: locate+fill {best:.3f} ms")

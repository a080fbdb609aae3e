"""Time fovea_locate_raster on the bench batch (64 x 1024^2) -- run with FOVEA_RAS_TILE_MAX=<pixels> to move the boundary
between the per-pixel sweep (8 lanes per triangle) and the whole-warp row-span path."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
import bench
from fovea import ops
wl = sys.argv[1] if len(sys.argv) > 1 else "b64_1024"
cfg = dict(bench.WORKLOADS[wl]); dev = torch.device("cuda", 0)
x, xs, pred = bench.make_inputs(dict(cfg, H=64, W=64), 0, device=dev)
path = bench.Path(dict(cfg, H=64, W=64), dev, "device")
grid = ops.saliency_to_grid(xs, path.g1x, path.g1y, 80, 80, 45, 45, "replication", (80, 80))
plan = ops.build_inverse_plan(grid, (cfg["H"], cfg["W"]), nchan=cfg["C"], triangulation="device")
best = 1e9
for rep in range(6):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    loc = ops._locate_raster(plan.pts, plan.mesh, plan.trirec, plan.ntri, grid, plan.winner, plan.h, plan.w, plan.cap, plan.tcap, False)
    b.record(); torch.cuda.synchronize()
    if rep: best = min(best, a.elapsed_time(b))
print(wl, "FOVEA_RAS_TILE_MAX", os.environ.get("FOVEA_RAS_TILE_MAX", "default"), f"raster + stamp {best*1e3:.0f} us", "same map:", torch.equal(loc, plan.loc))

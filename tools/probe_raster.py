"""Time fovea_locate_raster on the bench batch (64 x 1024^2) -- run with FOVEA_RAS_TILE_MAX=<pixels> to move the boundary
between the small-triangle path and the whole-warp row-span path; RAS_MODES=0,4,8 lists the FOVEA_RAS_MODE values compared
(0 = per-pixel sweep, 64 = span-start markers + row sweep, n = row spans with n lanes per triangle)."""
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
ref = None
for mode in [m for m in os.environ.get("RAS_MODES", "0,64,8").split(",")]:
    os.environ["FOVEA_RAS_MODE"] = mode          # read per call by fovea_locate_raster
    best = 1e9
    for rep in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loc = ops._locate_raster(plan.pts, plan.mesh, plan.trirec, plan.ntri, grid, plan.winner, plan.h, plan.w, plan.cap, plan.tcap, False)
        b.record(); torch.cuda.synchronize()
        if rep: best = min(best, a.elapsed_time(b))
    if ref is None: ref = loc.clone()
    print(wl, "FOVEA_RAS_TILE_MAX", os.environ.get("FOVEA_RAS_TILE_MAX", "default"), "mode", mode,
          f"raster + stamp {best*1e3:.0f} us", "same map as mode 0:", torch.equal(loc, ref), flush=True)

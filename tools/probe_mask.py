"""Per-kernel view of stage 3 in mask mode (run under `ncu --metrics gpu__time_duration.sum` for the launch list):
fovea_inverse_mask (node_argmax + triangle_candidates + inverse_mask) vs the all-channel fused argmax, 64 x 1024^2."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
import bench
from fovea import ops
cfg = dict(bench.WORKLOADS["b64_1024"])
dev = torch.device("cuda", 0)
x, xs, pred = bench.make_inputs(cfg, 0, device=dev)
path = bench.Path(cfg, dev, "device")
grid = ops.saliency_to_grid(xs, path.g1x, path.g1y, 80, 80, 45, 45, "replication", (80, 80))
plan = ops.build_inverse_plan(grid, (cfg["H"], cfg["W"]), nchan=cfg["C"], triangulation="device")
tab = ops.box4_table(pred)
ncand = None
for full in (False, True):
    ops._FULL_MASK_FILL = full
    for _ in range(3):
        ops._fill(plan, tab, cfg["C"], True, None, path.mask)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ops._fill(plan, tab, cfg["C"], True, None, path.mask)
    b.record(); torch.cuda.synchronize()
    print("all-channel" if full else "pruned", a.elapsed_time(b) / 5, "ms")
# survivor statistics
B, tcap = plan.loc.shape[0], plan.tcap
nbytes = int(ops._lib.load().fovea_inverse_mask_workspace_bytes(B, plan.h, plan.w, tcap))
ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
ops._lib.call("fovea_inverse_mask", ops._ptr(plan.loc), ops._ptr(plan.trirec), ops._ptr(plan.ntri), ops._ptr(tab), B, cfg["C"],
              tab.shape[2], plan.h, plan.w, plan.H, plan.W, tcap, ops._ptr(ws), ops._ptr(path.mask), 0, ops._stream())
torch.cuda.synchronize()
nc = ws[B * tcap * 512: B * tcap * 512 + B * tcap].view(B, tcap)
T = plan.ntri.long()
vals = torch.cat([nc[b, :T[b]] for b in range(B)]).float()
print("survivors per triangle: mean", vals[vals < 254].mean().item(), "overflow frac", (vals == 255).float().mean().item(),
      "nan frac", (vals == 254).float().mean().item(), "p99", vals[vals < 254].quantile(0.99).item())

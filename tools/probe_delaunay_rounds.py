"""Per-round clock / dirty-count trace of the Delaunay flip phase (needs a -DDT_PROFILE build; GPU box)."""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
from fovea import ops, _lib
from fovea.ops import _ptr, _stream
from bench import make_inputs, WORKLOADS, Path
cfg = dict(WORKLOADS["b64_1024"]); B, C, H, W, g, R = (cfg[k] for k in "BCHWgR")
dev = torch.device("cuda", 0)
x, xs, pred = make_inputs(dict(cfg, H=64, W=64), 0, device=dev)
path = Path(cfg, dev, "device")
grid = ops.saliency_to_grid(xs, path.g1x, path.g1y, g, g, R, R, "replication", (g, g))
plan = ops.build_inverse_plan(grid, (H, W), nchan=C, triangulation="device")
nbytes = int(_lib.load().fovea_delaunay_workspace_bytes(B, plan.cap))
ws = torch.zeros((nbytes + 3) // 4 + B * 512 * 2 + 4096, device=dev, dtype=torch.int32)
# the row-start scratch of the kernel sits right after the counters: keep the profile area behind everything
mesh = torch.empty(B, plan.tcap, 8, device=dev, dtype=torch.uint16); ntri = torch.empty(B, device=dev, dtype=torch.int32)
_lib.call("fovea_delaunay", _ptr(plan.pts), _ptr(plan.npts), B, plan.cap, plan.tcap, max(H, W), _ptr(mesh), _ptr(ntri), _ptr(ws), _stream())
torch.cuda.synchronize()
w = ws.cpu().numpy()
rounds = w[:B]
b = int(np.argmax(rounds))
o = 9 * B + B * (plan.cap + 2) // 2
prof = w[o: o + B * 1024].reshape(B, 512, 2)
print("image", b, "rounds", rounds[b])
t = prof[b, :rounds[b], 0].astype(np.int64) * 16 / 1965.0
d = prof[b, :rounds[b], 1]
for r in list(range(0, min(40, rounds[b]))) + list(range(40, rounds[b], 10)):
    dt = (t[r + 1] - t[r]) if r + 1 < len(t) else float("nan")
    print(f"round {r:4d}  t={t[r]:8.1f} us  dirty={d[r]:6d}  round_time={dt:6.2f} us")

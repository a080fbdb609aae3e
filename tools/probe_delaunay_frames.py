"""Per-frame view of the Delaunay kernel on the bench batch: sites, flip rounds and the clock at the end of each phase
(the kernel's always-on debug counters; clock64 >> 4 ticks at 1965 MHz).  With a -DDT_PROFILE build (FOVEA_B200_LIB) also
the per-round trace of the worst frame."""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
from fovea import ops, _lib
from fovea.ops import _ptr, _stream
from bench import make_inputs, WORKLOADS, Path
wl = sys.argv[1] if len(sys.argv) > 1 else "b64_1024"
cfg = dict(WORKLOADS[wl]); B, C, H, W, g, R = (cfg[k] for k in "BCHWgR")
dev = torch.device("cuda", 0)
x, xs, pred = make_inputs(dict(cfg, H=64, W=64), 0, device=dev)
path = Path(dict(cfg, H=64, W=64), dev, "device")
grid = ops.saliency_to_grid(xs, path.g1x, path.g1y, g, g, R, R, "replication", (g, g))
plan = ops.build_inverse_plan(grid, (H, W), nchan=C, triangulation="device")
nbytes = int(_lib.load().fovea_delaunay_workspace_bytes(B, plan.cap))
prof = os.environ.get("DT_PROFILE") == "1"
ws = torch.zeros((nbytes + 3) // 4 + (B * 512 * 2 + 4096 if prof else 0), device=dev, dtype=torch.int32)
mesh = torch.empty(B, plan.tcap, 8, device=dev, dtype=torch.uint16); ntri = torch.empty(B, device=dev, dtype=torch.int32)
for _ in range(3):
    ws.zero_()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.call("fovea_delaunay", _ptr(plan.pts), _ptr(plan.npts), B, plan.cap, plan.tcap, max(H, W), _ptr(mesh), _ptr(ntri), _ptr(ws), _stream())
    b_.record(); torch.cuda.synchronize()
print(wl, "kernel", a.elapsed_time(b_), "ms")
w = ws.cpu().numpy()
rounds, dbg, npts = w[:B], w[B:9 * B].reshape(B, 8), plan.npts.cpu().numpy()
us = lambda v: v.astype(np.int64) * 16 / 1965.0
tot, rows, strips, pockets = us(dbg[:, 0]), us(dbg[:, 1]), us(dbg[:, 2]), us(dbg[:, 3])
order = np.argsort(-tot)
print("frame  sites rounds  rows  strips pockets  flips  total [us]")
for b in list(order[:6]) + list(order[-3:]):
    print(f"{b:5d} {npts[b]:6d} {rounds[b]:6d} {rows[b]:6.1f} {strips[b]-rows[b]:6.1f} {pockets[b]-strips[b]:7.1f} {tot[b]-pockets[b]:7.1f} {tot[b]:7.1f}")
print(f"mean: rounds {rounds.mean():.1f}  construction {pockets.mean():.1f} us  flips {(tot-pockets).mean():.1f} us  total {tot.mean():.1f} us;"
      f"  sum over frames {tot.sum()/1e3:.2f} SM-ms; max {tot.max():.1f} us")
print(f"pocket rounds (ear-clipping rounds of the left / right side): mean {dbg[:, 4].mean():.1f} / {dbg[:, 5].mean():.1f}, max {dbg[:, 4].max()} / {dbg[:, 5].max()};"
      f"  pockets {(pockets - strips).mean():.1f} us on average")
if prof:
    b = int(np.argmax(rounds))
    o = 9 * B + B * (plan.cap + 2) // 2
    p = w[o: o + B * 1024].reshape(B, 512, 2)
    t, d = us(p[b, :rounds[b], 0]), p[b, :rounds[b], 1]
    print("worst frame", b, "rounds", rounds[b])
    for r in list(range(0, min(30, rounds[b]), 2)) + list(range(30, min(rounds[b], 512), 15)):
        dt = (t[r + 1] - t[r]) if r + 1 < len(t) else float("nan")
        print(f"round {r:4d}  t={t[r]:8.1f} us  dirty={d[r]:6d}  round_time={dt:6.2f} us")

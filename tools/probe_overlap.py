"""Cross-step software pipelining: plan(i+1) on a high-priority stream overlapping fill(i) on another (GPU box)."""
import os, sys, torch, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
from fovea import ops
from bench import make_inputs, WORKLOADS, Path
wl = sys.argv[1] if len(sys.argv) > 1 else "b64_1024"
cfg = dict(WORKLOADS[wl]); B, C, H, W, g, R = (cfg[k] for k in "BCHWgR")
dev = torch.device("cuda", 0)
x, xs, pred = make_inputs(cfg, 0, device=dev)
path = Path(cfg, dev, "device")
K = 20

def serial():
    for _ in range(K):
        path.step(x, xs, pred)

lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
sp = torch.cuda.Stream(priority=-1)   # plan: high priority
sf = torch.cuda.Stream(priority=0)    # fill
def pipelined(depth=2):
    live = collections.deque()
    cur = torch.cuda.current_stream()
    sp.wait_stream(cur); sf.wait_stream(cur)
    for i in range(K):
        with torch.cuda.stream(sp):
            if len(live) >= depth:
                sp.wait_event(live[0][2])          # bound the run-ahead (and the memory held by plans)
            grid = ops.saliency_to_grid(xs, path.g1x, path.g1y, g, g, R, R, "replication", (g, g))
            xsamp = ops.grid_sample(x, grid)
            plan = ops.build_inverse_plan(grid, (H, W), nchan=C, triangulation="device")
            ev = torch.cuda.Event(); ev.record(sp)
        with torch.cuda.stream(sf):
            sf.wait_event(ev)
            table = ops.box4_table(pred)
            ops._fill(plan, table, C, True, path.scores, None)
            done = torch.cuda.Event(); done.record(sf)
        live.append((plan, grid, done, xsamp, table))
        if len(live) > depth:
            live.popleft()
    cur.wait_stream(sp); cur.wait_stream(sf)
    return live

for name, fn in (("serial", serial), ("pipelined d2", lambda: pipelined(2)), ("pipelined d3", lambda: pipelined(3))):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); keep = fn(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    print(f"{wl} {name:14s} {ms:7.3f} ms/step  {B/ms*1e3:9.0f} frames/s  path HBM frac {4.0*C*H*W*B/ms/1e6/6542.4:.3f}")

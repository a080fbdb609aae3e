"""Times the stages of ResamplePipeline in isolation (GPU box): PCIe copies, the host-gather grid_sample, the kernels."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
from fovea import ops
from bench import make_inputs, WORKLOADS, Path

cfg = dict(WORKLOADS["b64_1024"]); B, C, H, W, g, R = (cfg[k] for k in "BCHWgR")
dev = torch.device("cuda", 0)
hx, hxs, hpred = make_inputs(cfg, 1, pinned=True)
hmask = torch.empty(B, H, W, dtype=torch.int64, pin_memory=True)
x = torch.empty(B, 3, H, W, device=dev); mask = torch.empty(B, H, W, dtype=torch.int64, device=dev)
path = Path(cfg, dev, "device")
xs, pred = hxs.to(dev), hpred.to(dev)

def timeit(name, fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name:40s} {a.elapsed_time(b)/n:8.3f} ms")

timeit("H2D image 805 MB", lambda: x.copy_(hx, non_blocking=True))
timeit("D2H mask int64 537 MB", lambda: hmask.copy_(mask, non_blocking=True))
grid = ops.saliency_to_grid(xs, path.g1x, path.g1y, g, g, R, R, "replication", (g, g))
timeit("grid_sample from HBM", lambda: ops.grid_sample(x, grid))
timeit("grid_sample gather from pinned host", lambda: ops.grid_sample(hx, grid))
timeit("path (scores)", lambda: path.step(x, xs, pred))
timeit("path (scores+mask)", lambda: path.step(x, xs, pred, want_scores=True, want_mask=True))
timeit("path (mask only)", lambda: path.step(x, xs, pred, want_scores=False, want_mask=True))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): x.copy_(hx, non_blocking=True)
    with torch.cuda.stream(s2): hmask.copy_(mask, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
timeit("H2D image || D2H mask", both)
def gather_and_d2h():
    with torch.cuda.stream(s1): ops.grid_sample(hx, grid)
    with torch.cuda.stream(s2): hmask.copy_(mask, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
timeit("gather || D2H mask", gather_and_d2h)

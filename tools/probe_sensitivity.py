"""What does each stage cost the PIPELINED step?  Every C-ABI call of the plan is idempotent, so issuing one of them
twice adds exactly its own work: the growth of the DevicePipeline step (64 x 1024^2, scores mode) is that stage's
marginal cost under overlap with the fill -- the number that says where the next optimisation pays (GPU box)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
import bench
from fovea import ops, _lib
from fovea.pipeline import DevicePipeline
wl = sys.argv[1] if len(sys.argv) > 1 else "b64_1024"
cfg = dict(bench.WORKLOADS[wl]); dev = torch.device("cuda", 0)
B, C, H, W = (cfg[k] for k in "BCHW")
x, xs, pred = bench.make_inputs(cfg, 0, device=dev)
scores = torch.empty(B, C, H, W, device=dev)
pipe = DevicePipeline(B, C, H, W, cfg["g"], cfg["R"], dev, "device", depth=2, scores=scores)
real_call = _lib.call
dup = set()
seen = []
def call(name, *a):
    if name not in seen: seen.append(name)
    r = real_call(name, *a)
    if name in dup: real_call(name, *a)
    return r
_lib.call = call; ops._lib.call = call
def measure(n=20 if wl == "b64_1024" else 16):
    for _ in range(4): pipe.submit(x, xs, pred)
    pipe.fence(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): pipe.submit(x, xs, pred)
    pipe.fence(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
base = measure()
print(f"baseline {base:.3f} ms per step; calls per step: {seen}")
for name in list(seen):
    dup.clear(); dup.add(name)
    t = measure()
    print(f"  {name:34s} twice: {t:.3f} ms  (+{(t - base) * 1e3:6.0f} us = {(t - base) * 148:5.1f} SM-ms)", flush=True)
dup.clear()
print(f"baseline again {measure():.3f} ms")

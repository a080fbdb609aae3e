"""How does inverse_fill's time depend on the SMs it gets?  A blocker kernel (tools/probes/sm_blocker.cu) holds n SMs
for 6 ms on a side stream while the fill of the bench batch runs on the main stream (GPU box)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
import bench
from fovea import ops
lib = ctypes.CDLL(os.path.join(ROOT, "tools", "_build", "libsm_blocker.so"))
lib.sm_blocker.argtypes = [ctypes.c_int, ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p]
cfg = dict(bench.WORKLOADS["b64_1024"]); dev = torch.device("cuda", 0)
x, xs, pred = bench.make_inputs(cfg, 0, device=dev)
path = bench.Path(cfg, dev, "device")
for _ in range(3):
    path.step(x, xs, pred)
grid = ops.saliency_to_grid(xs, path.g1x, path.g1y, 80, 80, 45, 45, "replication", (80, 80))
plan = ops.build_inverse_plan(grid, (cfg["H"], cfg["W"]), nchan=cfg["C"], triangulation="device")
sink = torch.zeros(4, device=dev, dtype=torch.int32)
side = torch.cuda.Stream(priority=-1)
for n in (0, 8, 16, 32, 48, 64, 96, 128):
    ts = []
    for rep in range(3):
        torch.cuda.synchronize()
        if n:
            lib.sm_blocker(n, 6_000_000, sink.data_ptr(), side.cuda_stream)
            torch.cuda._sleep(200_000)       # ~0.1 ms: let the blocker's CTAs get resident first
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.inverse_fill(plan, pred, want_scores=True, out=path.scores)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"blocked SMs {n:3d}: fill {min(ts):.3f} ms  (148-n)/148 model {2.28 * 148 / (148 - n):.3f} ms", flush=True)

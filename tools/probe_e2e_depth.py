"""e2e (host buffers -> int64 masks on the host) of ResamplePipeline for several pipeline depths and host-buffer counts."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
from bench import make_inputs, WORKLOADS
from fovea.pipeline import ResamplePipeline
cfg = dict(WORKLOADS["b64_1024"]); B, C, H, W = cfg["B"], cfg["C"], cfg["H"], cfg["W"]
dev = torch.device("cuda", 0)
nbuf = 4
host = [make_inputs(cfg, seed=100 + i, pinned=True) for i in range(nbuf)]
hmask = [torch.empty(B, H, W, dtype=torch.int64, pin_memory=True) for _ in range(nbuf)]
for depth in (2, 3, 4):
    for want_scores in (True, False):
        pipe = ResamplePipeline(B, C, H, W, cfg["g"], cfg["R"], dev, "device", depth=depth, image_on_host=True,
                                want_scores=want_scores)
        for i in range(4):
            pipe.submit(*host[i % nbuf], hmask[i % nbuf])
        pipe.drain(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k = 12
        e0.record()
        for i in range(k):
            pipe.submit(*host[i % nbuf], hmask[i % nbuf])
        pipe.fence(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / k
        print(f"depth {depth} scores={want_scores}: {ms:.2f} ms per 64 frames = {B / ms * 1e3:.0f} frames/s", flush=True)
        del pipe; torch.cuda.empty_cache()

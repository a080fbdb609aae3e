"""Write-only ceiling of the fill kernel's store pattern, with and without a 4 B/pixel side read (GPU box)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
from fovea import ops
B, C, H, W = 64, 51, 1024, 1024
scores = torch.empty(B, C, H, W, device="cuda")
side = torch.zeros(B, H, W, device="cuda", dtype=torch.int32)
for name, sr in (("stores only", None), ("stores + 4 B/px read", side)):
    best = 1e9
    for i in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.probe_store_ceiling(scores, sr); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"{name:24s} {best:.3f} ms  {4.0*B*C*H*W/best/1e6:.0f} GB/s")

"""Store-ceiling probe, direct 128-bit stores vs shared memory + TMA tile stores (FOVEA_PROBE_TMA = number of stages).
Each variant runs in its own process (the switch is read once); prints time, GB/s and a checksum of the output."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
    from fovea import ops
    B, C, H, W = 64, 51, int(sys.argv[2]), int(sys.argv[2])
    if H >= 2048:
        B = 16
    scores = torch.zeros(B, C, H, W, device="cuda")
    side = torch.zeros(B, H, W, device="cuda", dtype=torch.int32)
    for name, sr in (("stores only", None), ("stores + loc read", side)):
        best = 1e9
        for i in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.probe_store_ceiling(scores, sr); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        chk = scores[::7, ::5].double().sum().item() + scores[-1, -1, -1].double().sum().item()
        print(f"  {name:20s} {best:.3f} ms  {4.0*B*C*H*W/best/1e6:.0f} GB/s  checksum {chk:.1f}")
else:
    for size in (1024, 2048):
        for st in ("0", "2", "3"):
            print(f"canvas {size}^2, FOVEA_PROBE_TMA={st}", flush=True)
            subprocess.run([sys.executable, __file__, "child", str(size)], env=dict(os.environ, FOVEA_PROBE_TMA=st))

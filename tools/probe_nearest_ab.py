"""A/B of the two exact nearest-site implementations (GPU box): tiled (default) vs column/row scans (FOVEA_NEAREST_TILES=0);
checks that their source maps agree pixel for pixel (same tie rule) and prints the plan times."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
    from fovea import ops
    from bench import make_inputs, WORKLOADS, Path
    for wl in ("b64_1024", "b64_2048"):
        cfg = dict(WORKLOADS[wl]); B, C, H, W, g, R = (cfg[k] for k in "BCHWgR")
        if H > 1024: B = 16
        dev = torch.device("cuda", 0)
        x, xs, pred = make_inputs(dict(cfg, B=B, H=64, W=64), 0, device=dev)
        path = Path(dict(cfg, B=1, H=64, W=64), dev, "device")
        grid = ops.saliency_to_grid(xs, path.g1x, path.g1y, g, g, R, R, "replication", (g, g))
        winner = ops.grid_inv_scatter(grid, (H, W))
        for name, fn in (("nearest_locate (reference site rule)", lambda: ops.nearest_locate(winner, g, g, C)),
                         ("nearest_locate_all (every filled pixel)", lambda: ops.nearest_locate_all(winner, g, g))):
            loc = fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): loc = fn()
            b.record(); torch.cuda.synchronize()
            print(f"  {wl} B={B} {name:42s} {a.elapsed_time(b)/5:7.3f} ms  checksum {int(loc.long().sum())} "
                  f"{int((loc.long() * torch.arange(loc.numel(), device=dev).view_as(loc) % 1000003).sum())}")
else:
    for v in ("1", "0"):
        print(f"FOVEA_NEAREST_TILES={v}", flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, FOVEA_NEAREST_TILES=v))

"""Delaunay kernel with and without the single-warp tail (DT_TAIL): time per 64 frames, flip rounds, and a checksum of the
mesh (the tail runs the same rounds, so the meshes must be bit-identical).  Children load the library named by
FOVEA_B200_LIB.  The comparison library is not built by the Makefile:
    cd foveated-instance-segmentation_b200/csrc && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
        -I../../include -DDT_TAIL=0 -c delaunay.cu -o /tmp/dt0.o && nvcc -gencode arch=compute_100a,code=sm_100a -shared \
        -o ../../tools/_libfovea_notail.so common.o saliency.o grid.o grid_sample.o inverse.o nearest.o /tmp/dt0.o -lcudart"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
    from fovea import ops
    from bench import make_inputs, WORKLOADS, Path
    for wl in ("b64_1024", "b64_2048"):
        cfg = dict(WORKLOADS[wl]); B, C, H, W, g, R = (cfg[k] for k in "BCHWgR")
        dev = torch.device("cuda", 0)
        x, xs, pred = make_inputs(dict(cfg, H=64, W=64), 0, device=dev)
        path = Path(dict(cfg, B=1, H=64, W=64), dev, "device")
        grid = ops.saliency_to_grid(xs, path.g1x, path.g1y, g, g, R, R, "replication", (g, g))
        plan = ops.build_inverse_plan(grid, (H, W), nchan=C, triangulation="device")
        best = 1e9
        for i in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); mesh, ntri, ws = ops.delaunay_device(plan.pts, plan.npts, plan.cap, plan.tcap, max(H, W)); b.record()
            torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
        rounds = ws[:B].cpu()
        chk = 0
        for bb in range(B):
            m = mesh[bb, : int(ntri[bb])].to(torch.int64)
            chk += int((m * torch.arange(1, m.numel() + 1, device=dev).view_as(m) % 1000003).sum())
        print(f"  {wl}: {best:.3f} ms per {B} frames, rounds mean {rounds.float().mean():.1f} max {int(rounds.max())}, "
              f"triangles {int(ntri.sum())}, mesh checksum {chk}", flush=True)
else:
    for name, lib in (("tail (default build)", None), ("no tail (-DDT_TAIL=0)", os.path.join(ROOT, "tools", "_libfovea_notail.so"))):
        print(name, flush=True)
        env = dict(os.environ)
        if lib and not os.path.exists(lib):
            print("  (not built: see the docstring)")
            continue
        if lib: env["FOVEA_B200_LIB"] = lib
        subprocess.run([sys.executable, __file__, "child"], env=env)

"""Randomised stress of the stage-3 pipeline (GPU box): many seeds / sizes / saliency shapes; every device mesh is
validated exactly (tests/test_delaunay_gpu.check_mesh) and the device-mode scores are compared with the host-Qhull mode."""
import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200"), os.path.join(ROOT, "tests")]
from fovea import ops
from oracle import reference_port as rp
from test_delaunay_gpu import check_mesh
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.RandomState(123)
filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
worst = 1.0
t0 = time.time()
for case in range(n_cases):
    H = int(rng.choice([96, 200, 256, 512, 640, 1024, 1536])); W = int(rng.choice([96, 204, 256, 512, 640, 1024, 2048]))
    B = int(rng.randint(1, 4)); C = int(rng.choice([2, 5, 51]))
    g = torch.Generator().manual_seed(int(rng.randint(1 << 30)))
    sharp, noise, sig = float(rng.uniform(0, 12)), float(rng.uniform(0, 5)), float(rng.uniform(2, 30))
    gaze = torch.rand(B, 2, generator=g)
    ii = torch.arange(80.)[None, :, None]; jj = torch.arange(80.)[None, None, :]
    d2 = (ii - gaze[:, 0, None, None] * 79) ** 2 + (jj - gaze[:, 1, None, None] * 79) ** 2
    xs = torch.softmax((noise * torch.randn(B, 80, 80, generator=g) + sharp * torch.exp(-d2 / (2 * sig ** 2))).view(B, -1), 1).view(B, 1, 80, 80)
    grid, _ = rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))
    pred = torch.randn(B, C, 80, 80, generator=g).cuda()
    pd = ops.build_inverse_plan(grid.cuda(), (H, W), nchan=C, triangulation="device")
    ph = ops.build_inverse_plan(grid.cuda(), (H, W), nchan=C, triangulation="host")
    sd, md = ops.inverse_fill(pd, pred, want_scores=True, want_mask=True)
    sh, mh = ops.inverse_fill(ph, pred, want_scores=True, want_mask=True)
    pn = ops.build_nearest_plan(grid.cuda(), (H, W), nchan=C); sn, _ = ops.inverse_fill(pn, pred)
    torch.cuda.synchronize()
    assert torch.isfinite(sd).all() and torch.isfinite(sh).all() and torch.isfinite(sn).all()
    npts = pd.npts.cpu().numpy(); pts = pd.pts.cpu().numpy(); mesh = pd.mesh.cpu().numpy(); ntri = pd.ntri.cpu().numpy()
    for b in range(B):
        p = pts[b, :npts[b]]
        check_mesh(np.stack([p >> 16, p & 0xFFFF], 1), mesh[b], int(ntri[b]))
    close = ((sd - sh).abs() <= 1e-4 * sh.abs().max()).all(1).float().mean().item()
    worst = min(worst, close)
    print(f"case {case:3d} B={B} C={C:2d} {H}x{W} sharp={sharp:4.1f} noise={noise:3.1f} sig={sig:4.1f} npts={npts.tolist()} "
          f"device==host scores on {close:.4f} of the pixels", flush=True)
print(f"stress ok: {n_cases} cases in {time.time()-t0:.1f} s, worst device/host score agreement {worst:.4f}")

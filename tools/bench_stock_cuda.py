"""Inference hot path on STOCK PyTorch CUDA ops (the reference's own formulation) beside this repository's kernels, on the
same B200 and the same inputs ("the real competitor since no reference Blackwell kernel exists").

    python tools/bench_stock_cuda.py [--batch 64] [--size 1024] [--steps 10]

Stock arm = models/models.py as written, device parts only:
    A4  three dense 91x91 F.conv2d on the padded saliency + quotient/clamp (:594-637)
    A5  F.grid_sample(x, grid) (:909)
    A7  NaN canvas + index_put of the low-res indices (:640-655)
    A8  F.grid_sample(pred, grid_inv) + NaN overwrite (:935-938)
    A10 torch.argmax (:1044)
A9 (fillMissingValues_tensor: Qhull + find_simplex on the HOST, ~0.3 s per 1024^2 frame, bench.py's cpu_baseline) is
NOT in the stock arm -- it has no device implementation in the reference -- so the stock number is a lower bound of what
the reference needs; ours includes A9 (device Delaunay + point location + interpolation).  Prints one JSON line.
"""
import argparse, json, os, sys, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
from fovea import ops
from fovea.models import makeGaussian

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--classes", type=int, default=51)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
B, C, H, W, g, R = args.batch, args.classes, args.size, args.size, 80, 45
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.rand(B, 3, H, W, device=dev, generator=gen)
pred = torch.randn(B, C, g, g, device=dev, generator=gen)
ii = torch.arange(g, device=dev, dtype=torch.float32)
gaze = torch.rand(B, 2, device=dev, generator=gen) * (g - 1)
d2 = (ii[None, :, None] - gaze[:, 0, None, None]) ** 2 + (ii[None, None, :] - gaze[:, 1, None, None]) ** 2
xs = torch.softmax((3 * torch.randn(B, g, g, device=dev, generator=gen) + 6 * torch.exp(-d2 / 128)).view(B, -1), 1)
xs = xs.view(B, 1, g, g)
filt = torch.from_numpy(makeGaussian(2 * R + 1, fwhm=R)).float().to(dev)
g1x, g1y = (t.to(dev) for t in ops.separable_factors(filt))
conv_w = filt.view(1, 1, 2 * R + 1, 2 * R + 1)
G = g + 2 * R
jj = (torch.arange(G, device=dev, dtype=torch.float32) - R) / (g - 1.0)
P = torch.stack([jj[None, :].expand(G, -1), jj[:, None].expand(-1, G)])[None]


def stock():
    xs_hm = F.pad(xs, (R, R, R, R), mode="replicate")
    den = F.conv2d(xs_hm, conv_w)
    num = F.conv2d((P * xs_hm).view(-1, 1, G, G), conv_w).view(B, 2, g, g)
    grid = (num / den * 2 - 1).clamp(-1, 1).permute(0, 2, 3, 1).contiguous()
    x_sampled = F.grid_sample(x, grid, align_corners=False)
    gi = torch.full((2, B, H, W), float("nan"), device=dev)
    u = ((grid[..., 0] + 1) / 2 * (W - 1)).int().long()
    v = ((grid[..., 1] + 1) / 2 * (H - 1)).int().long()
    bb = torch.arange(B, device=dev)[:, None, None].expand(B, g, g)
    gi[0][bb, v, u] = ii[None, None, :].expand(B, g, g)
    gi[1][bb, v, u] = ii[None, :, None].expand(B, g, g)
    gi[0] = gi[0] / g * 2 - 1
    gi[1] = gi[1] / g * 2 - 1
    grid_inv = gi.permute(1, 2, 3, 0)
    unfilled = torch.isnan(grid_inv[..., 0])
    ps = F.grid_sample(pred, torch.nan_to_num(grid_inv, nan=0.0), align_corners=False)
    ps.masked_fill_(unfilled[:, None], float("nan"))
    return x_sampled, torch.argmax(torch.nan_to_num(ps, nan=0.0), dim=1)


def ours():
    grid = ops.saliency_to_grid(xs, g1x, g1y, g, g, R, R, "replication", (g, g))
    x_sampled = ops.grid_sample(x, grid)
    plan = ops.build_inverse_plan(grid, (H, W), nchan=C, triangulation="device")
    scores, mask = ops.inverse_fill(plan, pred, want_scores=True, want_mask=True)
    return x_sampled, mask


def time_it(fn):
    for _ in range(2):
        out = fn()
    del out
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = fn()
        del out
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps


ms_ours = time_it(ours)
torch.cuda.empty_cache()
ms_stock = time_it(stock)
print(json.dumps({"workload": f"B={B} {H}x{W} C={C}", "ours_ms_per_step": ms_ours, "ours_frames_s": B / ms_ours * 1e3,
                  "stock_cuda_ms_per_step_without_A9": ms_stock, "stock_cuda_frames_s_without_A9": B / ms_stock * 1e3,
                  "note": "stock = the reference's device ops only (A4,A5,A7,A8,A10); its A9 runs on the host "
                          "(see cpu_baseline in bench.py) and is excluded; ours = the whole path incl. A9, scores + "
                          "mask, one stream, allocations inside the step"}))

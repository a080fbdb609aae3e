"""Training-step hot path (BASELINE.json configs[3]): saliency net -> S1 grid -> S2 grid_sample (image + label) ->
backward through S2 (grad w.r.t. the grid) and S1 -> saliency-net gradients -> ONE flat NCCL all-reduce.

    python tools/bench_train_step.py [--batch 32] [--size 1024]            # 1 GPU
    torchrun --nproc-per-node N tools/bench_train_step.py                  # N GPUs (weak scaling, 32 frames / GPU)

Times, with CUDA events, (a) this repository's kernels and (b) the reference's formulation on stock PyTorch CUDA ops
(three dense 91x91 conv2d + F.grid_sample + autograd) on the same inputs, and checks the two saliency gradients agree.
The encoder/decoder are not part of the path and are replaced by a fixed random projection of x_sampled.
"""
import argparse, json, os, sys, torch
import torch.distributed as dist
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
from types import SimpleNamespace as NS
from fovea import ops
from fovea.models import CompressNet, makeGaussian
from fovea.parallel import FlatGradBucket
from fovea.saliency_network import fov_simple

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B, H, W, g, R = args.batch, args.size, args.size, 80, 45
cfg = NS(MODEL=NS(saliency_net="fovsimple", fov_deform=True))
torch.manual_seed(0)
sal, comp = fov_simple(cfg).to(dev), CompressNet(cfg).to(dev)
gen = torch.Generator(device=dev).manual_seed(1 + rank)
x = torch.rand(B, 3, H, W, device=dev, generator=gen)
y = (torch.rand(B, 1, H, W, device=dev, generator=gen) > 0.5).float()
x_low5 = torch.rand(B, 5, g, g, device=dev, generator=gen)
proj = torch.randn(3, g, g, device=dev, generator=gen)
filt = torch.from_numpy(makeGaussian(2 * R + 1, fwhm=R)).float().to(dev)
g1x, g1y = (t.to(dev) for t in ops.separable_factors(filt))
conv_w = filt.view(1, 1, 2 * R + 1, 2 * R + 1).clone().requires_grad_(True)     # reference: filter.weight requires grad
ii = torch.arange(g + 2 * R, device=dev, dtype=torch.float64)
P = torch.stack([((ii - R) / (g - 1.0))[None, :].expand(g + 2 * R, -1), ((ii - R) / (g - 1.0))[:, None].expand(-1, g + 2 * R)]).float()
bucket = FlatGradBucket([sal, comp])

def saliency():
    xs = comp(sal(x_low5))
    return torch.softmax(xs.view(B, -1), dim=1).view(B, 1, g, g)

def step_ours():
    xs = saliency()
    grid = ops.saliency_to_grid(xs, g1x, g1y, g, g, R, R, "replication", (g, g))
    xsamp = ops.grid_sample(x, grid)
    ysamp = ops.grid_sample(y, grid)
    loss = (xsamp * proj).sum() / B + ysamp.mean()
    loss.backward()

def step_stock():
    xs = saliency()
    xs_hm = F.pad(xs, (R, R, R, R), mode="replicate")                      # models/models.py:821
    den = F.conv2d(xs_hm, conv_w)                                          # :602-607
    num = F.conv2d((xs_hm * P[None]).view(-1, 1, g + 2 * R, g + 2 * R), conv_w).view(B, 2, g, g)
    grid = torch.clamp(num / den * 2 - 1, -1, 1).permute(0, 2, 3, 1)       # :609-637
    xsamp = F.grid_sample(x, grid, align_corners=False)                    # :909
    ysamp = F.grid_sample(y, grid, align_corners=False)                    # :880
    loss = (xsamp * proj).sum() / B + ysamp.mean()
    loss.backward()

def grads():
    return torch.cat([p.grad.flatten() for p in bucket.params])

def zero():
    for p in bucket.params:
        p.grad = None
    conv_w.grad = None

def timeit(fn, allreduce):
    for _ in range(3):
        zero(); fn()
        if allreduce and world > 1: bucket.allreduce()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        zero(); fn()
        if allreduce and world > 1: bucket.allreduce()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / args.steps], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

# Gradient agreement of the hot path (d loss / d xs), judged by an fp64 evaluation of the reference formulation on a
# 4-frame slice with smooth images.  (The saliency-NETWORK gradients are not compared: softmax backward subtracts the
# weighted mean of d/dxs, which amplifies the ~1e-4 fp32 rounding of either implementation to O(10 %).)
def hot_grad(dtype, n=4):
    xs = torch.softmax(torch.randn(n, g * g, device=dev, generator=gen) * 3, 1).view(n, 1, g, g).to(dtype).requires_grad_(True)
    xsm = F.interpolate(F.avg_pool2d(x[:n], 32), size=(H, W), mode="bicubic", align_corners=False)
    if dtype is None:
        xs = xs.float().detach().requires_grad_(True)
        grid = ops.saliency_to_grid(xs, g1x, g1y, g, g, R, R, "replication", (g, g))
        (ops.grid_sample(xsm, grid) * proj).sum().backward()
        return xs.grad
    wd = filt.to(dtype).view(1, 1, 2 * R + 1, 2 * R + 1)
    xs_hm = F.pad(xs, (R, R, R, R), mode="replicate")
    den = F.conv2d(xs_hm, wd)
    num = F.conv2d((xs_hm * P.to(dtype)[None]).view(-1, 1, g + 2 * R, g + 2 * R), wd).view(n, 2, g, g)
    grid = torch.clamp(num / den * 2 - 1, -1, 1).permute(0, 2, 3, 1)
    (F.grid_sample(xsm.to(dtype), grid, align_corners=False) * proj.to(dtype)).sum().backward()
    return xs.grad

state = gen.get_state()
g64 = hot_grad(torch.float64); gen.set_state(state)
g32 = hot_grad(torch.float32); gen.set_state(state)
gours = hot_grad(None)
rel = float((gours.double() - g64).norm() / g64.norm())
rel_stock = float((g32.double() - g64).norm() / g64.norm())
ms_ours = timeit(step_ours, True)
ms_stock = timeit(step_stock, True)
# the hot path alone (no saliency net): kernels vs stock ops, forward + backward to grad_xs
xs0 = saliency().detach()
def hot_ours():
    xs = xs0.clone().requires_grad_(True)
    grid = ops.saliency_to_grid(xs, g1x, g1y, g, g, R, R, "replication", (g, g))
    ((ops.grid_sample(x, grid) * proj).sum() / B + ops.grid_sample(y, grid).mean()).backward()
def hot_stock():
    xs = xs0.clone().requires_grad_(True)
    xs_hm = F.pad(xs, (R, R, R, R), mode="replicate")
    den = F.conv2d(xs_hm, conv_w)
    num = F.conv2d((xs_hm * P[None]).view(-1, 1, g + 2 * R, g + 2 * R), conv_w).view(B, 2, g, g)
    grid = torch.clamp(num / den * 2 - 1, -1, 1).permute(0, 2, 3, 1)
    ((F.grid_sample(x, grid, align_corners=False) * proj).sum() / B + F.grid_sample(y, grid, align_corners=False).mean()).backward()
ms_hot_ours, ms_hot_stock = timeit(hot_ours, False), timeit(hot_stock, False)
if rank == 0:
    print(json.dumps({"what": "training-step hot path (S1+S2 fwd+bwd + saliency net + flat all-reduce)", "n_gpus": world,
                      "frames_per_gpu": B, "size": H, "ours_ms_per_step": ms_ours, "stock_torch_cuda_ms_per_step": ms_stock,
                      "ours_frames_s": world * B / ms_ours * 1e3, "stock_frames_s": world * B / ms_stock * 1e3,
                      "hot_path_only_ours_ms": ms_hot_ours, "hot_path_only_stock_ms": ms_hot_stock,
                      "grad_xs_rel_err_vs_fp64": {"ours": rel, "stock_fp32": rel_stock}, "allreduce_numel": bucket.numel}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()

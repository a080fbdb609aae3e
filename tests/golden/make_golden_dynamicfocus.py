"""Golden vectors of DynamicFocus's deformed_unsampler from the UNMODIFIED reference source (build container only).

    python tests/golden/make_golden_dynamicfocus.py

DynamicFocus/d_model/nn_B0_deformed_sampler.py cannot be imported as a module here (its header imports plotting and
image-loading helpers that need matplotlib and cv2 GUI pieces), so the two function definitions this path uses --
`int_rount_scale_grid` (:83-102) and `deformed_unsampler` (:115-153) -- are taken from the file's own syntax tree and
executed unmodified, with exactly the names their bodies use (torch, F, scipy's distance_transform_edt).
"""
import ast
import os

import numpy as np
import torch
import torch.nn.functional as F
from scipy.ndimage import distance_transform_edt

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/DynamicFocus/d_model/nn_B0_deformed_sampler.py"


def load_reference_functions():
    tree = ast.parse(open(SRC).read(), SRC)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("int_rount_scale_grid", "deformed_unsampler")]
    assert len(keep) == 2
    ns = {"torch": torch, "F": F, "distance_transform_edt": distance_transform_edt}
    exec(compile(ast.Module(body=keep, type_ignores=[]), SRC, "exec"), ns)
    return ns["int_rount_scale_grid"], ns["deformed_unsampler"]


def deformed_grid(B, HS, WS, seed):
    """A foveated lattice in [-1,1] (channel 0 rows, 1 columns): denser around a random gaze point."""
    gen = torch.Generator().manual_seed(seed)
    gaze = torch.rand(B, 2, generator=gen) * 1.2 - 0.6
    u = torch.linspace(-1, 1, HS)[None, :, None].expand(B, HS, WS)
    v = torch.linspace(-1, 1, WS)[None, None, :].expand(B, HS, WS)

    def warp(t, c):
        d = t - c[:, None, None]
        return c[:, None, None] + d * (0.15 + 0.85 * d.abs() / (1 + c[:, None, None].abs()))
    g = torch.stack([warp(u, gaze[:, 0]), warp(v, gaze[:, 1])], 1)
    return (g + 0.004 * torch.randn(g.shape, generator=gen)).clamp(-1, 1)


def case(name, B, K, HS, WS, H, W, seed):
    scale, unsample = load_reference_functions()
    gen = torch.Generator().manual_seed(seed + 100)
    grid = deformed_grid(B, HS, WS, seed)
    labels = torch.randn(B, K, HS, WS, generator=gen)
    coords = scale(grid.clone(), H, W)
    out = unsample(labels.clone(), coords.clone(), H, W)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), grid=grid.numpy(), labels=labels.numpy(), coords=coords.numpy(),
                        out=out.numpy(), H=np.array(H), W=np.array(W))
    flat = coords[:, 0] * W + coords[:, 1]
    dup = sum(int(flat[b].numel() - flat[b].unique().numel()) for b in range(B))
    print(name, tuple(out.shape), "nodes sharing a pixel:", dup)


if __name__ == "__main__":
    case("unsampler_24_to_96x128", B=2, K=3, HS=24, WS=24, H=96, W=128, seed=1)
    case("unsampler_40x64_to_520", B=2, K=2, HS=40, WS=64, H=520, W=520, seed=2)

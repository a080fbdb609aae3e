"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py --functions     # create_grid / grid_inv / fill / Interp2D fixtures
    python tests/golden/make_golden.py --module        # DeformSegmentationModule.forward fixtures

Imports /root/reference through ref_shim.py, runs the reference's own `create_grid`, `F.grid_sample`
call pattern, NaN-mask inverse sampling and `fillMissingValues_tensor('tri')` (with the reference's own
`interp2d.Interp2D`) on seeded synthetic inputs, and stores inputs + outputs as small .npz fixtures.
The GPU box has no /root/reference: tests only read the .npz files.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_shim  # noqa: E402

rm, ri = ref_shim.install_all()
from oracle.reference_port import synthetic_saliency, synthetic_pred  # noqa: E402  (input generators only)


def build(cfg):
    torch.manual_seed(0)
    return rm.DeformSegmentationModule(None, None, None, None, None, cfg)


def pad(m, xs, mode):
    # models/models.py:819-825 verbatim call pattern
    px, py = m.padding_size_x, m.padding_size_y
    if mode == "replication":
        return nn.ReplicationPad2d((py, py, px, px))(xs)
    if mode == "reflect":
        return F.pad(xs, (py, py, px, px), mode="reflect")
    return F.pad(xs, (py, py, px, px), mode="constant")


def case_grid(name, sal, task, R, segSize, B, seed, pad_mode="replication", task_eval=(), rate=1):
    cfg = ref_shim.make_cfg(sal=sal, task=task, R=R, pad=pad_mode, task_eval=task_eval, rate=rate)
    m = build(cfg)
    xs, gaze = synthetic_saliency(B, m.grid_size_x, m.grid_size_y, seed=seed)
    xs_hm = pad(m, xs, pad_mode)
    with torch.no_grad():
        grid, grid_y = m.create_grid(xs_hm)
        grid2, grid_inv = m.create_grid(xs_hm, segSize=segSize, x_inv=1 - xs_hm)
    out = dict(xs=xs.numpy(), gaze=gaze.numpy(), filt=m.filter.weight.detach()[0, 0].numpy(),
               P_basis=m.P_basis.numpy(), grid=grid.numpy(), grid_y=grid_y.numpy(), grid_infer=grid2.numpy(),
               grid_inv=grid_inv.numpy(), segSize=np.array(segSize), sal=np.array(sal), task=np.array(task),
               task_eval=np.array(task_eval, dtype=np.int64), rate=np.array(rate), R=np.array(R),
               Rx=np.array(m.padding_size_x), Ry=np.array(m.padding_size_y), pad_mode=np.array(pad_mode))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items() if hasattr(v, "shape") and v.ndim > 1})
    return m, xs, grid, grid_inv


def case_inverse(name, m, grid, grid_inv, C, seed, image_hw):
    """models/models.py:935-940 call pattern: grid_sample(pred, grid_inv) -> NaN mask -> per-sample tri fill."""
    B, h, w, _ = grid.shape
    pred = synthetic_pred(B, C, h, w, seed=seed)
    gi = grid_inv.clone()
    unfilled = torch.isnan(gi[:, :, :, 0])
    gi[torch.isnan(gi)] = 0
    ps = F.grid_sample(pred, gi.float())
    ps[unfilled.unsqueeze(1).expand(ps.shape)] = float("nan")
    ps_nan = ps.clone()
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        for n in range(B):
            ps[n] = rm.fillMissingValues_tensor(ps[n], interp_mode="tri")
    # grid_sample of an image with the same grid (A5), through the reference's own call
    g = torch.Generator().manual_seed(seed + 7)
    x = torch.rand(B, 3, image_hw[0], image_hw[1], generator=g)
    x_sampled = F.grid_sample(x, grid)
    out = dict(pred=pred.numpy(), grid=grid.numpy(), pred_sampled_nan=ps_nan.numpy(), pred_sampled=ps.numpy(),
               x_seed=np.array(seed + 7), x_hw=np.array(image_hw), x_sampled=x_sampled.numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, ps.shape, "nan left:", int(torch.isnan(ps).sum()))


def case_interp2d(name, h, w, N, vdim, seed):
    """interp2d.py: the reference's own Interp2D class on random integer points + the 4 corners."""
    g = torch.Generator().manual_seed(seed)
    lin = torch.randperm(h * w, generator=g)[:N]
    corners = torch.tensor([0, w - 1, (h - 1) * w, h * w - 1])
    lin = torch.unique(torch.cat([lin, corners]))  # sorted = row-major, as torch.where would give
    pts = torch.stack([lin // w, lin % w], 1)
    vals = torch.randn(pts.shape[0], vdim, generator=g)
    out = ri.Interp2D(h, w)(pts, vals)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), points=pts.numpy(), values=vals.numpy(), out=out.numpy(),
                        hw=np.array([h, w]))
    print(name, out.shape)


if __name__ == "__main__" and "--functions" in sys.argv:
    torch.set_num_threads(8)
    # BASELINE geometry: saliency 80x80, task 80x80, R=45, replication pad
    m, xs, grid, gi = case_grid("grid_80_R45", (80, 80), (80, 80), 45, (128, 128), B=2, seed=1)
    case_inverse("inverse_80_to_128", m, grid, gi, C=5, seed=1, image_hw=(256, 256))
    # > 512 px: exercises the nearest-downscaled dilation branch (models.py:183-193), non-integer ratio
    m, xs, grid, gi = case_grid("grid_80_R45_seg520", (80, 80), (80, 80), 45, (520, 536), B=1, seed=2)
    case_inverse("inverse_80_to_520", m, grid, gi, C=2, seed=2, image_hw=(520, 536))
    # anisotropic saliency (gaussian_ap = 2 -> Ry = 2*Rx, bilinear-resized Gaussian), task != saliency size,
    # label grid at rate 2, reflect / zero padding
    case_grid("grid_40x80_R12_reflect", (40, 80), (64, 96), 12, (96, 160), B=2, seed=3, pad_mode="reflect", rate=2)
    case_grid("grid_40x80_R12_zero", (40, 80), (64, 96), 12, (96, 160), B=2, seed=4, pad_mode="zero", rate=2)
    case_grid("grid_32_R10_eval", (32, 32), (32, 32), 10, (64, 64), B=3, seed=5, task_eval=(48, 48))
    case_interp2d("interp2d_64x48", 64, 48, 300, 4, seed=6)


# ---------------------------------------------------------------------------------------------------------------
# Module-level golden: the reference's own DeformSegmentationModule.forward (models/models.py:666-1094) on CPU with
# its real saliency net (saliency_network.fov_simple) + CompressNet, tiny stand-in encoder/decoder, deform.yaml
# configuration.  pytorch_toolbelt is absent, so its DiceLoss('multiclass') is the restatement in fovea.models
# (loss values are outside the hot path; they are stored as a sanity check only).
# ---------------------------------------------------------------------------------------------------------------
def case_module(name, upsample, H=256, W=256, B=2, seed=11):
    import importlib, yaml, io, contextlib
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "foveated-instance-segmentation_b200"))
    sys.path.insert(0, os.path.dirname(HERE))
    from fovea.models import MulticlassDiceLoss
    from tiny_nets import TinyEncoder, TinyDecoder, synthetic_batch
    rm.DiceLoss = lambda *a, **k: MulticlassDiceLoss()
    defaults = importlib.import_module("config.defaults")
    cfg = defaults._C

    def merge(dst, src):
        for k, v in src.items():
            if isinstance(v, dict):
                merge(dst[k], v)
            elif isinstance(v, str) and not isinstance(dst.get(k), str):
                import ast
                dst[k] = ast.literal_eval(v)      # yacs decodes "(80,80)"-style strings the same way
            else:
                dst[k] = v
    merge(cfg, yaml.safe_load(open(os.path.join(ref_shim.REF, "config", "deform.yaml"))))
    # README.md:73/79 command-line overrides of the shipped recipe
    cfg.TRAIN.task_input_size = (80, 80)
    cfg.TRAIN.saliency_input_size = (80, 80)
    cfg.MODEL.gaussian_radius = 45
    cfg.TRAIN.deform_joint_loss = True
    cfg.VAL.no_upsample = True
    cfg.DATASET.grid_path = ""
    cfg.MODEL.upsample = upsample
    cfg.MODEL.rev_deform_interp = "tri"
    cfg.TRAIN.global_epoch = 1
    cfg.DIR = "/tmp/fovea_golden"
    cfg.TRAIN.num_gpus = 1
    for k in ("saliency_input_size", "task_input_size", "task_input_size_eval", "dynamic_task_input"):
        cfg.TRAIN[k] = tuple(cfg.TRAIN[k])
    torch.manual_seed(seed)
    import saliency_network as rs
    sal, comp = rs.fov_simple(cfg), rm.CompressNet(cfg)
    for net in (sal, comp):
        net.apply(rm.ModelBuilder.weights_init)
    enc, dec = TinyEncoder(), TinyDecoder(num_class=cfg.DATASET.num_class)
    m = rm.DeformSegmentationModule(enc, dec, sal, comp, None, cfg)
    m.eval()
    captured = {}
    sal.register_forward_pre_hook(lambda mod, inp: captured.__setitem__("x_low", inp[0].detach().clone()))
    enc.register_forward_pre_hook(lambda mod, inp: captured.__setitem__("x_sampled", inp[0].detach().clone()))
    orig_acc = m.pixel_acc
    m.pixel_acc = lambda p, l: (captured.__setitem__("scored", p.detach().clone()), orig_acc(p, l))[1]
    feed = synthetic_batch(B, H, W, seed)
    feed_in = {k: v.clone() for k, v in feed.items()}
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        out = m(feed, is_inference=True, rank=1, cur_iter=-1)
    names = ["loss", "acc", "edge_loss", "acc_bin_fg", "acc_cls_fbg", "acc_bin_fbg"]
    res = {n: np.array(float(v)) for n, v in zip(names, out)}
    sd = {}
    for tag, net in (("sal", sal), ("comp", comp), ("enc", enc), ("dec", dec)):
        for k, v in net.state_dict().items():
            sd[f"sd_{tag}__{k}"] = v.numpy()
    arrays = dict(H=np.array(H), W=np.array(W), upsample=np.array(upsample), seed=np.array(seed), B=np.array(B),
                  x_low=captured["x_low"].numpy(), x_sampled=captured["x_sampled"].numpy(),
                  seg_label_after=feed["seg_label"].numpy(), **res, **sd)
    if upsample:  # scores at full resolution would be 2*51*H*W floats: keep the argmax mask + a strided sample
        sc = captured["scored"]
        arrays["scored_argmax"] = torch.max(torch.nan_to_num(sc, nan=0.0), 1)[1].to(torch.uint8).numpy()
        arrays["scored_nan"] = np.packbits(torch.isnan(sc[:, 0]).numpy())
        arrays["scored_sub"] = sc[:, ::5, ::4, ::4].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print(name, res)


if __name__ == "__main__" and "--module" in sys.argv:
    case_module("module_256_lowres", upsample=False)
    case_module("module_256_upsample", upsample=True)


def case_nearest(name, src_name):
    """fillMissingValues_tensor(..., interp_mode='nearest') of the UNMODIFIED reference (models/models.py:213-272: cv2.dilate
    site selection + SciPy NearestNDInterpolator on the host) on the NaN-masked tensors of an existing inverse fixture."""
    src = np.load(os.path.join(HERE, src_name + ".npz"))
    ps_nan = torch.from_numpy(src["pred_sampled_nan"])
    out = ps_nan.clone()
    for n in range(out.shape[0]):
        out[n] = rm.fillMissingValues_tensor(out[n], interp_mode="nearest")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), pred_sampled_nearest=out.numpy(), source=np.array(src_name))
    print(name, out.shape, "nan left:", int(torch.isnan(out).sum()))


if __name__ == "__main__" and "--nearest" in sys.argv:
    case_nearest("nearest_80_to_128", "inverse_80_to_128")
    case_nearest("nearest_80_to_520", "inverse_80_to_520")

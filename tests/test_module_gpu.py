"""Drop-in module test: fovea.models.DeformSegmentationModule against the reference's own forward
(tests/golden/module_256_*.npz, produced by tests/golden/make_golden.py --module on CPU with identical weights).

Tolerances: x_low (stock ops) 1e-5.  x_sampled is the image sampled at the *grid*, and the reference's fp32 91x91
convolution puts its grid ~2e-5 from the exact value (tests/test_parity_gpu.py); on a U[0,1) noise image of width W
that moves samples by ~2e-5 * W/2 pixels, so |x_sampled - ref| is bounded by ~1e-4 * W * |d img/dx| -- asserted as a
small median and a bounded maximum.  Losses / accuracies are sanity checks (stock PyTorch code outside the path).
"""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from tiny_nets import TinyDecoder, TinyEncoder, synthetic_batch

pytestmark = pytest.mark.gpu

# the stock saliency / encoder convolutions must run in true fp32 for a parity check (cuDNN defaults to TF32)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def make_cfg(upsample):
    return NS(
        TRAIN=NS(saliency_input_size=(80, 80), task_input_size=(80, 80), task_input_size_eval=(), dynamic_task_input=(1,),
                 def_saliency_pad_mode="replication", deform_joint_loss=True, opt_deform_LabelEdge_norm=True,
                 opt_deform_LabelEdge=False, edge_loss_scale=100.0, global_epoch=1, num_gpus=1),
        MODEL=NS(saliency_output_size_short=0, gaussian_radius=45, gaussian_ap=0.0, upsample=upsample,
                 rev_deform_interp="tri", uniform_sample="", gt_gradient=False, loss_at_high_res=False,
                 saliency_net="fovsimple"),
        DATASET=NS(segm_downsampling_rate=1, num_class=51),
        VAL=NS(no_upsample=True),
    )


def build(golden, upsample, triangulation):
    from fovea.models import CompressNet, DeformSegmentationModule
    from fovea.saliency_network import fov_simple
    cfg = make_cfg(upsample)
    nets = {"sal": fov_simple(cfg), "comp": CompressNet(cfg), "enc": TinyEncoder(), "dec": TinyDecoder(num_class=51)}
    for tag, net in nets.items():
        sd = {k[len(f"sd_{tag}__"):]: torch.from_numpy(v) for k, v in golden.items() if k.startswith(f"sd_{tag}__")}
        # strict=False as the reference's ModelBuilder (models/models.py:1213-1216): its SynchronizedBatchNorm2d keeps
        # extra bookkeeping buffers (_tmp_running_mean, ...); every real parameter / running stat must be present
        res = net.load_state_dict(sd, strict=False)
        assert all(k.endswith("num_batches_tracked") for k in res.missing_keys), res.missing_keys
    m = DeformSegmentationModule(nets["enc"], nets["dec"], nets["sal"], nets["comp"], None, cfg,
                                 triangulation=triangulation)
    return m.cuda().eval(), nets


@pytest.mark.parametrize("upsample,triangulation", [(False, "device"), (True, "host"), (True, "device")])
def test_module_forward_matches_reference(golden_dir, upsample, triangulation):
    name = "module_256_upsample" if upsample else "module_256_lowres"
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    m, nets = build(g, upsample, triangulation)
    H, W, B = int(g["H"]), int(g["W"]), int(g["B"])
    feed = {k: v.cuda() for k, v in synthetic_batch(B, H, W, int(g["seed"])).items()}
    cap = {}
    nets["sal"].register_forward_pre_hook(lambda mod, inp: cap.__setitem__("x_low", inp[0].detach()))
    nets["enc"].register_forward_pre_hook(lambda mod, inp: cap.__setitem__("x_sampled", inp[0].detach()))
    orig = m.pixel_acc
    m.pixel_acc = lambda p, l: (cap.__setitem__("scored", p.detach()), orig(p, l))[1]
    from fovea import ops
    from oracle import reference_port as rp
    calls, stock = [], ops.grid_sample
    ops.grid_sample = lambda inp, grid: (calls.append((inp.detach().cpu(), grid.detach().cpu())), stock(inp, grid))[1]
    try:
        with torch.no_grad():
            out = m(feed, is_inference=True, rank=1, cur_iter=-1)
    finally:
        ops.grid_sample = stock
    assert len(out) == 6
    np.testing.assert_allclose(cap["x_low"].cpu().numpy(), g["x_low"], rtol=0, atol=1e-5)
    err = np.abs(cap["x_sampled"].cpu().numpy() - g["x_sampled"])
    assert np.median(err) < 2e-4 and err.max() < 1e-4 * W, (np.median(err), err.max())
    # feed_dict['seg_label'] is mutated like the reference (models/models.py:951): y_sampled.long().  Truncating a
    # bilinear sample of a {0,1} disc is a rounding lottery in the reference itself (the four weights of a pixel
    # whose taps are all 1 sum to 1.0 or 1-ulp: the golden has ~2 % holes inside the disc), so the label is pinned
    # twice: bit-exact against the oracle's F.grid_sample on the grid this module produced, and within the lottery
    # rate against the reference's label on its own (CPU-convolution) grid.
    lab = feed["seg_label"].cpu().numpy()
    (y_in, grid_y) = calls[0]
    assert y_in.shape[1] == 1
    want = rp.grid_sample(y_in, grid_y).squeeze(1).long().numpy()
    assert np.array_equal(lab, want)
    assert (lab != g["seg_label_after"]).mean() < 2e-2
    for k, v in zip(["loss", "acc", "edge_loss", "acc_bin_fg", "acc_cls_fbg", "acc_bin_fbg"], out):
        assert abs(float(v) - float(g[k])) <= 2e-2 * max(1.0, abs(float(g[k]))), (k, float(v), float(g[k]))
    if upsample:
        sc = cap["scored"]
        assert sc.shape == (B, 51, H, W)
        mask = torch.max(torch.nan_to_num(sc, nan=0.0), 1)[1].cpu().numpy()
        agree = (mask == g["scored_argmax"]).mean()
        nan_ref = np.unpackbits(g["scored_nan"])[: B * H * W].reshape(B, H, W).astype(bool)
        nan_agree = (torch.isnan(sc[:, 0]).cpu().numpy() == nan_ref).mean()
        print(f"upsample/{triangulation}: full-res mask agreement {agree:.4f}, NaN-pattern agreement {nan_agree:.4f}")
        # measured: 0.9999 / 0.9999 with the reference's own mesh (host Qhull), 0.9996 / 0.9999 with the device mesh (the
        # remainder: the reference's undefined winners at collision pixels, and co-circular cells for the device mesh)
        assert agree > (0.999 if triangulation == "host" else 0.998) and nan_agree > 0.999


def test_module_training_step_backward():
    """Gradients reach the saliency network through grid_sample -> create_grid (autograd.Functions on the kernels)."""
    from fovea.models import CompressNet, DeformSegmentationModule
    from fovea.saliency_network import fov_simple
    cfg = make_cfg(False)
    torch.manual_seed(0)
    sal, comp, enc, dec = fov_simple(cfg), CompressNet(cfg), TinyEncoder(), TinyDecoder()
    m = DeformSegmentationModule(enc, dec, sal, comp, None, cfg).cuda().train()
    feed = {k: v.cuda() for k, v in synthetic_batch(4, 192, 192, 5).items()}
    loss, acc, edge = m(feed, rank=1, cur_iter=-1)
    loss.backward()
    g_sal = torch.cat([p.grad.flatten() for p in sal.parameters() if p.grad is not None])
    g_enc = torch.cat([p.grad.flatten() for p in enc.parameters()])
    assert torch.isfinite(g_sal).all() and g_sal.abs().sum() > 0
    assert torch.isfinite(g_enc).all() and g_enc.abs().sum() > 0
    assert "filter.weight" in m.state_dict() and m.state_dict()["filter.weight"].shape == (1, 1, 91, 91)


def test_create_grid_and_inference_interfaces():
    from fovea.models import CompressNet, DeformSegmentationModule
    from fovea.saliency_network import fov_simple
    from oracle import reference_port as rp
    cfg = make_cfg(True)
    m = DeformSegmentationModule(TinyEncoder(), TinyDecoder(), fov_simple(cfg), CompressNet(cfg), None, cfg).cuda().eval()
    xs, _ = rp.synthetic_saliency(2, seed=4)
    xs_hm = rp.pad_saliency(xs, 45, 45).cuda()
    grid, grid_y = m.create_grid(xs_hm)
    assert grid.shape == (2, 80, 80, 2) and grid_y.shape == (2, 80, 80, 2)
    grid2, grid_inv = m.create_grid(xs_hm, segSize=(128, 160), x_inv=1 - xs_hm)
    assert grid_inv.shape == (2, 128, 160, 2)
    want = rp.grid_inverse(grid2.cpu(), (128, 160), tie="max")
    assert np.array_equal(grid_inv.cpu().numpy(), want.numpy(), equal_nan=True)
    feed = {k: v.cuda() for k, v in synthetic_batch(2, 128, 160, 9).items()}
    with torch.no_grad():
        cfg.VAL.no_upsample = False
        pred_sampled, pred, y_sampled = m(feed, segSize=(128, 160))
    assert pred_sampled.shape == (2, 51, 128, 160) and pred.shape == (2, 51, 80, 80) and y_sampled.shape == (2, 80, 80)
    assert not torch.isnan(pred_sampled).any()


def test_fill_missing_values_and_interp2d_api(golden_dir):
    """The reference-named entry points on arbitrary tensors / point sets (host triangulation = exact parity)."""
    from fovea.interp2d import Interp2D
    from fovea.models import fillMissingValues_tensor
    from oracle import reference_port as rp
    g = dict(np.load(os.path.join(golden_dir, "interp2d_64x48.npz")))
    h, w = (int(v) for v in g["hw"])
    pts, vals = torch.from_numpy(g["points"]), torch.from_numpy(g["values"])
    for tri_mode in ("host", "device"):
        out = Interp2D(h, w, triangulation=tri_mode)(pts.cuda(), vals.cuda()).cpu().numpy()
        close = np.isclose(out, g["out"], rtol=0, atol=1e-5)
        print("Interp2D", tri_mode, "agreement", close.mean())
        assert close.mean() > (0.999 if tri_mode == "host" else 0.95)   # measured 1.0 / 0.981 (co-circular cells)
    gi = dict(np.load(os.path.join(golden_dir, "inverse_80_to_128.npz")))
    t = torch.from_numpy(gi["pred_sampled_nan"][0]).cuda()
    want = rp.fill_missing_values_tensor(torch.from_numpy(gi["pred_sampled_nan"][0]).clone())
    got = fillMissingValues_tensor(t, copy=True, triangulation="host").cpu()
    same_nan = (torch.isnan(got) == torch.isnan(want)).float().mean().item()
    close = torch.isclose(got, want, rtol=0, atol=1e-5, equal_nan=True).float().mean().item()
    assert same_nan > 0.999 and close > 0.999, (same_nan, close)
    assert torch.equal(torch.isnan(t), torch.from_numpy(np.isnan(gi["pred_sampled_nan"][0])).cuda())  # copy=True


def test_resample_pipeline_matches_direct_path():
    """fovea.pipeline.ResamplePipeline (3 streams x 2 slots, host buffers) gives the same masks as the direct ops path,
    for both image-ingest modes (bulk H2D copy / on-demand PCIe gather from the pinned host image)."""
    from fovea import ops
    from fovea.pipeline import ResamplePipeline
    from oracle import reference_port as rp
    B, C, H, W, g, R = 3, 7, 256, 320, 80, 45
    batches = []
    for seed in range(4):
        xs, _ = rp.synthetic_saliency(B, seed=seed)
        pred = rp.synthetic_pred(B, C, seed=seed)
        x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(seed))
        batches.append((x.pin_memory(), xs.pin_memory(), pred.pin_memory()))
    g1x, g1y = (t.cuda() for t in ops.separable_factors(rp.gaussian_filter_weight(R, R, R)))
    want = []
    for x, xs, pred in batches:
        grid = ops.saliency_to_grid(xs.cuda(), g1x, g1y, g, g, R, R, "replication", (g, g))
        plan = ops.build_inverse_plan(grid, (H, W), nchan=C, triangulation="device")
        _, mask = ops.inverse_fill(plan, pred.cuda(), want_scores=False, want_mask=True)
        want.append((ops.grid_sample(x.cuda(), grid).cpu(), mask.cpu()))
    for on_host in (False, True, 0.34):
        pipe = ResamplePipeline(B, C, H, W, g, R, torch.device("cuda", 0), "device", depth=2, image_on_host=on_host)
        outs = [torch.empty(B, H, W, dtype=torch.int64).pin_memory() for _ in batches]
        sampled = [pipe.submit(*bt, out) for bt, out in zip(batches, outs)]
        pipe.drain()
        for (xs_want, mask_want), got, xs_got in zip(want, outs, sampled):
            assert torch.equal(xs_got.cpu(), xs_want), f"x_sampled differs (image_on_host={on_host})"
            assert torch.equal(got, mask_want), f"{int((got != mask_want).sum())} mask pixels differ (image_on_host={on_host})"
    with pytest.raises(Exception):
        pipe.submit(batches[0][0].clone(), batches[0][1], batches[0][2], outs[0])   # unpinned image is refused


def test_fill_missing_values_nearest_api(golden_dir):
    """fillMissingValues_tensor(t, interp_mode='nearest') on a NaN-holed tensor against the oracle / the reference golden."""
    from fovea.models import fillMissingValues_tensor
    from oracle import reference_port as rp
    s = dict(np.load(os.path.join(golden_dir, "inverse_80_to_128.npz")))
    g = dict(np.load(os.path.join(golden_dir, "nearest_80_to_128.npz")))
    t = torch.from_numpy(s["pred_sampled_nan"][0])
    got = fillMissingValues_tensor(t.clone().cuda(), copy=True, interp_mode="nearest").cpu()
    want = torch.from_numpy(g["pred_sampled_nearest"][0])
    assert not torch.isnan(got).any()
    valid = ~torch.isnan(t)
    assert torch.equal(got[valid], t[valid])
    # on a 128x128 integer lattice ~12 % of the pixels have several equidistant sites, among which the reference's
    # KD-tree and the kernel's (|dx|, left, upper) rule choose differently; tests/test_parity_gpu.py::_check_nearest
    # proves pixel by pixel that every disagreement is such a tie
    close = ((got - want).abs() <= 1e-5 * want.abs().max()).all(0)
    assert close.float().mean().item() > 0.8


def test_module_nearest_mode_and_async_plan():
    """cfg.MODEL.rev_deform_interp='nearest' (config/deform.yaml:17) through the module's inference branch, and the
    side-stream plan (overlapping the encoder) against the plan built after the decoder."""
    from fovea import ops
    from fovea.models import CompressNet, DeformSegmentationModule
    from fovea.saliency_network import fov_simple
    torch.manual_seed(3)
    feed = {k: v.cuda() for k, v in synthetic_batch(2, 128, 160, 9).items()}
    outs = {}
    for mode in ("nearest", "tri"):
        cfg = make_cfg(True)
        cfg.MODEL.rev_deform_interp = mode
        cfg.VAL.no_upsample = False
        torch.manual_seed(3)
        m = DeformSegmentationModule(TinyEncoder(), TinyDecoder(), fov_simple(cfg), CompressNet(cfg), None, cfg).cuda().eval()
        with torch.no_grad():
            ps_async, pred, _ = m(feed, segSize=(128, 160))
            cfg.DATASET.num_class = None                      # no class count in the config -> plan after the decoder
            ps_sync, pred2, _ = m(feed, segSize=(128, 160))
        assert torch.equal(pred, pred2) and torch.equal(ps_async, ps_sync)
        assert ps_async.shape == (2, 51, 128, 160) and not torch.isnan(ps_async).any()
        outs[mode] = ps_async
    # both modes leave the pixels that received a node untouched, and differ elsewhere (step function vs. linear)
    assert not torch.equal(outs["nearest"], outs["tri"])


def test_module_inference_accepts_uint8_frames():
    """SURVEY 8f row 3 through the module: a uint8 img_data (the loader's decode, before ToTensor) gives bit for bit the
    outputs of the converted fp32 image at inference; under autograd it is refused (the grid gradient needs fp32 taps)."""
    from fovea.models import CompressNet, DeformSegmentationModule
    from fovea.saliency_network import fov_simple
    cfg = make_cfg(True)
    cfg.VAL.no_upsample = False
    torch.manual_seed(4)
    m = DeformSegmentationModule(TinyEncoder(), TinyDecoder(), fov_simple(cfg), CompressNet(cfg), None, cfg).cuda().eval()
    feed = {k: v.cuda() for k, v in synthetic_batch(2, 128, 160, 11).items()}
    img8 = (feed["img_data"] * 255).round().to(torch.uint8)
    # ToTensor divides on the CPU (a true fp32 division); torch's CUDA `/ 255.0` multiplies by the reciprocal instead
    feed32 = dict(feed, img_data=(img8.cpu().float() / 255.0).cuda())
    feed8 = dict(feed, img_data=img8)
    with torch.no_grad():
        a = m(dict(feed32), segSize=(128, 160))
        b = m(dict(feed8), segSize=(128, 160))
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    m.train()
    with pytest.raises(NotImplementedError):
        m(dict(feed8))


def test_inference_return_tuples_match_the_reference_signature():
    """models_instance.py:1112-1121: (pred_sampled, pred, y_sampled) -- also with VAL.no_upsample (that flag only renames
    tensors for the reference's visualisation) --, a 4-tuple with VAL.y_sampled_reverse, (pred_sampled, loss) for
    F_Xlr_acc_map; rev_deform_interp='BI' runs on the same kernels with the NB site rule."""
    from fovea.models import CompressNet, DeformSegmentationModule
    from fovea.saliency_network import fov_simple
    feed = {k: v.cuda() for k, v in synthetic_batch(2, 128, 160, 9).items()}
    cfg = make_cfg(True)
    torch.manual_seed(5)
    m = DeformSegmentationModule(TinyEncoder(), TinyDecoder(), fov_simple(cfg), CompressNet(cfg), None, cfg,
                                 triangulation="device").cuda().eval()
    with torch.no_grad():
        cfg.VAL.no_upsample = True
        a = m(dict(feed), segSize=(128, 160))
        cfg.VAL.no_upsample = False
        b = m(dict(feed), segSize=(128, 160))
        assert len(a) == len(b) == 3 and all(torch.equal(u, v) for u, v in zip(a, b))
        assert a[0].shape == (2, 51, 128, 160) and a[1].shape == (2, 51, 80, 80) and a[2].shape == (2, 80, 80)
        cfg.VAL.y_sampled_reverse = True
        c = m(dict(feed), segSize=(128, 160))
        assert len(c) == 4 and c[3].shape == (2, 128, 160) and c[3].dtype == torch.int64
        assert set(c[3].unique().tolist()) <= {0, 1}                  # the label is {0,1}: its inverse upsampling too
        # the label pushed down and up again agrees with the label on most of the frame (the intrinsic upsampling error)
        assert (c[3] == feed["seg_label"].squeeze(1).long()).float().mean().item() > 0.9
        cfg.VAL.y_sampled_reverse = False
        ps, loss = m(dict(feed), segSize=(128, 160), F_Xlr_acc_map=True)
        assert torch.equal(ps, a[0]) and loss.dim() == 0 and torch.isfinite(loss)
        cfg.MODEL.rev_deform_interp = "BI"
        d = m(dict(feed), segSize=(128, 160))
        assert d[0].shape == a[0].shape and torch.isfinite(d[0]).all() and not torch.equal(d[0], a[0])


def test_device_pipeline_matches_direct_path():
    """fovea.pipeline.DevicePipeline (plan of batch i+1 on a high-priority stream over the fill of batch i): scores mode and
    mask mode (want_scores=False: the pruned arg-max fill) give the direct path's results; x_sampled is ordered on the
    caller's stream when submit returns."""
    from fovea import ops
    from fovea.pipeline import DevicePipeline
    from oracle import reference_port as rp
    B, C, H, W, g, R = 3, 9, 256, 320, 80, 45
    g1x, g1y = (t.cuda() for t in ops.separable_factors(rp.gaussian_filter_weight(R, R, R)))
    batches, want = [], []
    for seed in range(4):
        xs, _ = rp.synthetic_saliency(B, seed=seed)
        x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(seed)).cuda()
        pred = rp.synthetic_pred(B, C, seed=seed).cuda()
        batches.append((x, xs.cuda(), pred))
        grid = ops.saliency_to_grid(xs.cuda(), g1x, g1y, g, g, R, R, "replication", (g, g))
        plan = ops.build_inverse_plan(grid, (H, W), nchan=C, triangulation="device")
        s, m = ops.inverse_fill(plan, pred, want_scores=True, want_mask=True)
        want.append((ops.grid_sample(x, grid), s, m))
    pipe = DevicePipeline(B, C, H, W, g, R, triangulation="device", depth=2)
    for (x, xs, pred), (xs_want, s_want, _) in zip(batches, want):
        x_sampled, scores = pipe.submit(x, xs, pred)
        assert torch.equal(x_sampled, xs_want)            # consumed on the caller's stream, no explicit fence
        pipe.fence()
        assert torch.equal(scores, s_want)
    pipe.check()
    mpipe = DevicePipeline(B, C, H, W, g, R, triangulation="device", depth=2, want_mask=True, want_scores=False)
    for (x, xs, pred), (_, _, m_want) in zip(batches, want):
        _, mask = mpipe.submit(x, xs, pred)
        mpipe.fence()
        assert mask.dtype == torch.int64 and torch.equal(mask, m_want)
    with pytest.raises(Exception):
        DevicePipeline(B, C, H, W, g, R, want_mask=False, want_scores=False)

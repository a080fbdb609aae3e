"""Mirror of the reference's deform-segmentation interface (models/models.py) on the sm_100a kernels.

Exports the names the reference's callers import (`train_deform_semantic.py:22`, `eval.py:16`):
`DeformSegmentationModule`, `CompressNet`, `fillMissingValues_tensor`, `makeGaussian`, `FocalLoss` (+ `b_imresize`).
Encoder / decoder / saliency networks are whatever `nn.Module`s the caller passes in (HRNet / Segformer / DeepLab and
`C1` stay stock PyTorch, north_star); only the foveated resampling path runs on libfovea_b200.so:

    create_grid            models/models.py:594-657   -> fovea_grid_fwd/bwd (+ fovea_grid_resize, grid_inv kernels)
    F.grid_sample(x|y, .)  models/models.py:880, 909  -> fovea_grid_sample_fwd/bwd
    grid_inv + grid_sample(pred, grid_inv) + NaN mask + fillMissingValues_tensor('tri')
                           models/models.py:933-940, models_instance.py:883-893, 940 -> fovea_inverse_fill

Dead work of the reference forward whose result is unobservable is skipped and listed in DESIGN.md (the per-sample
PIL blur/edge loop models.py:776-800, the second/third create_grid calls :849-852, PNG dumps :973-1051, d(filter.weight)).
Config branches the shipped `config/deform.yaml` never takes (uniform_sample, gt_gradient, deep supervision,
dynamic_task_input) raise NotImplementedError instead of silently running something else.

Triangulation of rev_deform_interp='tri': the reference runs Qhull on the host (interp2d.py:55).  The default here is the
same ("host": bit-for-bit the reference's mesh); `triangulation="device"` (argument, or FOVEA_TRIANGULATION=device in
the environment) selects the sm_100a Delaunay kernel, which differs from Qhull only inside co-circular lattice cells.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import os

from . import ops
from ._lib import FoveaError
from .interp2d import Interp2D, interp2d_scores


def default_triangulation():
    """"host" (Qhull, the reference's own mesh) unless FOVEA_TRIANGULATION=device opts into the Delaunay kernel."""
    mode = os.environ.get("FOVEA_TRIANGULATION", "host")
    if mode not in ("host", "device"):
        raise FoveaError(f"FOVEA_TRIANGULATION must be 'host' or 'device', got {mode!r}")
    return mode


def b_imresize(im, size, interp="bilinear"):
    """dataset.py:30-31."""
    return F.interpolate(im, size, mode=interp)


def makeGaussian(size, fwhm=3, center=None):
    """models/models.py:140-157."""
    x = np.arange(0, size, 1, float)
    y = x[:, np.newaxis]
    x0 = y0 = size // 2 if center is None else None
    if center is not None:
        x0, y0 = center[0], center[1]
    return np.exp(-4 * np.log(2) * ((x - x0) ** 2 + (y - y0) ** 2) / fwhm ** 2)


class FocalLoss(nn.Module):
    """models/models.py:87-120 (stock PyTorch, not on the hot path)."""

    def __init__(self, gamma=0, size_average=True):
        super().__init__()
        self.gamma, self.size_average = gamma, size_average

    def forward(self, input, target):
        if input.dim() > 2:
            input = input.view(input.size(0), input.size(1), -1).transpose(1, 2).contiguous().view(-1, input.size(1))
        logpt = F.log_softmax(input, dim=1).gather(1, target.view(-1, 1)).view(-1)
        pt = logpt.detach().exp()
        loss = -1 * (1 - pt) ** self.gamma * logpt
        return loss.mean() if self.size_average else loss.sum()


class MulticlassDiceLoss(nn.Module):
    """pytorch_toolbelt.losses.dice.DiceLoss('multiclass') as the reference constructs it (models/models.py:482):
    from_logits=True, smooth=0, eps=1e-7, classes absent from the target are masked out of the mean."""

    def forward(self, y_pred, y_true):
        bs, nc = y_pred.shape[:2]
        p = y_pred.log_softmax(dim=1).exp().view(bs, nc, -1)
        t = F.one_hot(y_true.view(bs, -1).long(), nc).permute(0, 2, 1).type_as(p)
        inter = torch.sum(p * t, dim=(0, 2))
        card = torch.sum(p + t, dim=(0, 2))
        dice = (2.0 * inter) / card.clamp_min(1e-7)
        loss = (1.0 - dice) * (t.sum(dim=(0, 2)) > 0).to(p.dtype)
        return loss.mean()


class CompressNet(nn.Module):
    """models/models.py:360-372."""

    def __init__(self, cfg):
        super().__init__()
        cin = 24 if cfg.MODEL.saliency_net == "fovsimple" else 256
        self.conv_last = nn.Conv2d(cin, 1, kernel_size=1, padding=0, stride=1)
        self.act = nn.ReLU(inplace=False)

    def forward(self, x):
        return self.conv_last(self.act(x))


def _fill_nearest(t):
    """interp_mode='nearest' on a CUDA tensor [C,H,W] (models/models.py:213-250, 259-272): every NaN pixel takes the value
    of the nearest site = valid pixel with a NaN pixel directly above/below (getPixelsForInterp_NB; cv2.dilate reads the
    [C,H,W] array as rows=C, cols=H, channels=W, so its cross never looks left/right); no forced corners."""
    C, H, W = t.shape
    invalid = torch.isnan(t[0])
    valid = ~invalid
    n = int(valid.sum())
    if n == 0:
        return t                                  # no site: the reference's KD-tree would fail on an empty point set
    if n + 2 > 32768:
        raise FoveaError(f"fillMissingValues_tensor('nearest'): {n} valid pixels exceed the value-table limit (32766)")
    winner = torch.full((1, H, W), -1, device=t.device, dtype=torch.int32)
    winner[0][valid] = torch.arange(n, device=t.device, dtype=torch.int32)
    Cs = (C + 7) // 8 * 8
    table = torch.zeros(1, n + 2, Cs, device=t.device, dtype=torch.float32)
    table[0, :n, :C] = t[:, valid].T
    table[0, n] = float("nan")
    loc = ops.nearest_locate(winner, n, 1, C)
    plan = ops.InversePlan(winner, None, None, None, None, None, None,
                           torch.zeros(1, 1, 16, device=t.device, dtype=torch.int32), loc, n, 1, H, W, n + 4, 1, "nearest")
    out = torch.empty(1, C, H, W, device=t.device, dtype=torch.float32)
    ops.inverse_fill_table(plan, table, C, zero_residual=False, scores=out)
    t[:, invalid] = out[0][:, invalid]            # :272
    return t


def fillMissingValues_tensor(target_for_interp, copy=False, interp_mode="tri", triangulation=None):
    """models/models.py:159-286 for interp_mode='tri' | 'nearest' on a CUDA tensor [C,H,W]; in place unless `copy`.

    The NaN pattern is taken from channel 0 (the reference's own point extraction uses `mask_for_interp[0]`,
    :265, and requires the pattern to be identical across channels for its `.view(C,-1)`, :268)."""
    if interp_mode not in ("tri", "nearest"):
        raise NotImplementedError("rev_deform_interp='BI' (host SciPy LinearNDInterpolator) is not on the GPU path; "
                                  "use 'tri' (the same piecewise-linear interpolant) or 'nearest'")
    t = target_for_interp.clone() if copy else target_for_interp
    if not t.is_cuda:
        raise FoveaError("fillMissingValues_tensor: expected a CUDA tensor (there is no CPU fallback)")
    if interp_mode == "nearest":
        return _fill_nearest(t)
    C, H, W = t.shape
    invalid = torch.isnan(t[0])
    if not bool(invalid.any()):                                                  # :254-255
        return t
    kernel = torch.tensor([[0., 1., 0.], [1., 1., 1.], [0., 1., 0.]], device=t.device).view(1, 1, 3, 3)
    inv_f = invalid.float()[None, None]
    if max(C, H, W) > 512:                                                      # :183-193
        dr = max(C, H, W) / 512
        scaled = F.interpolate(inv_f, (int(H / dr), int(W / dr)), mode="nearest")
        dil = torch.clamp(F.conv2d(scaled, kernel, padding=1), 0, 1)
        dil = F.interpolate(dil, (H, W), mode="nearest")
    else:
        dil = torch.clamp(F.conv2d(inv_f, kernel, padding=1), 0, 1)
    mask = (dil[0, 0] > 0) & ~invalid                                           # :200
    mask[0, 0] = mask[0, -1] = mask[-1, 0] = mask[-1, -1] = True                # :202-209
    rr, cc = torch.where(mask)                                                  # :265-267
    values = t[:, rr, cc].T.contiguous()                                        # :268 (NaN at unfilled corners)
    interp = interp2d_scores(torch.stack([rr, cc], 1), values, H, W, triangulation or default_triangulation())
    t[:, invalid] = interp[:, invalid]                                          # :280
    return t


class SegmentationModuleBase(nn.Module):
    """Pixel-accuracy metrics of models/models.py:378-474 (vectorised over the batch; same values)."""

    @staticmethod
    def _preds(pred):
        return torch.max(pred, dim=1)[1]

    def pixel_acc(self, pred_all, label_all):
        preds = self._preds(pred_all)
        valid, valid1 = label_all < 50, preds < 50
        acc_sum = (valid & (preds == label_all)).flatten(1).sum(1).float()
        union = (valid | valid1).flatten(1).sum(1).float()
        return (acc_sum / (union + 1e-10)).mean()

    def fg_bin_pixel_acc(self, pred_all, label_all):
        preds = self._preds(pred_all)
        valid, valid1 = label_all < 50, preds < 50
        acc_sum = (valid & valid1).flatten(1).sum(1).float()
        union = (valid | valid1).flatten(1).sum(1).float()
        return (acc_sum / (union + 1e-10)).mean()

    def fbg_cls_pixel_acc(self, pred_all, label_all):
        preds = self._preds(pred_all)
        same = preds == label_all
        out = 0.0
        for lab_is, pred_is in ((label_all < 50, preds < 50), (label_all == 50, preds == 50)):
            acc = (lab_is & same).flatten(1).sum(1).float() / ((lab_is | pred_is).flatten(1).sum(1).float() + 1e-10)
            out = out + 0.5 * acc
        return out.mean()

    def fbg_bin_pixel_acc(self, pred_all, label_all):
        preds = self._preds(pred_all)
        out = 0.0
        for lab_is, pred_is in ((label_all < 50, preds < 50), (label_all == 50, preds == 50)):
            acc = (lab_is & pred_is).flatten(1).sum(1).float() / ((lab_is | pred_is).flatten(1).sum(1).float() + 1e-10)
            out = out + 0.5 * acc
        return out.mean()


class DeformSegmentationModule(SegmentationModuleBase):
    """models/models.py:476-1094 (training / eval-at-low-res forward) + models_instance.py:840-1121 (inference)."""

    def __init__(self, net_encoder, net_decoder, net_saliency, net_compress, crit, cfg, deep_sup_scale=None,
                 triangulation=None):
        super().__init__()
        self.encoder, self.decoder = net_encoder, net_decoder
        self.localization, self.net_compress = net_saliency, net_compress
        self.crit = MulticlassDiceLoss()                                        # models.py:482 ignores `crit` too
        self.crit_mse = nn.MSELoss()
        self.cfg, self.deep_sup_scale = cfg, deep_sup_scale
        self.triangulation = triangulation or default_triangulation()
        if deep_sup_scale is not None:
            raise NotImplementedError("deep supervision is not part of the foveated path")
        sal = cfg.TRAIN.saliency_input_size
        short = cfg.MODEL.saliency_output_size_short
        self.grid_size_x = sal[0] if short == 0 else short                      # :490-494
        self.grid_size_y = sal[1] // (sal[0] // self.grid_size_x)
        self.padding_size_x = cfg.MODEL.gaussian_radius                         # :495-500
        ap = sal[1] // sal[0] if cfg.MODEL.gaussian_ap == 0.0 else cfg.MODEL.gaussian_ap
        self.padding_size_y = int(ap * self.padding_size_x)
        self.global_size_x = self.grid_size_x + 2 * self.padding_size_x
        self.global_size_y = self.grid_size_y + 2 * self.padding_size_y
        self.input_size = tuple(sal)
        self.input_size_net = tuple(cfg.TRAIN.task_input_size)
        self.input_size_net_eval = tuple(cfg.TRAIN.task_input_size_eval)
        self.input_size_net_infer = self.input_size_net_eval if len(self.input_size_net_eval) else self.input_size_net
        Kx, Ky = 2 * self.padding_size_x + 1, 2 * self.padding_size_y + 1
        g = torch.FloatTensor(makeGaussian(Kx, fwhm=cfg.MODEL.gaussian_radius))  # :510-515
        g = b_imresize(g[None, None], (Kx, Ky), interp="bilinear")[0, 0]
        self.filter = nn.Conv2d(1, 1, kernel_size=(Kx, Ky), bias=False)         # kept: state_dict / DDP shapes match
        self.filter.weight.data[0, 0] = g
        i = np.arange(self.global_size_x, dtype=np.float64)[:, None]            # :517-522
        j = np.arange(self.global_size_y, dtype=np.float64)[None, :]
        P = np.stack([(j - self.padding_size_y) / (self.grid_size_y - 1.0) + 0 * i,
                      (i - self.padding_size_x) / (self.grid_size_x - 1.0) + 0 * j])
        self.P_basis = torch.from_numpy(P.astype(np.float32))                   # plain attribute, as the reference
        self._factors = None
        self._plan_stream = None

    # -- stage 1 -----------------------------------------------------------------------------------------------
    def _g1(self, device):
        """1-D factors of the (rank-1) Gaussian filter weight, re-derived if the parameter object changed."""
        w = self.filter.weight
        key = (w.data_ptr(), w._version, str(device))
        if self._factors is None or self._factors[0] != key:
            g1x, g1y = ops.separable_factors(w)
            self._factors = (key, g1x.to(device), g1y.to(device))
        return self._factors[1], self._factors[2]

    def _grid_sizes(self, segSize):
        infer = len(self.input_size_net_eval) != 0 and segSize is not None      # :621-625
        size = self.input_size_net_infer if infer else self.input_size_net
        if segSize is None:                                                     # :627-631
            size_y = tuple(int(v) // self.cfg.DATASET.segm_downsampling_rate for v in self.input_size_net)
        else:
            size_y = tuple(self.input_size_net_infer)
        return size, size_y

    def create_grid(self, x, segSize=None, x_inv=None):
        """models/models.py:594-657 on the PADDED saliency map x = xs_hm [B,1,G,G] (same signature and returns)."""
        g1x, g1y = self._g1(x.device)
        size, size_y = self._grid_sizes(segSize)
        grid = ops.saliency_to_grid(x, g1x, g1y, self.grid_size_x, self.grid_size_y, self.padding_size_x,
                                    self.padding_size_y, "none", size)
        if segSize is not None and x_inv is not None:                          # :640-655
            winner = ops.grid_inv_scatter(grid, segSize)
            return grid, ops.grid_inv_canvas(winner, grid.shape[1], grid.shape[2])
        return grid, ops.grid_resize(grid, size_y)

    def _grid_from_saliency(self, xs, segSize=None):
        """create_grid with the padding (:819-825) fused into the filter taps: xs is the UNPADDED saliency."""
        mode = self.cfg.TRAIN.def_saliency_pad_mode
        g1x, g1y = self._g1(xs.device)
        size, size_y = self._grid_sizes(segSize)
        grid = ops.saliency_to_grid(xs, g1x, g1y, self.grid_size_x, self.grid_size_y, self.padding_size_x,
                                    self.padding_size_y, mode, size)
        return grid, ops.grid_resize(grid, size_y)

    # -- stage 0: saliency branch (the network itself is stock PyTorch) -------------------------------------------
    def _saliency(self, x, focus_point):
        # :684-705: focus map + b_imresize + two cats in one launch (x may be fp32 or the loader's uint8 frame)
        x_low = ops.saliency_input(x, focus_point.to(x.device), self.input_size)
        xs = self.net_compress(self.localization(x_low))                        # :711-713
        if tuple(xs.shape[-2:]) != (self.grid_size_x, self.grid_size_y):        # nn.Upsample, identity by default
            xs = F.interpolate(xs, (self.grid_size_x, self.grid_size_y), mode="bilinear")
        xs = ops.saliency_softmax(xs.reshape(-1, self.grid_size_x * self.grid_size_y))  # :715-723
        return xs.view(-1, 1, self.grid_size_x, self.grid_size_y)

    @staticmethod
    def _sample_image(x, grid):
        """F.grid_sample(x, grid) (:909); a uint8 frame (the loader's decode, before ToTensor) is sampled directly with the
        /255 folded into the taps -- bit-identical to converting first, a quarter of the bytes."""
        if x.dtype == torch.uint8:
            if torch.is_grad_enabled() and grid.requires_grad:
                raise NotImplementedError("a uint8 img_data cannot carry the grid gradient: pass the fp32 image when training")
            return ops.grid_sample_u8(x, grid.detach())
        return ops.grid_sample(x, grid)

    def _check_cfg(self):
        c = self.cfg
        if c.MODEL.uniform_sample != "" or getattr(c.MODEL, "gt_gradient", False) or c.TRAIN.dynamic_task_input[0] != 1:
            raise NotImplementedError("only the foveated (deform.yaml) configuration runs on the B200 path")
        if not (c.TRAIN.deform_joint_loss and c.TRAIN.opt_deform_LabelEdge_norm):
            raise NotImplementedError("only the joint-loss / normalised edge-loss configuration of deform.yaml")

    def _build_plan(self, grid, segSize, nchan):
        mode = self.cfg.MODEL.rev_deform_interp
        if mode == "nearest":                      # config/deform.yaml:17
            return ops.build_nearest_plan(grid.detach(), segSize, nchan=nchan)
        if mode == "tri":
            return ops.build_inverse_plan(grid.detach(), segSize, nchan=nchan, triangulation=self.triangulation)
        if mode == "BI":                           # models/models.py:248-250 (see ops.build_inverse_plan, sites="nb")
            return ops.build_inverse_plan(grid.detach(), segSize, nchan=nchan, triangulation=self.triangulation,
                                          sites="nb")
        raise NotImplementedError(f"rev_deform_interp={mode!r}: expected 'tri', 'nearest' or 'BI'")

    def plan_async(self, grid, segSize):
        """Start the saliency-only half of stage 3 (A7 scatter, A9 point selection, Delaunay, point location) on a
        high-priority side stream as soon as the grid exists, so that it overlaps the encoder/decoder; returns a handle
        for `inverse_upsample(..., plan=handle)`.  Needs the class count before the decoder has run: taken from
        cfg.DATASET.num_class (it only enters the >512-pixel rule of the point selection, models/models.py:183); when
        the config has none, returns None and the plan is built after the decoder instead."""
        nchan = getattr(getattr(self.cfg, "DATASET", None), "num_class", None)
        if nchan is None:
            return None
        if self._plan_stream is None:
            self._plan_stream = torch.cuda.Stream(grid.device, priority=-1)
        side = self._plan_stream
        side.wait_stream(torch.cuda.current_stream(grid.device))
        with torch.cuda.stream(side):
            plan = self._build_plan(grid, segSize, int(nchan))
            done = torch.cuda.Event()
            done.record(side)
        grid.record_stream(side)
        return plan, done, int(nchan)

    def inverse_upsample(self, pred, grid, segSize, zero_residual, want_mask=False, plan=None):
        """grid_inv + F.grid_sample(pred, grid_inv) + NaN mask + per-sample fill (models.py:933-940) fused.
        Differentiable w.r.t. `pred` (the reference's autograd graph through F.grid_sample + Interp2D)."""
        if tuple(pred.shape[-2:]) != tuple(grid.shape[1:3]):
            raise NotImplementedError(f"inverse upsampling needs the decoder output {tuple(pred.shape[-2:])} at the "
                                      f"sampling-grid resolution {tuple(grid.shape[1:3])}")
        if plan is not None and plan[2] == pred.shape[1]:
            plan, done, _ = plan
            torch.cuda.current_stream(pred.device).wait_event(done)
            for t in (plan.loc, plan.trirec, plan.winner):           # allocated on the side stream, consumed here
                t.record_stream(torch.cuda.current_stream(pred.device))
        else:
            plan = self._build_plan(grid, segSize, pred.shape[1])
        out = ops.inverse_fill(plan, pred, want_scores=True, want_mask=want_mask, zero_residual=zero_residual)
        ops.check_plan(plan)     # (after the fill is enqueued: the one host sync reads a [B] status word)
        return out

    # -- forward ------------------------------------------------------------------------------------------------
    def forward(self, feed_dict, *, writer=None, segSize=None, F_Xlr_acc_map=False, count=None, epoch=None,
                feed_dict_info=None, feed_batch_count=None, cur_iter=None, is_inference=False, rank=None):
        self._check_cfg()
        cfg = self.cfg
        x = feed_dict["img_data"]
        H_HS, W_HS = x.shape[-2:]
        xs = self._saliency(x, feed_dict["focus_point"])
        y = feed_dict["seg_label"].clone()
        if segSize is not None:
            return self._forward_inference(feed_dict, x, xs, y, segSize, F_Xlr_acc_map)

        # ---- training / eval at low resolution (models/models.py:828-1094)
        xs_target = F.interpolate(y if y.dim() == 4 else y.unsqueeze(0), size=(self.grid_size_x, self.grid_size_y),
                                  mode="area")                                  # :730
        grid, grid_y = self._grid_from_saliency(xs)                             # :821 + :845
        y_sampled = ops.grid_sample(y.float(), grid_y).squeeze(1)               # :880
        xs_n = (xs - xs.min()) / (xs.max() - xs.min())                          # :889-898
        xt_n = (xs_target - xs_target.min()) / (xs_target.max() - xs_target.min())
        edge_loss = 0.05 * self.crit_mse(xs_n, xt_n) * cfg.TRAIN.edge_loss_scale
        high_res = bool(getattr(cfg.MODEL, "loss_at_high_res", False))
        upsample = cfg.MODEL.upsample or high_res
        rate = int(cfg.DATASET.segm_downsampling_rate)
        inv_size = (H_HS // rate, W_HS // rate)                                 # :873 ori_size // segm_downsampling_rate
        plan = self.plan_async(grid, inv_size) if upsample else None            # overlaps the encoder/decoder
        x_sampled = self._sample_image(x, grid)                                    # :909
        pred = self.decoder(self.encoder(x_sampled, return_feature_maps=True))  # :926
        if high_res:                                                            # :945-947, :962-965, :1070-1071
            # the loss is taken on the inverse-upsampled scores: the gradient reaches `pred` through the inverse path
            pred_sampled, _ = self.inverse_upsample(pred, grid, inv_size, zero_residual=True, plan=plan)
            loss = self.crit(pred_sampled, feed_dict["seg_label"]) + edge_loss
            acc = self.pixel_acc(pred_sampled, feed_dict["seg_label"].reshape(pred_sampled.shape[0], *pred_sampled.shape[-2:]))
            if is_inference:
                raise NotImplementedError("is_inference with MODEL.loss_at_high_res: the reference's own forward "
                                          "references unset metrics on this branch (models/models.py:1070-1090)")
            return loss, acc, edge_loss
        seg_low = y_sampled.long()
        y_hs = feed_dict["seg_label"].squeeze(1)
        feed_dict["seg_label"] = seg_low                                        # :951 (the reference mutates it too)
        cls = feed_dict["cls_label"].to(pred.device)
        ground_truth = seg_low * cls[:, :, None] + (1 - seg_low) * 50           # :968
        loss = self.crit(pred, ground_truth) + FocalLoss(gamma=5.0)(pred, ground_truth) + edge_loss  # :1057-1069
        if not upsample:
            target, scored = ground_truth, pred
        else:                                                                   # :933-940, :971
            with torch.no_grad():                                               # only the metrics read it on this branch
                scored, _ = self.inverse_upsample(pred, grid, inv_size, zero_residual=False, plan=plan)
            target = (y_hs * cls[:, :, None] + (1 - y_hs) * 50).long()
        acc = self.pixel_acc(scored, target)
        if not is_inference:
            return loss, acc, edge_loss
        return (loss, acc, edge_loss, self.fg_bin_pixel_acc(scored, target), self.fbg_cls_pixel_acc(scored, target),
                self.fbg_bin_pixel_acc(scored, target))

    def _forward_inference(self, feed_dict, x, xs, y, segSize, F_Xlr_acc_map=False):
        """models_instance.py:840-1121 with rev_deform_opt == 51 ('ours deformed case').  Returns what the reference
        returns: (pred_sampled, pred, y_sampled[, y_sampled_reverse]) -- also when VAL.no_upsample is set (that flag
        only renames tensors for the reference's visualisation, :942-949) -- or (pred_sampled, loss) for F_Xlr_acc_map."""
        cfg = self.cfg
        if getattr(cfg.MODEL, "rev_deform_opt", 51) != 51:
            raise NotImplementedError("only MODEL.rev_deform_opt == 51 (the deformed inverse) runs on the B200 path")
        grid, grid_y = self._grid_from_saliency(xs, segSize=segSize)            # :844-845
        plan = self.plan_async(grid, segSize)
        x_sampled = self._sample_image(x, grid)                                    # :851-852
        if tuple(x_sampled.shape[-2:]) != tuple(self.input_size_net_infer):
            x_sampled = F.interpolate(x_sampled, self.input_size_net_infer, mode="bilinear")
        pred = self.decoder(self.encoder(x_sampled, return_feature_maps=True), segSize=tuple(self.input_size_net_infer))
        y4 = y.float() if y.dim() == 4 else y.float().unsqueeze(1)
        y_sampled = F.grid_sample(y4, grid_y, mode="nearest", align_corners=False).long().squeeze(1)   # :866 (stock)
        pred_sampled, _ = self.inverse_upsample(pred, grid, segSize, zero_residual=True, plan=plan)   # :883-893, :940
        if F_Xlr_acc_map:                                                       # :1112-1114
            return pred_sampled, self.crit(pred_sampled, feed_dict["seg_label"])
        if getattr(cfg.VAL, "y_sampled_reverse", False):                        # :904-930
            if cfg.MODEL.rev_deform_interp != "tri":
                raise NotImplementedError("VAL.y_sampled_reverse is built for rev_deform_interp='tri' only")
            nclass = int(cfg.DATASET.num_class)
            onehot = F.one_hot(y_sampled.clamp(0, nclass - 1), nclass).permute(0, 3, 1, 2).float().contiguous()
            onehot = onehot * (y_sampled[:, None] == torch.arange(nclass, device=y_sampled.device)[None, :, None, None])
            rev_plan = self._build_plan(grid, segSize, nclass)
            rev, _ = ops.inverse_fill(rev_plan, onehot, want_scores=True, zero_residual=False)
            ops.check_plan(rev_plan)
            return pred_sampled, pred, y_sampled, torch.max(rev, dim=1)[1].long()   # :930 (NaN pixels: argmax 0, as torch)
        return pred_sampled, pred, y_sampled

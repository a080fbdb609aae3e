"""Reference-side integration: swap the foveated resampling path of an imported FovealSeg checkout for the
B200 kernels WITHOUT touching its encoders, decoder, losses, metrics, launcher or checkpoints.

    import models.models as ref_models          # the reference's own module (train_deform_semantic.py:22)
    import fovea_dropin; fovea_dropin.install(ref_models)

After `install`, `ref_models.DeformSegmentationModule.create_grid`, the module-level `F.grid_sample` used by its
forward, `ref_models.fillMissingValues_tensor` and `ref_models.Interp2D` dispatch to libfovea_b200.so for CUDA fp32
tensors and raise for anything else (there is no CPU fallback).  See INTEGRATION.md.
"""
import types

import torch
import torch.nn.functional as F

from fovea import ops
from fovea.interp2d import Interp2D
from fovea.models import fillMissingValues_tensor
from fovea._lib import FoveaError


def _create_grid(self, x, segSize=None, x_inv=None):
    """Replacement for DeformSegmentationModule.create_grid (models/models.py:594-657), same signature/returns."""
    w = self.filter.weight
    key = (w.data_ptr(), w._version, str(x.device))
    cache = getattr(self, "_fovea_factors", None)
    if cache is None or cache[0] != key:
        g1x, g1y = ops.separable_factors(w)
        cache = (key, g1x.to(x.device), g1y.to(x.device))
        self._fovea_factors = cache
    infer = len(self.input_size_net_eval) != 0 and segSize is not None
    size = tuple(self.input_size_net_infer) if infer else tuple(self.input_size_net)
    grid = ops.saliency_to_grid(x, cache[1], cache[2], self.grid_size_x, self.grid_size_y, self.padding_size_x,
                                self.padding_size_y, "none", size)
    if segSize is not None and x_inv is not None:
        winner = ops.grid_inv_scatter(grid, segSize)
        return grid, ops.grid_inv_canvas(winner, grid.shape[1], grid.shape[2])
    if segSize is None:
        size_y = tuple(int(v) // self.cfg.DATASET.segm_downsampling_rate for v in self.input_size_net)
    else:
        size_y = tuple(self.input_size_net_infer)
    return grid, ops.grid_resize(grid, size_y)


class _Functional(types.ModuleType):
    """`torch.nn.functional` with grid_sample routed to the sm_100a kernel for the reference's call pattern."""

    def __init__(self):
        super().__init__("torch.nn.functional[fovea]")
        self.__dict__.update({k: v for k, v in F.__dict__.items() if not k.startswith("__")})
        self.grid_sample = _grid_sample


def _grid_sample(input, grid, mode="bilinear", padding_mode="zeros", align_corners=None):
    """The reference only calls F.grid_sample with its defaults on 4-D fp32 tensors (models/models.py:865-937)."""
    if mode == "bilinear" and padding_mode == "zeros" and not align_corners and input.dim() == 4 \
            and input.dtype == torch.float32:
        return ops.grid_sample(input, grid)      # raises FoveaError for CPU tensors: no fallback
    return F.grid_sample(input, grid, mode=mode, padding_mode=padding_mode, align_corners=align_corners)


def install(ref_models):
    """Patch an imported reference `models.models` module in place; returns the list of replaced names."""
    if not hasattr(ref_models, "DeformSegmentationModule"):
        raise FoveaError("install(): expected the reference's models.models module")
    ref_models.DeformSegmentationModule.create_grid = _create_grid
    ref_models.F = _Functional()
    ref_models.fillMissingValues_tensor = fillMissingValues_tensor
    ref_models.Interp2D = Interp2D
    return ["DeformSegmentationModule.create_grid", "F.grid_sample", "fillMissingValues_tensor", "Interp2D"]

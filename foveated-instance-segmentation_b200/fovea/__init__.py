"""fovea -- host side of the B200-native foveated resampling path.

Sub-modules: `_lib` (ctypes binding of libfovea_b200.so), `ops` (torch-facing operators).
The reference-facing mirror (DeformSegmentationModule, Interp2D, fillMissingValues_tensor, fov_simple, ...) lives
in `fovea.models`, `fovea.interp2d`, `fovea.saliency_network`.
"""
from ._lib import FoveaError, load  # noqa: F401

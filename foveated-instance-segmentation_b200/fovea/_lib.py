"""ctypes binding of libfovea_b200.so (the C ABI declared in include/fovea_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or a call fails, a
`FoveaError` is raised.  PyTorch is used only for device memory and streams -- no torch type crosses the ABI.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FOVEA_B200_LIB", os.path.join(_HERE, "libfovea_b200.so"))

FOVEA_OK = 0
PAD_NONE, PAD_REPLICATION, PAD_REFLECT, PAD_ZERO = 0, 1, 2, 3
PAD_MODES = {"none": PAD_NONE, "replication": PAD_REPLICATION, "reflect": PAD_REFLECT, "zero": PAD_ZERO}
HINT_CELL_W, HINT_CELL_H = 32, 8
ABI_VERSION = 13


class FoveaError(RuntimeError):
    pass


_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64

# name -> (restype, argtypes); mirrors include/fovea_b200.h one to one
PROTOTYPES = {
    "fovea_abi_version": (_i, []),
    "fovea_last_error": (C.c_char_p, []),
    "fovea_saliency_input": (_i, [_p, _i, C.c_float, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "fovea_saliency_softmax": (_i, [_p, _i, _i, _p, _p]),
    "fovea_saliency_softmax_bwd": (_i, [_p, _p, _i, _i, _p, _p]),
    "fovea_grid_fwd": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p, _p]),
    "fovea_grid_bwd": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p]),
    "fovea_grid_resize": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "fovea_grid_resize_bwd": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "fovea_grid_sample_fwd": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "fovea_grid_sample_fwd_u8": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, C.c_float, _p, _p]),
    "fovea_grid_sample_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "fovea_grid_inv_scatter": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "fovea_grid_inv_canvas": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "fovea_box4_table": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "fovea_select_points": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "fovea_select_points_nb": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "fovea_select_points_sparse": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "fovea_delaunay_workspace_bytes": (_i64, [_i, _i]),
    "fovea_delaunay": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "fovea_delaunay_hints_fused": (_i, [_i, _i, _i]),
    "fovea_delaunay_with_hints": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "fovea_locate_hints_workspace_bytes": (_i64, [_i, _i, _i]),
    "fovea_locate_hints": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "fovea_triangle_setup": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "fovea_locate_pixels": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "fovea_locate_raster_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "fovea_locate_raster": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "fovea_locate_raster_targets": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "fovea_inverse_fill": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p]),
    "fovea_inverse_mask_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "fovea_inverse_mask": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p]),
    "fovea_inverse_fill_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "fovea_nearest_workspace_bytes": (_i64, [_i, _i, _i]),
    "fovea_nearest_locate": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "fovea_scatter_nodes": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "fovea_nearest_locate_all": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "fovea_node_table": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "fovea_probe_store_ceiling": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "fovea_relabel_mask": (_i, [_p, _p, _i, _i64, _i, _p, _p]),
    "fovea_argmax_classes": (_i, [_p, _i, _i, _i64, _p, _p]),
}

_lib = None


def load():
    """Load the shared library once; raise FoveaError (never fall back) if it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FoveaError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` or `make -C "
            f"foveated-instance-segmentation_b200/csrc` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise FoveaError(f"{LIB_PATH} does not export {name}; rebuild the library") from e
        fn.restype = res
        fn.argtypes = args
    if lib.fovea_abi_version() != ABI_VERSION:
        raise FoveaError(f"ABI version mismatch: library {lib.fovea_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != FOVEA_OK:
        msg = lib.fovea_last_error().decode("utf-8", "replace")
        raise FoveaError(f"{name} failed (rc={rc}): {msg}")

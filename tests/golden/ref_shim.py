"""Import shim for the UNMODIFIED reference (/root/reference) -- golden generation only.

TEST INFRASTRUCTURE. Runs only in the build container (the GPU box has no /root/reference).
It stubs the reference's missing third-party imports, makes the hard-coded
`.cuda()` calls no-ops so the module runs on CPU, and injects a `spatial.qhull` stand-in built on
stock SciPy that restates the one functional patch of the vendored fork
(`spatial/qhull.pyx:2075-2163`, `find_simplex(..., return_c=True)`), because the fork itself cannot
be built here (Cython 0.29 / numpy.distutils / py3.7 artefacts).
"""
import sys
import types

import numpy as np
import torch

REF = "/root/reference"


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Delaunay:
    """scipy.spatial.Delaunay + the fork's `return_c` patch (spatial/qhull.pyx:2138-2163)."""

    def __init__(self, points):
        from scipy.spatial import Delaunay
        pts = np.asarray(points, dtype=np.float64)
        self._tri = Delaunay(pts)  # default options "Qbb Qc Qz Q12" (+Qt), as qhull.pyx:1874-1883
        self.simplices = self._tri.simplices
        self.neighbors = self._tri.neighbors
        self.points = self._tri.points

    def find_simplex(self, xi, bruteforce=False, tol=None, return_c=False):
        xi = np.asarray(xi, dtype=np.float64)
        isimplex = self._tri.find_simplex(xi, bruteforce=bruteforce, tol=tol)
        if not return_c:
            return isimplex
        # c as left by _find_simplex_directed/_barycentric_inside (qhull.pyx:1210-1264): computed from
        # tri.transform of the simplex found; for -1 (outside) the fork leaves the last tried simplex's
        # coordinates -- unreachable on this path because the 4 image corners are always points.
        T = self._tri.transform[np.maximum(isimplex, 0)]          # [n,3,2]
        d = xi - T[:, 2, :]
        c01 = np.einsum("nij,nj->ni", T[:, :2, :], d)
        c = np.concatenate([c01, 1.0 - c01.sum(1, keepdims=True)], axis=1)
        return isimplex, c


def install_all():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    _stub("yacs"); _stub("yacs.config", CfgNode=_AttrDict)
    _stub("torchsnooper", snoop=lambda *a, **k: (lambda f: f))
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        _stub("matplotlib"); _stub("matplotlib.pyplot", get_cmap=lambda *a, **k: None)
    _stub("pytorch_toolbelt"); _stub("pytorch_toolbelt.losses")
    _stub("pytorch_toolbelt.losses.dice", DiceLoss=lambda *a, **k: torch.nn.Identity())
    _stub("albumentations")
    _stub("peft", get_peft_model=None, LoraConfig=None, TaskType=None)
    _stub("segmentation_models_pytorch")
    _stub("spatial"); q = _stub("spatial.qhull", Delaunay=_Delaunay); sys.modules["spatial"].qhull = q
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.reset_max_memory_allocated = lambda *a, **k: None
    import models.models as rm
    import interp2d as ri
    from DynamicFocus.utility import torch_tools as tt
    rm.gen_grid_mtx_2xHxW = lambda H, W, device=None: tt.gen_grid_mtx_2xHxW(H, W, device=None)
    rm.Interp2D = ri.Interp2D          # the reference forgets this import (SURVEY surprise 3)
    return rm, ri


def make_cfg(sal=(80, 80), task=(80, 80), R=45, num_class=51, pad="replication", task_eval=(), rate=1):
    C = _AttrDict
    cfg = C(
        TRAIN=C(saliency_input_size=tuple(sal), task_input_size=tuple(task), task_input_size_eval=tuple(task_eval),
                opt_deform_LabelEdge=False, deform_joint_loss=True, num_gpus=1, def_saliency_pad_mode=pad,
                dynamic_task_input=(1,), global_epoch=1),
        MODEL=C(saliency_output_size_short=0, gaussian_radius=R, gaussian_ap=0.0, upsample=False,
                rev_deform_interp="tri", uniform_sample=""),
        DATASET=C(segm_downsampling_rate=rate, num_class=num_class, grid_path="", list_train="ADE"),
        VAL=C(),
    )
    return cfg

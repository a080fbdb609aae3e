"""The C-ABI library loads and exports every symbol include/fovea_b200.h declares (no compute, CPU only)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "fovea_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fovea_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fovea import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in fovea_b200.h but not exported"
    # the ctypes prototypes cover exactly the declared surface
    assert sorted(_lib.PROTOTYPES) == declared


def test_binding_loads_and_reports_version():
    from fovea import _lib
    lib = _lib.load()
    assert lib.fovea_abi_version() == _lib.ABI_VERSION
    assert lib.fovea_last_error() == b""


def test_argument_errors_do_not_need_a_gpu():
    """Argument validation happens before any CUDA call and reports through fovea_last_error()."""
    from fovea import _lib
    lib = _lib.load()
    rc = lib.fovea_grid_fwd(None, 1, 80, 80, 45, 45, 1, None, None, 80, 80, None, None, None)
    assert rc == -1
    assert b"null pointer" in lib.fovea_last_error()
    with pytest.raises(_lib.FoveaError):
        _lib.call("fovea_box4_table", None, 1, 3, 8, 8, 4, None, None)


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Stage-0, C1-tail and deformed_unsampler entry points: argument errors are reported before any CUDA call."""
    from fovea import _lib
    lib = _lib.load()
    assert lib.fovea_saliency_input(None, 0, 255.0, None, 1, 3, 8, 8, 4, 4, None, None) == -1
    assert b"null pointer" in lib.fovea_last_error()
    assert lib.fovea_saliency_softmax(None, 1, 16, None, None) == -1
    assert lib.fovea_relabel_mask(None, None, 1, 16, 3, None, None) == -1
    assert lib.fovea_scatter_nodes(None, 1, 4, 4, 8, 8, None, None) == -1
    assert lib.fovea_node_table(None, 1, 3, 4, 4, 4, None, None) == -1
    assert lib.fovea_nearest_locate_all(None, 1, 4, 4, 8, 8, None, None, None) == -1
    assert lib.fovea_nearest_workspace_bytes(2, 64, 64) >= 2 * 64 * 64 * 2      # covers the scan map and the site lists


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: handing a CPU tensor to the product path fails loudly."""
    import torch
    from fovea import ops, FoveaError
    with pytest.raises(FoveaError):
        ops.grid_sample(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 2, 2))
    with pytest.raises(FoveaError):
        ops.grid_inv_scatter(torch.zeros(1, 4, 4, 2), (8, 8))
    with pytest.raises(FoveaError):
        ops.saliency_input(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2), (4, 4))
    with pytest.raises(FoveaError):
        ops.saliency_softmax(torch.zeros(2, 16))
    with pytest.raises(FoveaError):
        ops.scatter_nodes(torch.zeros(1, 2, 4, 4, dtype=torch.int64), (8, 8))
    from fovea import dynamic_focus
    with pytest.raises(FoveaError):
        dynamic_focus.deformed_unsampler(torch.zeros(1, 1, 2, 2), torch.zeros(1, 2, 2, 2, dtype=torch.int64), 8, 8)


def test_separable_factors_reject_non_gaussian():
    import torch
    from fovea import ops, FoveaError
    from oracle import reference_port as rp
    w = rp.gaussian_filter_weight(12, 24, 12)
    g1x, g1y = ops.separable_factors(w)
    assert torch.allclose(torch.outer(g1x, g1y), w, rtol=0, atol=1e-6)
    with pytest.raises(FoveaError):
        ops.separable_factors(w + 0.1 * torch.rand_like(w))

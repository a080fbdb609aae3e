"""Reference-side integration (dropin.install) -- structural test, CPU only; needs /root/reference (build container)."""
import os
import sys

import pytest

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout only exists in the build container")
def test_install_patches_the_reference_module(golden_dir):
    sys.path.insert(0, golden_dir)
    import ref_shim
    rm, _ = ref_shim.install_all()
    import dropin
    from fovea.interp2d import Interp2D
    from fovea.models import fillMissingValues_tensor
    names = dropin.install(rm)
    assert len(names) == 4
    assert rm.DeformSegmentationModule.create_grid is dropin._create_grid
    assert rm.fillMissingValues_tensor is fillMissingValues_tensor and rm.Interp2D is Interp2D
    assert rm.F.grid_sample is not __import__("torch").nn.functional.grid_sample
    assert rm.F.interpolate is __import__("torch").nn.functional.interpolate      # everything else untouched
    # CPU tensors are refused loudly (no fallback) by the patched path
    import torch
    from fovea import FoveaError
    cfg = ref_shim.make_cfg()
    m = rm.DeformSegmentationModule(None, None, None, None, None, cfg)
    with pytest.raises(FoveaError):
        m.create_grid(torch.rand(1, 1, 170, 170))


def test_mirror_exports_reference_names():
    import fovea.models as fm
    import fovea.saliency_network as fs
    import fovea.interp2d as fi
    for name in ("DeformSegmentationModule", "CompressNet", "fillMissingValues_tensor", "makeGaussian", "FocalLoss",
                 "b_imresize"):
        assert hasattr(fm, name)
    assert hasattr(fs, "fov_simple") and hasattr(fs, "FovSimModule") and hasattr(fi, "Interp2D")
    import inspect
    sig = inspect.signature(fm.DeformSegmentationModule.forward)
    for p in ("feed_dict", "writer", "segSize", "F_Xlr_acc_map", "count", "epoch", "feed_dict_info", "feed_batch_count",
              "cur_iter", "is_inference", "rank"):
        assert p in sig.parameters

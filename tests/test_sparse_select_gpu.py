"""fovea_select_points_sparse + fovea_locate_raster_targets (no dense A7 winner map: the node targets are sorted in shared
memory, the last node of a run of equal pixels wins it, "is this pixel filled" is a binary search) must build the plan of
fovea_grid_inv_scatter + fovea_select_points + fovea_locate_raster bit for bit -- sites, their order, their table rows,
and the per-pixel source map -- on canvases with many node collisions (small), with the dilation run on the
nearest-downscaled mask (> 512 px, models/models.py:183-193) and without (<= 512 px), and on non-square ones."""
import pytest
import torch

from oracle import reference_port as rp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.fail("these tests need a CUDA device (run on the B200 box with -m gpu)")
    from fovea import ops as _ops
    _ops._lib.load()
    return _ops


def _grid(B, seed, g=80):
    xs, _ = rp.synthetic_saliency(B, gh=g, gw=g, seed=seed)
    filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(g, g, 45, 45)
    return rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, g, g, (g, g))[0].cuda().float().contiguous()


@pytest.mark.parametrize("H,W,nchan,tri", [(64, 64, 3, "device"), (128, 160, 51, "device"), (512, 512, 51, "device"),
                                           (520, 392, 51, "host"), (1024, 1024, 51, "device"), (2048, 2048, 51, "device"),
                                           (256, 256, 600, "device")])
def test_sparse_plan_equals_dense_plan(ops, H, W, nchan, tri):
    B = 3
    grid = _grid(B, seed=H * 7 + W)
    dense = ops.build_inverse_plan(grid, (H, W), nchan=nchan, triangulation=tri)
    sparse = ops.build_inverse_plan(grid, (H, W), nchan=nchan, triangulation=tri, dense_winner=False)
    assert dense.winner is not None and sparse.winner is None
    assert torch.equal(dense.npts, sparse.npts)
    for b in range(B):
        n = int(dense.npts[b])
        assert torch.equal(dense.pts[b, :n], sparse.pts[b, :n]), "sites differ"
        assert torch.equal(dense.src[b, :n], sparse.src[b, :n]), "table rows of the sites differ"
        T = int(dense.ntri[b])
        assert T == int(sparse.ntri[b])
        assert torch.equal(dense.mesh[b, :T].view(torch.int16), sparse.mesh[b, :T].view(torch.int16))
    assert torch.equal(dense.loc, sparse.loc), "per-pixel source maps differ"


def test_sparse_plan_with_nan_and_out_of_range_targets(ops):
    """Grid entries that are NaN or map outside the canvas index nothing (models/models.py:644-651) on both paths."""
    grid = _grid(2, seed=5)
    grid[0, 3, 5, 0] = float("nan")
    grid[0, 10, 10, 1] = 1.5
    grid[1, 0, 0, :] = -1.2
    dense = ops.build_inverse_plan(grid, (256, 256), nchan=51, triangulation="device")
    sparse = ops.build_inverse_plan(grid, (256, 256), nchan=51, triangulation="device", dense_winner=False)
    assert torch.equal(dense.npts, sparse.npts)
    for b in range(2):
        n = int(dense.npts[b])
        assert torch.equal(dense.pts[b, :n], sparse.pts[b, :n]) and torch.equal(dense.src[b, :n], sparse.src[b, :n])
    assert torch.equal(dense.loc, sparse.loc)


def test_sparse_plan_on_the_yaml_lattice_64x128(ops):
    """config/deform.yaml:42's (64,128) saliency = 8 196 keys per frame: the 16 384-key sort, the Delaunay kernel's large
    layout and the marker raster on a 1024 x 2048 canvas, sparse against dense."""
    B, gh, gw, Rx, Ry, H, W = 2, 64, 128, 15, 30, 1024, 2048
    xs, _ = rp.synthetic_saliency(B, gh, gw, seed=64)
    filt, P = rp.gaussian_filter_weight(Rx, Ry, Rx), rp.p_basis(gh, gw, Rx, Ry)
    grid = rp.create_grid(rp.pad_saliency(xs, Rx, Ry), filt, P, gh, gw, (gh, gw))[0].cuda().float().contiguous()
    dense = ops.check_plan(ops.build_inverse_plan(grid, (H, W), nchan=3, triangulation="device"))
    sparse = ops.check_plan(ops.build_inverse_plan(grid, (H, W), nchan=3, triangulation="device", dense_winner=False))
    assert torch.equal(dense.npts, sparse.npts) and (dense.npts > 6600).all()
    for b in range(B):
        n = int(dense.npts[b])
        assert torch.equal(dense.pts[b, :n], sparse.pts[b, :n]) and torch.equal(dense.src[b, :n], sparse.src[b, :n])
    assert torch.equal(dense.loc, sparse.loc)

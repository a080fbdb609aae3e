"""Pin the CPU oracle (oracle/reference_port.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import reference_port as rp

GRID_CASES = ["grid_80_R45", "grid_80_R45_seg520", "grid_40x80_R12_reflect", "grid_40x80_R12_zero", "grid_32_R10_eval"]


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


@pytest.mark.parametrize("name", GRID_CASES)
def test_filter_and_pbasis_match_reference(golden_dir, name):
    g = _load(golden_dir, name)
    Rx, Ry, R = int(g["Rx"]), int(g["Ry"]), int(g["R"])
    gh, gw = g["xs"].shape[-2:]
    filt = rp.gaussian_filter_weight(Rx, Ry, R)
    assert np.array_equal(filt.numpy(), g["filt"])
    assert np.array_equal(rp.p_basis(gh, gw, Rx, Ry).numpy(), g["P_basis"])
    # rank-1 (SURVEY surprise 6): the dense filter is an outer product of two 1-D profiles
    u, s, vt = np.linalg.svd(g["filt"].astype(np.float64))
    assert s[1] / s[0] < 1e-6


@pytest.mark.parametrize("name", GRID_CASES)
def test_create_grid_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    xs = torch.from_numpy(g["xs"])
    Rx, Ry = int(g["Rx"]), int(g["Ry"])
    gh, gw = xs.shape[-2:]
    filt, P = torch.from_numpy(g["filt"]), torch.from_numpy(g["P_basis"])
    xs_hm = rp.pad_saliency(xs, Rx, Ry, str(g["pad_mode"]))
    task, task_eval, rate = tuple(g["task"]), tuple(g["task_eval"]), int(g["rate"])
    grid, grid_y = rp.create_grid(xs_hm, filt, P, gh, gw, task, task_eval, rate)
    assert np.array_equal(grid.numpy(), g["grid"])
    assert np.array_equal(grid_y.numpy(), g["grid_y"])
    seg = tuple(int(s) for s in g["segSize"])
    grid2, grid_inv = rp.create_grid(xs_hm, filt, P, gh, gw, task, task_eval, rate, segSize=seg, x_inv=1 - xs_hm,
                                     tie="torch")
    assert np.array_equal(grid2.numpy(), g["grid_infer"])
    # Deterministic tie rule ("max"): identical to the reference wherever no two low-res nodes collide; on
    # collisions the reference's winner is undefined (it is merely ONE of the candidates).
    B, h, w = grid2.shape[:3]
    gi_max = rp.grid_inverse(grid2, seg, tie="max").numpy()
    assert np.array_equal(np.isnan(gi_max), np.isnan(g["grid_inv"]))
    u = (((grid2[..., 0] + 1) / 2) * (seg[1] - 1)).int().long().view(B, -1)
    v = (((grid2[..., 1] + 1) / 2) * (seg[0] - 1)).int().long().view(B, -1)
    counts = torch.zeros(B, seg[0] * seg[1], dtype=torch.int64).scatter_add_(1, v * seg[1] + u, torch.ones_like(u))
    single = (counts.view(B, *seg) == 1).numpy()
    assert single.sum() > 0
    assert np.array_equal(gi_max[single], g["grid_inv"][single])
    assert np.array_equal(grid_inv.numpy()[single], g["grid_inv"][single])   # replayed torch ops, tie-free pixels
    assert np.array_equal(np.isnan(grid_inv.numpy()), np.isnan(g["grid_inv"]))
    ref_j = np.rint((g["grid_inv"][..., 0] + 1) / 2 * w)
    ref_i = np.rint((g["grid_inv"][..., 1] + 1) / 2 * h)
    multi = np.argwhere(counts.view(B, *seg).numpy() > 1)
    for b, y, x in multi[:200]:
        cands = torch.where((u[b] == x) & (v[b] == y))[0].tolist()
        assert int(ref_i[b, y, x] * w + ref_j[b, y, x]) in cands
        assert rp.grid_inverse_winner(grid2, seg)[b, y, x] == max(cands)


@pytest.mark.parametrize("name,grid_name", [("inverse_80_to_128", "grid_80_R45"), ("inverse_80_to_520", "grid_80_R45_seg520")])
def test_inverse_path_matches_reference(golden_dir, name, grid_name):
    g = _load(golden_dir, name)
    gg = _load(golden_dir, grid_name)
    pred, grid = torch.from_numpy(g["pred"]), torch.from_numpy(g["grid"])
    seg = tuple(int(s) for s in gg["segSize"])
    # the reference's duplicate-target winners are thread-timing dependent even on CPU, so the golden's own
    # grid_inv is the input here: this pins A8 + A9 exactly, independent of the tie rule
    gi = torch.from_numpy(gg["grid_inv"])
    ps_nan = rp.inverse_sample(pred, gi)
    assert np.array_equal(ps_nan.numpy(), g["pred_sampled_nan"], equal_nan=True)
    ps = ps_nan.clone()
    for n in range(ps.shape[0]):
        ps[n] = rp.fill_missing_values_tensor(ps[n])
    ref = g["pred_sampled"]
    assert np.array_equal(np.isnan(ps.numpy()), np.isnan(ref))
    np.testing.assert_allclose(ps.numpy(), ref, rtol=0, atol=1e-6, equal_nan=True)
    # A5 through the reference's own F.grid_sample call, and the explicit float64 restatement of the op
    gen = torch.Generator().manual_seed(int(g["x_seed"]))
    x = torch.rand(grid.shape[0], 3, int(g["x_hw"][0]), int(g["x_hw"][1]), generator=gen)
    xs_ = rp.grid_sample(x, grid)
    assert np.array_equal(xs_.numpy(), g["x_sampled"])
    # exact (float64) restatement: fp32 pixel-coordinate rounding is W*2^-24 px, times |d img/dx| <= 1 here
    Wx = x.shape[-1]
    np.testing.assert_allclose(rp.grid_sample_explicit(x.numpy(), grid.numpy()), g["x_sampled"], rtol=1e-5,
                               atol=4 * Wx * 2.0 ** -24)


def test_interp2d_matches_reference(golden_dir):
    g = _load(golden_dir, "interp2d_64x48")
    h, w = (int(v) for v in g["hw"])
    out = rp.interp2d_forward(torch.from_numpy(g["points"]), torch.from_numpy(g["values"]), h, w)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=1e-6)


def test_a8_is_two_by_two_box_average():
    """sampling pred at grid_inv coordinates equals the zero-padded 2x2 box mean."""
    pred = rp.synthetic_pred(1, 3, 16, 20, seed=3)
    grid = torch.stack(torch.meshgrid(torch.linspace(-1, 1, 16), torch.linspace(-1, 1, 20), indexing="ij")[::-1], -1)[None]
    gi = rp.grid_inverse(grid, (40, 44))
    ps = rp.inverse_sample(pred, gi)
    win = rp.grid_inverse_winner(grid, (40, 44))[0]
    pp = torch.nn.functional.pad(pred, (1, 0, 1, 0))
    box = 0.25 * (pp[..., :-1, :-1] + pp[..., :-1, 1:] + pp[..., 1:, :-1] + pp[..., 1:, 1:])
    ys, xs = torch.where(win >= 0)
    got = ps[0, :, ys, xs]
    want = box[0].reshape(3, -1)[:, win[ys, xs]]
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name,src", [("nearest_80_to_128", "inverse_80_to_128"), ("nearest_80_to_520", "inverse_80_to_520")])
def test_nearest_fill_matches_reference(golden_dir, name, src):
    """interp_mode='nearest' (models/models.py:213-272) of the oracle == the unmodified reference, bit for bit."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    s = np.load(os.path.join(golden_dir, src + ".npz"))
    want = torch.from_numpy(g["pred_sampled_nearest"])
    for n in range(want.shape[0]):
        got = rp.fill_missing_values_nearest(torch.from_numpy(s["pred_sampled_nan"][n]).clone())
        assert torch.equal(got, want[n])


def test_cv2_dilate_reads_chw_as_rows_cols_channels():
    """The reading of getPixelsForInterp_NB that the 'nearest' kernels implement: cv2.dilate on a [C,H,W] array spans the
    class and row axes, never the column axis (so with one NaN pattern in every class only vertical neighbours count)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    a = (rng.rand(5, 40, 37) > 0.9).astype("uint8")
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3))
    d = cv2.dilate(a, k, borderType=cv2.BORDER_CONSTANT, borderValue=int(0))
    o = a.copy()
    o[1:] |= a[:-1]; o[:-1] |= a[1:]
    o[:, 1:] |= a[:, :-1]; o[:, :-1] |= a[:, 1:]
    assert np.array_equal(d, o)
    b = np.repeat(a[:1], 5, 0)
    v = b.copy()
    v[:, 1:] |= b[:, :-1]; v[:, :-1] |= b[:, 1:]
    assert np.array_equal(cv2.dilate(b, k, borderType=cv2.BORDER_CONSTANT, borderValue=int(0)), v)


@pytest.mark.parametrize("name", ["module_256_lowres", "module_256_upsample"])
def test_saliency_input_matches_reference_module(golden_dir, name):
    """A0 (models/models.py:684-705): the oracle's focus map + x_low against the tensor the UNMODIFIED reference module
    handed to its saliency network (captured by make_golden.py --module with a forward pre-hook)."""
    from tiny_nets import synthetic_batch
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    feed = synthetic_batch(int(g["B"]), int(g["H"]), int(g["W"]), int(g["seed"]))
    got = rp.saliency_input(feed["img_data"], feed["focus_point"], (80, 80))
    np.testing.assert_allclose(got.numpy(), g["x_low"], rtol=0, atol=1e-7)


@pytest.mark.parametrize("name", ["unsampler_24_to_96x128", "unsampler_40x64_to_520"])
def test_deformed_unsampler_matches_reference(golden_dir, name):
    """the oracle's restatement of DynamicFocus's deformed_unsampler against the output of the unmodified
    reference functions (tests/golden/make_golden_dynamicfocus.py), bit for bit -- including the duplicate-target rule
    (last write = largest node index) and SciPy's own choice among equidistant pixels."""
    g = _load(golden_dir, name)
    H, W = int(g["H"]), int(g["W"])
    coords = rp.int_round_scale_grid(torch.from_numpy(g["grid"]).clone(), H, W)
    assert np.array_equal(coords.numpy(), g["coords"])
    out = rp.deformed_unsampler(torch.from_numpy(g["labels"]), coords, H, W)
    assert np.array_equal(out.numpy(), g["out"])

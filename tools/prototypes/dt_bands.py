"""Could the Delaunay kernel's start mesh use BANDS -- consecutive occupied pixel rows whose sites all belong to ONE lattice
row merged into one x-sorted chain, so that chains stay y-separated and the closed-form zipper + pocket clipping apply
unchanged -- instead of single pixel rows?  Measured: no, the rule almost never fires (see __main__)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools', 'prototypes'))
import numpy as np, torch
from oracle import reference_port as rp


def frame(seed, H, W):
    xs, _ = rp.synthetic_saliency(1, seed=seed)
    filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
    grid, _ = rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))
    ps = rp.inverse_sample(rp.synthetic_pred(1, 1, seed=seed), rp.grid_inverse(grid, (H, W)))
    mask, inv = rp.pixels_for_interp(ps[0]); rr, cc = torch.where(mask[0])
    win = rp.grid_inverse_winner(grid, (H, W))[0]
    pts = np.stack([rr.numpy(), cc.numpy()], 1).astype(np.int64)
    return pts, win[rr, cc].numpy()          # sorted row-major; node = -1 at forced corners without a node


def make_chains(pts, node, merge):
    rows = pts[:, 0]
    starts = np.flatnonzero(np.r_[True, rows[1:] != rows[:-1]]); ends = np.r_[starts[1:], len(pts)]
    lat = np.where(node >= 0, node // 80, -1)
    key = []                                   # lattice row of a pixel row if it is unique, else -1 (never merged)
    for s, e in zip(starts, ends):
        l = np.unique(lat[s:e]); key.append(int(l[0]) if len(l) == 1 and l[0] >= 0 else -1)
    chains, cur = [], None
    for r, (s, e) in enumerate(zip(starts, ends)):
        idx = np.arange(s, e)
        if merge and cur is not None and key[r] >= 0 and key[r] == cur[0]:
            cand = np.concatenate([cur[1], idx]); cand = cand[np.lexsort((pts[cand, 0], pts[cand, 1]))]
            if (np.diff(pts[cand, 1]) > 0).all():       # strictly x-monotone: keep merging
                cur = (cur[0], cand); continue
        if cur is not None: chains.append(cur[1])
        cur = (key[r], idx)
    chains.append(cur[1])
    return chains


if __name__ == "__main__":
    # Result (round 2): the rule almost never fires -- 951 -> 899, 804 -> 796, 655 -> 638, 1018 -> 1018 chains on four
    # 1024^2 frames (1 299 -> 1 268 at 2048^2): the deformation tilts a lattice row by more than the row spacing, so a pixel
    # row nearly always holds sites of two or three lattice rows and consecutive lattice rows are NOT y-separated.  A start
    # mesh from the lattice therefore needs general (non-monotone) polygon handling, not a variant of the row zipper.
    for (H, W, seed) in [(1024, 1024, s) for s in range(4)] + [(2048, 2048, 20), (256, 256, 30)]:
        pts, node = frame(seed, H, W)
        n0, n1 = len(make_chains(pts, node, False)), len(make_chains(pts, node, True))
        print(f"{H}^2 seed {seed:2d}: sites {len(pts)}: {n0} pixel-row chains -> {n1} chains after merging single-lattice-row bands")

"""Parallel Lawson rounds started from the lattice-row chains INCLUDING their zero-area / inverted triangles (two-sided
legality test, a flip must create two positive triangles): 17-50 rounds and 2-4 k flips on six frames (pixel-row start:
80-300 rounds, 15-35 k flips), but runs of collinear sites shared by two chains leave zero-area triangles that no flip
removes -- the dense core needs the pixel-row construction."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools', 'prototypes'))
import numpy as np, torch
from dt_proto import orient, incircle, build_adjacency
from dt_lattice_chains import chains_zipper
from oracle import reference_port as rp
from scipy.spatial import Delaunay

def flips_two_sided(pts, tris, nb, seed=0, max_rounds=400):
    """Parallel Lawson rounds; an edge is illegal if EITHER adjacent non-degenerate triangle has the opposite vertex strictly
    inside its circumcircle, or if one of the two triangles is degenerate (zero area) and the edge is its longest one."""
    rng = np.random.default_rng(seed); T = len(tris); rounds = total = 0
    while rounds < max_rounds:
        tt = np.repeat(np.arange(T), 3); kk = np.tile(np.arange(3), T); uu = nb[tt, kk]
        m = uu > tt; tt, kk, uu = tt[m], kk[m], uu[m]
        a = pts[tris[tt, kk]]; b = pts[tris[tt, (kk + 1) % 3]]; c = pts[tris[tt, (kk + 2) % 3]]
        k2 = np.argmax(nb[uu] == tt[:, None], axis=1); d = pts[tris[uu, k2]]
        ot = orient(a, b, c); ou = orient(d, c, b)
        bad = ((ot > 0) & (incircle(a, b, c, d) > 0)) | ((ou > 0) & (incircle(d, c, b, a) > 0))
        # a flip must produce two positively oriented triangles (a,b,d) and (a,d,c)
        ok = (orient(a, b, d) > 0) & (orient(a, d, c) > 0)
        degen = ((ot == 0) | (ou == 0)) & ok
        bad = (bad | degen) & ok
        tt, kk, uu, k2 = tt[bad], kk[bad], uu[bad], k2[bad]
        if len(tt) == 0: break
        pri = rng.permutation(len(tt)).astype(np.int64); owner = np.full(T, np.iinfo(np.int64).max)
        grp = [tt, uu, nb[tt, (kk + 1) % 3], nb[tt, (kk + 2) % 3], nb[uu, (k2 + 1) % 3], nb[uu, (k2 + 2) % 3]]
        for g in grp:
            o_ = g >= 0; np.minimum.at(owner, g[o_], pri[o_])
        win = np.ones(len(tt), bool)
        for g in grp:
            o_ = g >= 0; win &= (~o_) | (owner[np.where(o_, g, 0)] == pri)
        for t, k, u, ku in zip(tt[win], kk[win], uu[win], k2[win]):
            a_ = tris[t, k]; b_ = tris[t, (k + 1) % 3]; c_ = tris[t, (k + 2) % 3]; d_ = tris[u, ku]
            n_ab = nb[t, (k + 2) % 3]; n_ca = nb[t, (k + 1) % 3]
            iu_b = [i for i in range(3) if tris[u, i] == b_][0]; iu_c = [i for i in range(3) if tris[u, i] == c_][0]
            n_bd = nb[u, iu_c]; n_dc = nb[u, iu_b]
            tris[t] = (a_, b_, d_); nb[t] = (n_bd, u, n_ab); tris[u] = (a_, d_, c_); nb[u] = (n_dc, n_ca, t)
            if n_bd >= 0: nb[n_bd][nb[n_bd] == u] = t
            if n_ca >= 0: nb[n_ca][nb[n_ca] == t] = u
        total += int(win.sum()); rounds += 1
    return rounds, total

filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
for (H, W, seed) in [(1024, 1024, 7), (1024, 1024, 11), (1024, 1024, 3), (2048, 2048, 4), (1024, 1024, 21), (1024, 1024, 22)]:
    xs, _ = rp.synthetic_saliency(1, seed=seed)
    grid, _ = rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))
    ps = rp.inverse_sample(rp.synthetic_pred(1, 1, seed=seed), rp.grid_inverse(grid, (H, W)))
    mask, inv = rp.pixels_for_interp(ps[0]); rr, cc = torch.where(mask[0])
    win = rp.grid_inverse_winner(grid, (H, W))[0]; node = win[rr, cc].numpy()
    pts = np.stack([rr.numpy(), cc.numpy()], 1).astype(np.int64); keep = node >= 0; pts, node = pts[keep], node[keep]
    tris, o, lens = chains_zipper(pts, node // 80)
    ninv = int((o < 0).sum()); nzero = int((o == 0).sum())
    tris[o < 0] = tris[o < 0][:, [0, 2, 1]]
    nb, nbound = build_adjacency(tris)
    r, tot = flips_two_sided(pts, tris, nb)
    o2 = orient(pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]])
    area = int(o2.sum())
    print(f"{H}^2 seed {seed}: zero-area {nzero}, inverted {ninv} -> rounds {r}, flips {tot}; after flips: zero-area {int((o2 == 0).sum())}, "
          f"negative {int((o2 < 0).sum())}", flush=True)

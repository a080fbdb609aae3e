"""How many flip rounds would the Delaunay kernel save if a flip claimed only its TWO triangles instead of the six it
claims today (the two plus the four outer neighbours whose back links it rewrites)?  Simulation of the kernel's
random-priority rounds on the pixel-row start mesh of three frames."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools', 'prototypes'))
import numpy as np, torch
from dt_proto import incircle, build_adjacency, check_delaunay
from dt_start_experiments import zipper
from oracle import reference_port as rp


def rounds(pts, tris, nb, narrow, seed=0):
    rng = np.random.default_rng(seed); T = len(tris); r = 0; total = 0; hist = []
    while True:
        tt = np.repeat(np.arange(T), 3); kk = np.tile(np.arange(3), T); uu = nb[tt, kk]
        m = uu > tt; tt, kk, uu = tt[m], kk[m], uu[m]
        a = pts[tris[tt, kk]]; b = pts[tris[tt, (kk + 1) % 3]]; c = pts[tris[tt, (kk + 2) % 3]]
        k2 = np.argmax(nb[uu] == tt[:, None], axis=1); dpt = pts[tris[uu, k2]]
        bad = incircle(a, b, c, dpt) > 0
        tt, kk, uu, k2 = tt[bad], kk[bad], uu[bad], k2[bad]
        if len(tt) == 0: break
        hist.append(len(tt))
        pri = rng.permutation(len(tt)).astype(np.int64)
        owner = np.full(T, np.iinfo(np.int64).max)
        grp = [tt, uu] if narrow else [tt, uu, nb[tt, (kk + 1) % 3], nb[tt, (kk + 2) % 3], nb[uu, (k2 + 1) % 3], nb[uu, (k2 + 2) % 3]]
        for g in grp:
            ok = g >= 0; np.minimum.at(owner, g[ok], pri[ok])
        win = np.ones(len(tt), bool)
        for g in grp:
            ok = g >= 0; win &= (~ok) | (owner[np.where(ok, g, 0)] == pri)
        for t, k, u, ku in zip(tt[win], kk[win], uu[win], k2[win]):      # (sequential here: the back links are exact)
            a = tris[t, k]; b = tris[t, (k + 1) % 3]; c = tris[t, (k + 2) % 3]
            ku = [i for i in range(3) if tris[u, i] not in (b, c)][0]; d = tris[u, ku]
            n_ab = nb[t, (k + 2) % 3]; n_ca = nb[t, (k + 1) % 3]
            iu_b = [i for i in range(3) if tris[u, i] == b][0]; iu_c = [i for i in range(3) if tris[u, i] == c][0]
            n_bd = nb[u, iu_c]; n_dc = nb[u, iu_b]
            tris[t] = (a, b, d); nb[t] = (n_bd, u, n_ab); tris[u] = (a, d, c); nb[u] = (n_dc, n_ca, t)
            if n_bd >= 0: nb[n_bd][nb[n_bd] == u] = t
            if n_ca >= 0: nb[n_ca][nb[n_ca] == t] = u
        total += int(win.sum()); r += 1
    return r, total, hist


if __name__ == "__main__":
    for (H, W, seed) in [(1024, 1024, 3), (1024, 1024, 7), (2048, 2048, 4)]:
        xs, _ = rp.synthetic_saliency(1, seed=seed)
        filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
        grid, _ = rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))
        ps = rp.inverse_sample(rp.synthetic_pred(1, 1, seed=seed), rp.grid_inverse(grid, (H, W)))
        mask, inv = rp.pixels_for_interp(ps[0]); rr, cc = torch.where(mask[0])
        pts = np.stack([rr.numpy(), cc.numpy()], 1).astype(np.int64)
        tris = zipper(pts, False); nb, _ = build_adjacency(tris)
        for narrow in (False, True):
            t0 = time.time()
            r, tot, hist = rounds(pts, tris.copy(), nb.copy(), narrow)
            print(f"{H}^2 seed {seed} claims={'2' if narrow else '6'}: rounds={r} flips={tot} illegal edges per round {hist[:5]} .. "
                  f"rounds with > 100 illegal edges: {sum(1 for h in hist if h > 100)}, with <= 15: {sum(1 for h in hist if h <= 15)}  ({time.time()-t0:.0f}s)", flush=True)

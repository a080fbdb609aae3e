import sys, time
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tools', 'prototypes'))
import numpy as np, torch
from dt_proto import orient, incircle, build_adjacency, check_delaunay
from oracle import reference_port as rp

def zipper(pts, greedy):
    rows = pts[:, 0]
    starts = np.flatnonzero(np.r_[True, rows[1:] != rows[:-1]]); ends = np.r_[starts[1:], len(pts)]
    tris = []
    for k in range(len(starts) - 1):
        t = np.arange(starts[k], ends[k]); b = np.arange(starts[k + 1], ends[k + 1])
        i = j = 0
        while i < len(t) - 1 or j < len(b) - 1:
            if j == len(b) - 1: adv_top = True
            elif i == len(t) - 1: adv_top = False
            elif not greedy: adv_top = pts[t[i + 1], 1] <= pts[b[j + 1], 1]
            else:
                # candidate triangles (t_i, t_{i+1}, b_j) vs (t_i, b_{j+1}, b_j): pick the one whose circumcircle excludes the other candidate
                a, c_, d_, e_ = pts[t[i]], pts[b[j]], pts[t[i + 1]], pts[b[j + 1]]
                # triangle ccw? rows increase downward: use incircle with orientation fix
                tri = np.array([a, d_, c_]); o = orient(tri[0], tri[1], tri[2])
                if o < 0: tri = tri[[0, 2, 1]]
                adv_top = incircle(tri[0], tri[1], tri[2], e_) <= 0
            if adv_top: tris.append((t[i], t[i + 1], b[j])); i += 1
            else: tris.append((t[i], b[j + 1], b[j])); j += 1
    for side in (0, 1):
        chain = list(starts if side == 0 else ends - 1); st = [chain[0]]
        for v in chain[1:]:
            while len(st) >= 2:
                a, b_ = st[-2], st[-1]; o = orient(pts[a], pts[b_], pts[v])
                if (o > 0) if side == 0 else (o < 0):
                    tris.append((a, b_, v) if side == 0 else (a, v, b_)); st.pop()
                else: break
            st.append(v)
    tris = np.array(tris, dtype=np.int64)
    o = orient(pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]])
    assert (o != 0).all()
    tris[o < 0] = tris[o < 0][:, [0, 2, 1]]
    return tris

def flip_rounds_random(pts, tris, nb, seed=0):
    rng = np.random.default_rng(seed); T = len(tris); rounds = 0; total = 0; hist = []
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
        grp = [tt, uu, nb[tt, (kk + 1) % 3], nb[tt, (kk + 2) % 3], nb[uu, (k2 + 1) % 3], nb[uu, (k2 + 2) % 3]]
        for g in grp:
            ok = g >= 0; np.minimum.at(owner, g[ok], pri[ok])
        win = np.ones(len(tt), bool)
        for g in grp:
            ok = g >= 0; win &= (~ok) | (owner[np.where(ok, g, 0)] == pri)
        for t, k, u, ku in zip(tt[win], kk[win], uu[win], k2[win]):
            a = tris[t, k]; b = tris[t, (k + 1) % 3]; c = tris[t, (k + 2) % 3]; d = tris[u, ku]
            n_ab = nb[t, (k + 2) % 3]; n_ca = nb[t, (k + 1) % 3]
            iu_b = [i for i in range(3) if tris[u, i] == b][0]; iu_c = [i for i in range(3) if tris[u, i] == c][0]
            n_bd = nb[u, iu_c]; n_dc = nb[u, iu_b]
            tris[t] = (a, b, d); nb[t] = (n_bd, u, n_ab); tris[u] = (a, d, c); nb[u] = (n_dc, n_ca, t)
            if n_bd >= 0: nb[n_bd][nb[n_bd] == u] = t
            if n_ca >= 0: nb[n_ca][nb[n_ca] == t] = u
        total += int(win.sum()); rounds += 1
    return rounds, total, hist

if __name__ == "__main__":
  for (H, W, seed) in [(1024, 1024, 3), (1024, 1024, 7), (2048, 2048, 4)]:
      xs, _ = rp.synthetic_saliency(1, seed=seed)
      filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
      grid, _ = rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))
      ps = rp.inverse_sample(rp.synthetic_pred(1, 1, seed=seed), rp.grid_inverse(grid, (H, W)))
      mask, inv = rp.pixels_for_interp(ps[0]); rr, cc = torch.where(mask[0])
      pts = np.stack([rr.numpy(), cc.numpy()], 1).astype(np.int64)
      for greedy in (False, True):
          t0 = time.time(); tris = zipper(pts, greedy); nb, _ = build_adjacency(tris)
          r, tot, hist = flip_rounds_random(pts, tris.copy(), nb.copy())
          print(f"{H}^2 seed {seed} greedy={greedy}: N={len(pts)} T={len(tris)} rounds={r} flips={tot} illegal edges per round (first 6) {hist[:6]} ... tail>{sum(1 for h in hist if h<=15)} rounds with <=15  ({time.time()-t0:.0f}s)", flush=True)

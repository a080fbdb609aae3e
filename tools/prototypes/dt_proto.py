"""Prototype of the device Delaunay algorithm (row strips + pockets + parallel Lawson flips) in NumPy."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np


def orient(a, b, c):
    return (b[..., 1] - a[..., 1]) * (c[..., 0] - a[..., 0]) - (b[..., 0] - a[..., 0]) * (c[..., 1] - a[..., 1])


def incircle(a, b, c, d):
    """>0 if d strictly inside circumcircle of ccw (a,b,c) [x=col,y=row]."""
    ax, ay = a[..., 1] - d[..., 1], a[..., 0] - d[..., 0]
    bx, by = b[..., 1] - d[..., 1], b[..., 0] - d[..., 0]
    cx, cy = c[..., 1] - d[..., 1], c[..., 0] - d[..., 0]
    return ((ax * ax + ay * ay) * (bx * cy - by * cx) - (bx * bx + by * by) * (ax * cy - ay * cx)
            + (cx * cx + cy * cy) * (ax * by - ay * bx))


def initial_triangulation(pts):
    """pts [N,2] (row,col) int64 sorted row-major, corners included. Returns tris [T,3] ccw (orient>0)."""
    rows = pts[:, 0]
    starts = np.flatnonzero(np.r_[True, rows[1:] != rows[:-1]])
    ends = np.r_[starts[1:], len(pts)]
    tris = []
    # strips
    for k in range(len(starts) - 1):
        t = np.arange(starts[k], ends[k]); b = np.arange(starts[k + 1], ends[k + 1])
        i = j = 0
        while i < len(t) - 1 or j < len(b) - 1:
            if j == len(b) - 1 or (i < len(t) - 1 and pts[t[i + 1], 1] <= pts[b[j + 1], 1]):
                tris.append((t[i], t[i + 1], b[j])); i += 1
            else:
                tris.append((t[i], b[j + 1], b[j])); j += 1
    # pockets: left chain (first of each row), right chain (last of each row)
    for side in (0, 1):
        chain = list(starts if side == 0 else ends - 1)
        # monotone mountain ear clipping (sequential here); interior side: left pocket lies at smaller col
        st = [chain[0]]
        for v in chain[1:]:
            while len(st) >= 2:
                a, b_ = st[-2], st[-1]
                o = orient(pts[a], pts[b_], pts[v])
                # left side: chain goes downward (row increasing); pocket is to the left (smaller col).
                # vertex b_ is an ear if the turn a->b_->v bulges away from the pocket, i.e. pocket-side convex
                if (o > 0) if side == 0 else (o < 0):
                    tris.append((a, b_, v) if side == 0 else (a, v, b_)); st.pop()
                else:
                    break
            st.append(v)
    tris = np.array(tris, dtype=np.int64)
    o = orient(pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]])
    assert (o != 0).all(), "degenerate triangle in initial triangulation"
    neg = o < 0
    tris[neg] = tris[neg][:, [0, 2, 1]]
    return tris


def build_adjacency(tris):
    T = len(tris)
    nb = -np.ones((T, 3), dtype=np.int64)
    d = {}
    for t in range(T):
        for k in range(3):
            a, b = tris[t, (k + 1) % 3], tris[t, (k + 2) % 3]
            key = (min(a, b), max(a, b))
            if key in d:
                t2, k2 = d.pop(key)
                nb[t, k] = t2; nb[t2, k2] = t
            else:
                d[key] = (t, k)
    return nb, len(d)


def flip_rounds(pts, tris, nb, max_rounds=100000):
    T = len(tris)
    rounds = 0; total = 0
    while True:
        # candidate edges: (t,k) with neighbor u>t... evaluate all
        cand = []
        tt = np.repeat(np.arange(T), 3); kk = np.tile(np.arange(3), T)
        uu = nb[tt, kk]
        m = uu > tt
        tt, kk, uu = tt[m], kk[m], uu[m]
        a = pts[tris[tt, kk]]; b = pts[tris[tt, (kk + 1) % 3]]; c = pts[tris[tt, (kk + 2) % 3]]
        # opposite vertex in u: the vertex of u not in edge (b,c)
        k2 = np.argmax(nb[uu] == tt[:, None], axis=1)
        dpt = pts[tris[uu, k2]]
        bad = incircle(a, b, c, dpt) > 0
        tt, kk, uu, k2 = tt[bad], kk[bad], uu[bad], k2[bad]
        if len(tt) == 0:
            break
        # independent set: claim triangles t,u and their outer neighbors; priority = edge id (min wins)
        owner = np.full(T, np.iinfo(np.int64).max)
        eid = tt * 3 + kk
        grp = [tt, uu, nb[tt, (kk + 1) % 3], nb[tt, (kk + 2) % 3], nb[uu, (k2 + 1) % 3], nb[uu, (k2 + 2) % 3]]
        for g in grp:
            ok = g >= 0
            np.minimum.at(owner, g[ok], eid[ok])
        win = np.ones(len(tt), bool)
        for g in grp:
            ok = g >= 0
            win &= (~ok) | (owner[np.where(ok, g, 0)] == eid)
        tt, kk, uu, k2 = tt[win], kk[win], uu[win], k2[win]
        # perform flips sequentially (they are independent)
        for t, k, u, ku in zip(tt, kk, uu, k2):
            a = tris[t, k]; b = tris[t, (k + 1) % 3]; c = tris[t, (k + 2) % 3]; d = tris[u, ku]
            # t = (a,b,c), u shares edge (b,c) with apex d. new: t=(a,b,d), u=(a,d,c)
            n_ab = nb[t, (k + 2) % 3]  # opposite c : edge (a,b)
            n_ca = nb[t, (k + 1) % 3]  # opposite b : edge (c,a)
            # in u: vertices order ... find edges
            iu_b = [i for i in range(3) if tris[u, i] == b][0]
            iu_c = [i for i in range(3) if tris[u, i] == c][0]
            n_bd = nb[u, iu_c]  # opposite c in u: edge (d,b)
            n_dc = nb[u, iu_b]  # opposite b in u: edge (c,d)
            tris[t] = (a, b, d); nb[t] = (n_bd, u, n_ab)
            tris[u] = (a, d, c); nb[u] = (n_dc, n_ca, t)
            if n_bd >= 0:
                nb[n_bd][nb[n_bd] == u] = t
            if n_ca >= 0:
                nb[n_ca][nb[n_ca] == t] = u
        total += len(tt); rounds += 1
        if rounds >= max_rounds:
            break
    return rounds, total


def check_delaunay(pts, tris, nb):
    T = len(tris)
    o = orient(pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]])
    assert (o > 0).all()
    for k in range(3):
        u = nb[:, k]; m = u >= 0
        t = np.flatnonzero(m); u = u[m]
        k2 = np.argmax(nb[u] == t[:, None], axis=1)
        assert (nb[u, k2] == t).all()
        inc = incircle(pts[tris[t, 0]], pts[tris[t, 1]], pts[tris[t, 2]], pts[tris[u, k2]])
        assert (inc <= 0).all(), "not Delaunay"
    return True


if __name__ == "__main__":
    import torch
    from oracle import reference_port as rp
    for (H, W, seed) in [(128, 128, 1), (256, 256, 2), (1024, 1024, 3), (2048, 2048, 4)]:
        xs, _ = rp.synthetic_saliency(1, seed=seed)
        filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
        grid, _ = rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))
        pred = rp.synthetic_pred(1, 1, seed=seed)
        ps = rp.inverse_sample(pred, rp.grid_inverse(grid, (H, W)))
        mask, inv = rp.pixels_for_interp(ps[0])
        rr, cc = torch.where(mask[0]); pts = np.stack([rr.numpy(), cc.numpy()], 1).astype(np.int64)
        t0 = time.time()
        tris = initial_triangulation(pts)
        nb, nbound = build_adjacency(tris)
        area2 = orient(pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]]).sum()
        print(f"{H}x{W}: N={len(pts)} rows={len(np.unique(pts[:,0]))} T={len(tris)} boundary edges={nbound} "
              f"area ok={area2 == 2*(H-1)*(W-1)}  init {time.time()-t0:.1f}s")
        t0 = time.time()
        rounds, total = flip_rounds(pts, tris, nb)
        print(f"   flips: rounds={rounds} total flips={total} ({time.time()-t0:.1f}s)")
        check_delaunay(pts, tris, nb)
        from scipy.spatial import Delaunay
        sd = Delaunay(pts.astype(float))
        print("   scipy T =", len(sd.simplices), " ours T =", len(tris))

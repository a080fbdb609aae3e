"""How does the reference's Qhull split co-circular cells -- and can the device Delaunay kernel imitate it?  (CPU experiment;
needs oracle/_ref/libqhull_ref.so = the reference's own vendored Qhull 2019.1, `make -C oracle`.)

Result: every cell of >= 4 co-circular sites comes out of `Qt` as a FAN around the vertex with the
largest Qhull vertex id, i.e. the site the incremental hull inserted last; and that vertex is not predictable from the
cell: nine local rules each hit it about as often as a blind guess.

    python tools/prototypes/qt_fan_rule.py
"""
import collections, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import qhull_ref, reference_port as rp


def sites(seed, H=1024, W=1024):
    xs, _ = rp.synthetic_saliency(1, seed=seed)
    filt, P = rp.gaussian_filter_weight(45, 45, 45), rp.p_basis(80, 80, 45, 45)
    grid, _ = rp.create_grid(rp.pad_saliency(xs, 45, 45), filt, P, 80, 80, (80, 80))
    gi = rp.grid_inverse(grid, (H, W))
    mask, _ = rp.pixels_for_interp(gi[0, :, :, 0][None])
    return torch.stack(torch.where(mask[0]), 1).numpy()


tot = collections.Counter()
for seed in range(3):
    pts = sites(seed)
    simp, owner, vid = qhull_ref.delaunay_probe(pts)
    cells = collections.defaultdict(list)
    for t in np.flatnonzero(owner != -1):
        cells[int(owner[t])].append(t)
    cen = pts.mean(0)
    hits, fans, sizes = collections.Counter(), 0, collections.Counter()
    for ts in cells.values():
        verts = sorted(set(simp[ts].ravel().tolist()))
        common = set.intersection(*[set(simp[t].tolist()) for t in ts])
        apex = max(verts, key=lambda p: vid[p])
        fans += apex in common
        sizes[len(verts)] += 1
        rules = {"max index": max(verts), "min index": min(verts),
                 "max row": max(verts, key=lambda p: (pts[p][0], pts[p][1])), "max col": max(verts, key=lambda p: (pts[p][1], pts[p][0])),
                 "min col": min(verts, key=lambda p: (pts[p][1], pts[p][0])),
                 "farthest from centroid": max(verts, key=lambda p: ((pts[p] - cen) ** 2).sum()),
                 "nearest to centroid": min(verts, key=lambda p: ((pts[p] - cen) ** 2).sum()),
                 "max lift": max(verts, key=lambda p: (pts[p] ** 2).sum()), "min lift": min(verts, key=lambda p: (pts[p] ** 2).sum())}
        for k, p in rules.items():
            hits[k] += p == apex
    n = len(cells)
    print(f"seed {seed}: {len(pts)} sites, {len(simp)} triangles, {n} co-circular cells (sizes {dict(sizes)}), "
          f"fans around the largest vertex id: {fans}/{n}")
    print("   local rules that pick the apex:", {k: f"{v / n:.3f}" for k, v in hits.items()})
    tot["cells"] += n; tot["fans"] += fans
print(f"total: {tot['fans']} / {tot['cells']} cells are fans around the last-inserted vertex")

"""Writes tests/golden/zones_outlier.json: a small cloud with the answers of plain numpy boolean logic (zone slicing =
chains of inclusive float32 PassThrough windows) and of an O(n^2) numpy count (radius outlier removal: float32 squared
distance (dx*dx + dy*dy) + dz*dz strictly below (float)(r*r), the point itself counted, keep iff count > min_pts) --
produced without the C++ oracle and without the CUDA path, so the file pins both. The reference ships no golden vectors
(parity unpinned). Run: python tests/golden/make_aux_fixtures.py
"""
import json
import os

import numpy as np

F = np.float32


def r9(v):
    v = float(v)
    return "nan" if v != v else float("%.9g" % v)


rng = np.random.default_rng(77)
n = 400
x = np.column_stack([rng.uniform(-2, 6, n), rng.uniform(-1, 1, n), rng.uniform(-0.6, 1.2, n), rng.uniform(0, 255, n)]).astype(F)
x[:60, :3] = (x[60:120, :3] + rng.normal(0, 0.06, (60, 3))).astype(F)   # close pairs, some inside, some outside the radius
x[7, 0] = F(2.0); x[8, 0] = F(4.0)                                       # points exactly on shared window ends
x[9, 2] = F(0.5); x[10, 2] = F(0.51)                                     # on the z window end / inside the 0.01 gap
x[11, 1] = np.nan                                                        # non-finite: passes no stage, never kept

zones = [[(0, 0.0, 2.0, 0), (2, -0.5, 0.5, 0)], [(0, 0.0, 2.0, 0), (2, float(F(np.float64(F(0.5)) + 0.01)), 1.0, 0)],
         [(0, 2.0, 4.0, 0), (2, -0.5, 0.5, 0)], [(0, 2.0, 4.0, 0), (1, -0.25, 0.25, 1)], []]
zone_idx = []
for chain in zones:
    keep = np.ones(n, bool)
    for axis, lo, hi, neg in chain:
        v = x[:, axis]
        with np.errstate(invalid="ignore"):
            inside = (v >= F(lo)) & (v <= F(hi))
        fin = np.isfinite(x[:, 0]) & np.isfinite(x[:, 1]) & np.isfinite(x[:, 2])
        keep &= fin & ((~inside & np.isfinite(v)) if neg else inside)
    zone_idx.append([int(i) for i in np.nonzero(keep)[0]])

radius = float(F(0.15))
r2 = F(radius * radius)
fin = np.isfinite(x[:, :3]).all(axis=1)
d = x[:, None, :3] - x[None, :, :3]
with np.errstate(invalid="ignore"):
    acc = (d[..., 0] * d[..., 0]).astype(F)
    acc = (acc + (d[..., 1] * d[..., 1]).astype(F)).astype(F)
    acc = (acc + (d[..., 2] * d[..., 2]).astype(F)).astype(F)
    cnt = ((acc < r2) & fin[None, :] & fin[:, None]).sum(axis=1)
outlier = {str(m): [int(i) for i in np.nonzero(fin & (cnt > m))[0]] for m in (0, 1, 2)}

doc = {"source": "numpy boolean windows / O(n^2) float32 count (tests/golden/make_aux_fixtures.py)",
       "xyzi": [[r9(v) for v in row] for row in x], "zones": zones, "zone_indices": zone_idx,
       "radius": r9(radius), "outlier_kept_by_min_pts": outlier}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "zones_outlier.json"), "w") as f:
    json.dump(doc, f)
print("wrote zones_outlier.json", [len(z) for z in zone_idx], {k: len(v) for k, v in outlier.items()})

"""Writes tests/golden/plane_ransac.json: small RANSAC ground-plane cases with the answers of the numpy restatement
(oracle/np_oracle.py: PCL 1.8.1's fixed-seed sampler, plane model, stopping rule; no refit) -- produced without the C++
oracle and without the CUDA path, so the file pins both. The raw engine outputs come from numpy's own MT19937
(init_genrand seeding). The reference ships no golden vectors (parity unpinned). Run: python tests/golden/make_plane_fixture.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import np_oracle  # noqa: E402


def r9(v):
    return float("%.9g" % float(v))


def scene(seed, n, frac):
    rng = np.random.default_rng(seed)
    g = int(n * frac)
    p = np.zeros((n, 4), np.float32)
    p[:, 0] = rng.uniform(-30, 30, n)
    p[:, 1] = rng.uniform(-10, 10, n)
    p[:g, 2] = -1.8 + 0.02 * p[:g, 0] + rng.normal(0, 0.03, g)
    p[g:, 2] = rng.uniform(-1.5, 1.0, n - g)
    p[:, 3] = np.arange(n)
    return p[rng.permutation(n)]


cases = []
for seed, n, frac, thr, prob, max_it, order in ((1, 48, 0.75, 0.3, 0.99, 1000, 0), (2, 64, 0.4, 0.1, 0.99, 1000, 1),
                                                (3, 40, 0.6, 0.3, 0.999, 25, 2), (4, 3, 1.0, 0.3, 0.99, 1000, 0)):
    x = scene(seed, n, frac)
    thr32, prob32 = float(np.float32(thr)), float(np.float32(prob))   # the reference's parameters are floats
    r = np_oracle.plane_ransac(x, thr32, prob32, max_it, 12345, order)
    mask = np_oracle.plane_inlier_mask(x, r["coeff"], thr32, order) if r["found"] else np.zeros(n, bool)
    cases.append({"xyzi": [[r9(v) for v in row] for row in x], "threshold": r9(thr32), "probability": r9(prob32),
                  "max_iterations": max_it, "seed": 12345, "sum_order": order,
                  "expected": {"found": bool(r["found"]), "iterations": int(r["iterations"]), "draws": int(r["draws"]),
                               "best_count": int(r["best_count"]), "sample": [int(v) for v in r["sample"]],
                               "coeff_ransac_bits": [int(v) for v in np.asarray(r["coeff"], np.float32).view(np.uint32)],
                               "inliers": [int(i) for i in np.nonzero(mask)[0]]}})
raw = np.random.RandomState(12345)._bit_generator.random_raw(6)
doc = {"source": "oracle/np_oracle.py (numpy restatement of PCL 1.8.1 SampleConsensusModelPlane + RandomSampleConsensus), optimize off",
       "mt19937_seed_12345_first_outputs": [int(v) for v in raw], "cases": cases}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "plane_ransac.json"), "w") as f:
    json.dump(doc, f)
print("wrote plane_ransac.json", [c["expected"]["iterations"] for c in cases])

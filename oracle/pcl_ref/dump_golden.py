#!/usr/bin/env python
"""Writes golden fixtures FROM THE REAL PCL (oracle/_ref/libcm_pcl_ref.so) into tests/golden/pcl_golden.npz. Run it on a
machine that has PCL 1.8.x (see README.md) and commit the file: tests/test_oracle.py::test_pcl_golden_fixtures then pins
the CPU oracle -- and tests/test_gpu_parity.py::test_pcl_golden_fixtures_gpu the CUDA path -- against PCL's own outputs
everywhere, including this repository's PCL-less build image and GPU box. Inputs are the seeded synthetic clouds of
cloud_merger_b200/synth.py, so the file holds outputs only (plus the seeds)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cloud_merger_b200 import synth  # noqa: E402
from oracle.pcl_ref import pcl_ref_py as pcl  # noqa: E402

SEED = 4242
LEAVES = [0.1, 0.05, 0.5]


def inputs():
    """(clouds, mats) of one cfg1-shaped frame with 0.5 % non-finite points: what every golden entry is computed from."""
    clouds = [synth.lidar_cloud(SEED, s, 0, 32, 512, nan_frac=0.005) for s in range(2)]
    mats = [synth.extrinsic(s, 2) for s in range(2)]
    return clouds, mats


def main():
    if not pcl.available():
        raise SystemExit("no PCL reference library: " + pcl.why_not())
    clouds, mats = inputs()
    out = {"pcl_version": np.array(pcl.version()), "seed": np.array(SEED)}
    merged = None
    for s, (c, m) in enumerate(zip(clouds, mats)):
        t = pcl.transform(c, m[:3].reshape(-1), is_dense=False)
        out["transform_%d" % s] = t
        cur = t
        for k, (axis, lo, hi, neg) in enumerate(synth.ROI_BOX):
            keep = pcl.passthrough(cur, axis, lo, hi, bool(neg))
            out["roi_%d_pass%d_idx" % (s, k)] = keep
            cur = np.ascontiguousarray(cur[keep])
        merged = cur if merged is None else pcl.concat(merged, True, s, cur, True, s + 1)[0]
    out["merged"] = merged
    for leaf in LEAVES:
        for mp in (1, 2):
            v, grid = pcl.voxelgrid(merged, leaf, mp)
            out["voxel_%g_%d" % (leaf, mp)] = v
            out["grid_%g_%d" % (leaf, mp)] = grid
    out["outlier_0.15_1_idx"] = pcl.radius_outlier(merged, 0.15, 1)
    low = np.ascontiguousarray(merged[pcl.passthrough(merged, 2, -0.5, 0.5)])
    coeff, inl = pcl.plane_ransac(low, 0.3, 0.99, 1000, True)
    out["plane_input_n"] = np.array(len(low))
    out["plane_coeff"] = coeff
    out["plane_inliers"] = inl
    q = np.array([0.02, -0.01, 0.3826834323650898, 0.9238795325112867]); q /= np.linalg.norm(q)
    out["tf_q"] = q
    out["tf_t"] = np.array([1.2, -0.7, 1.9])
    out["tf_matrix"] = pcl.tf_to_matrix(q, out["tf_t"])
    path = os.path.join(ROOT, "tests", "golden", "pcl_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "from", pcl.version())


if __name__ == "__main__":
    main()

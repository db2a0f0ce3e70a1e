"""Writes tests/golden/known_answer.json: the hand-checkable vector of SURVEY.md section 8c.

The INPUTS and EXPECTED values below were worked out by hand from the PCL 1.8.1 formulas (they are not produced by any
code in this repository); the file pins both oracle restatements and the GPU path. The reference itself ships no golden
vectors (parity unpinned). Run: python tests/golden/make_known_answer.py
"""
import json
import os

import numpy as np

f = np.float32


def r9(v):
    """9 significant digits: enough to round-trip a float32 exactly."""
    return float("%.9g" % float(v))


# The few values that are not exact decimals, derived step by step in float32 (each line is one rounded operation):
yB0 = f(f(1) * f(-1.97)) + f(2)            # row (1,0,0,2) of B on point (-1.97, 1.04, 0.52): x + 2
yB3 = f(f(1) * f(-1.96)) + f(2)            # same row on point (-1.96, 1.01, 0.55)
zA2 = f(-1.2) + f(0.5)                     # row (0,0,1,0.5) of A on z = -1.2 (cropped away; documents -0.70000005)
cx0 = (f(-1.04) + f(-1.01)) / f(2)         # centroid of voxel 0, x
cy0 = (yB0 + yB3) / f(2)
cz0 = (f(0.52) + f(0.55)) / f(2)
cx21 = (f(f(0.01) + f(1)) + f(f(0.04) + f(1))) / f(2)
cy21 = (f(0.02) + f(0.03)) / f(2)
cz21 = (f(f(0) + f(0.5)) + f(f(0.01) + f(0.5))) / f(2)
nan = float("nan")
doc = {
    "source": "SURVEY.md section 8c, hand-computed with PCL 1.8.1 formulas in float32",
    "extrinsics_row_major_3x4": {
        "A": [1, 0, 0, 1, 0, 1, 0, 0, 0, 0, 1, 0.5],
        "B": [0, -1, 0, 0, 1, 0, 0, 2, 0, 0, 1, 0],
    },
    "sensor_A_xyzi": [[0.01, 0.02, 0, 10], [0.04, 0.03, 0.01, 20], [0.26, 0, -1.2, 30], [5, 5, 2.5, 40]],
    "sensor_A_is_dense": 1,
    "sensor_B_xyzi": [[-1.97, 1.04, 0.52, 50], [0.5, 0.5, 3.5, 60], ["nan", 0, 0, 70], [-1.96, 1.01, 0.55, 80]],
    "sensor_B_is_dense": 0,
    "crop_passes": [[2, -0.5, 3.0, 0]],
    "leaf": [0.1, 0.1, 0.1],
    "cropped_away_z_of_A2": r9(zA2),
    "expected": {
        "survivor_src": [0, 1, 3, 4, 7],
        "survivor_xyz": [[1.01, 0.02, 0.5], [1.04, 0.03, 0.51], [6, 5, 3], [-1.04, r9(yB0), 0.52],
                         [-1.01, r9(yB3), 0.55]],
        "min_b": [-11, 0, 5], "max_b": [60, 50, 30], "div_b": [72, 51, 26],
        "point_idx": [21, 21, 95471, 0, 0],
        "min_points_1": {"idx": [0, 21, 95471], "count": [2, 2, 1],
                         "centroid": [[r9(cx0), r9(cy0), r9(cz0), 65], [r9(cx21), r9(cy21), r9(cz21), 15], [6, 5, 3, 40]]},
        "min_points_2": {"idx": [0, 21], "count": [2, 2],
                         "centroid": [[r9(cx0), r9(cy0), r9(cz0), 65], [r9(cx21), r9(cy21), r9(cz21), 15]]},
    },
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "known_answer.json"), "w") as f:
    json.dump(doc, f, indent=1)
print("wrote known_answer.json")

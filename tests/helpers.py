"""Shared helpers of the test-suite: input construction and comparisons (tests only)."""
from __future__ import annotations

import json
import os

import numpy as np

from cloud_merger_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


def known_answer():
    with open(os.path.join(HERE, "golden", "known_answer.json")) as f:
        d = json.load(f)

    def arr(rows):
        return np.array([[float("nan") if v == "nan" else float(v) for v in r] for r in rows], np.float32)
    d["A"] = arr(d["sensor_A_xyzi"])
    d["B"] = arr(d["sensor_B_xyzi"])
    return d


def cloud_dict(xyzi, m, is_dense=1, point_step=16, off_x=0, off_y=4, off_z=8, off_i=12):
    """Oracle-side description of one sensor cloud (bytes in the given PointCloud2 layout)."""
    data = synth.pack_cloud(xyzi, point_step, off_x, off_y, off_z, off_i)
    return dict(data=data, n_points=len(xyzi), point_step=point_step, off_x=off_x, off_y=off_y, off_z=off_z,
                off_i=off_i, is_dense=int(is_dense), m=np.asarray(m, np.float32).reshape(-1)[:12].copy())


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
    if not same.all():
        bad = np.argwhere(~same)
        raise AssertionError("%s: %d of %d values differ bitwise, first at %s: %r vs %r" % (
            what, len(bad), a.size, bad[0], a[tuple(bad[0])], b[tuple(bad[0])]))


def rel_err(g, o):
    g = np.asarray(g, np.float64)
    o = np.asarray(o, np.float64)
    return np.abs(g - o) / np.maximum(np.abs(o), 1e-30)


CENTROID_RTOL = 1e-5  # north_star: centroids within 1e-5 relative tolerance


def assert_centroids_close(g, o_f64, what=""):
    """|g - o| <= 1e-5 * max(|o|, tiny) per component, with an absolute floor of 1e-5 * leaf-scale for values that
    cancel to ~0 (a centroid coordinate near 0 has no meaningful relative error)."""
    g = np.asarray(g, np.float64)
    o = np.asarray(o_f64, np.float64)
    assert g.shape == o.shape, "%s: shape %s vs %s" % (what, g.shape, o.shape)
    err = np.abs(g - o)
    tol = CENTROID_RTOL * np.maximum(np.abs(o), 1e-2)
    if not (err <= tol).all():
        i = np.argmax(err - tol)
        raise AssertionError("%s: centroid off by %g (tol %g) at flat index %d" % (what, err.flat[i], tol.flat[i], i))
    return float((err / np.maximum(np.abs(o), 1e-2)).max()) if err.size else 0.0


def reference_front_zones():
    """The ten PassThrough chains proceedFront runs on one ROI-cropped cloud (pc_preprocessing_main.cpp:228-270 with
    pcl_preprocessing/src/Parameter.h:31-55): five x windows [deviation, deviation + length] (float arithmetic), each
    followed by the ground window z in [-zg, zg] and the no-ground window z in [zg + 0.01, roi_z_max] -- the `+ 0.01` is
    a double addition narrowed to float by setFilterLimits (removeGround, :80-92)."""
    f32 = np.float32
    roi_mid, roi_z_max = f32(15), f32(3.0)
    front, mid, mid2, veh, rear = f32(30.0), f32(15.0), f32(11.0), f32(8.0), f32(11)
    parts = [  # (length, deviation, z_max_ground)
        (front, -roi_mid + rear + veh + mid + mid2, f32(2.5)),
        (mid2, -roi_mid + rear + veh + mid, f32(2.0)),
        (mid, -roi_mid + rear + veh, f32(1.5)),
        (veh, -roi_mid + rear, f32(0.3)),
        (rear, -roi_mid, f32(0.5)),
    ]
    zones = []
    for length, dev, zg in parts:
        x_pass = (0, float(f32(dev)), float(f32(dev) + f32(length)), 0)
        zones.append([x_pass, (2, float(-zg), float(zg), 0)])
        zones.append([x_pass, (2, float(f32(np.float64(zg) + 0.01)), float(roi_z_max), 0)])
    return zones

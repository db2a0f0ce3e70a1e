"""The whole body of a proceedX of the reference after getROI (pc_preprocessing_main.cpp:228-312: five x windows, per window
the two z windows, the RANSAC plane + ExtractIndices, outlierRemoval, appended) on one sensor ROI cloud: GPU (one
zone-slicing pass, one multi-cloud RANSAC pass, one multi-cloud radius outlier removal; device-resident between the stages) next
to the CPU restatement (oracle, one core) of the same sequence. One JSON line; wall clock, results compared."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cloud_merger_b200 import ROI_PASSES, CloudMerger, synth
from oracle import cm_oracle_py as oracle  # CPU side of the comparison (bench-side use only)

THR, PROB, RADIUS = float(np.float32(0.3)), float(np.float32(0.99)), float(np.float32(0.15))
ROI_Z_MAX = 3.0
# {length, deviation, z_max_ground} of proceedFront (Parameter.h:45-55 with roi_mid = 15)
PARTS = [(30.0, 30.0, 2.5), (11.0, 19.0, 2.0), (15.0, 4.0, 1.5), (8.0, -4.0, 0.3), (11.0, -15.0, 0.5)]


def roi_cloud(seed, rings, az):
    c = synth.lidar_cloud(seed, 0, 0, rings, az)
    m = synth.extrinsic(0, 4)[:3]
    cur = oracle.transform(c, m.reshape(-1))
    for (axis, lo, hi, neg) in ROI_PASSES:
        cur = np.ascontiguousarray(cur[oracle.passthrough(cur, axis, lo, hi, bool(neg))])
    return cur


def zones_of(parts):
    z = []
    for (length, dev, zg) in parts:
        x = (0, dev, dev + length, 0)
        z.append([x, (2, -zg, zg, 0)])
    for (length, dev, zg) in parts:
        x = (0, dev, dev + length, 0)
        z.append([x, (2, float(np.float32(np.float64(np.float32(zg)) + 0.01)), ROI_Z_MAX, 0)])
    return z


def cpu_chain(cloud, zones):
    k = len(PARTS)
    parts = oracle.zone_split(cloud, zones)
    no_ground, ground = [], []
    for i in range(k):
        low, high = parts[i][0], parts[k + i][0]
        r = oracle.plane_ransac(low, THR, PROB)
        inl = r["inliers"]
        rest = np.ascontiguousarray(np.delete(low, inl, axis=0))
        keep = oracle.radius_outlier(rest, RADIUS, 1)
        no_ground += [rest[keep], high]
        ground.append(low[inl])
    return np.concatenate(no_ground), np.concatenate(ground)


def gpu_chain(cms, buf, n, scratch):
    """cms = (zones, plane, ror) handles: each stage writes its results into its own handle's zone outputs."""
    cz, cp, cr = cms
    k = len(PARTS)
    cz.dev_zone_split(buf.ptr, n)
    z_xyzi, _, z_begin = cz.zone_out_raw()   # the 2k zones on the device: k ground windows first, then the k upper ones
    begin = np.array(z_begin, np.int64)
    res = cp.dev_plane_ransac_multi(z_xyzi, begin[:k + 1], THR, PROB)
    p_xyzi, _, p_begin = cp.zone_out_raw()
    pb = np.array(p_begin, np.int64)
    # what is not ground, zone by zone: the rest clouds are zones 1, 3, 5, ... of the plane handle; gather them side by
    # side in the scratch buffer and filter them in one multi-cloud pass
    rb = [0]
    for i in range(k):
        n_rest = int(pb[2 * i + 2] - pb[2 * i + 1])
        cr.memcpy_d2d(scratch.ptr + rb[-1] * 16, p_xyzi + int(pb[2 * i + 1]) * 16, n_rest * 16)
        rb.append(rb[-1] + n_rest)
    cr.dev_radius_outlier_multi(scratch.ptr, rb, RADIUS, 1)
    kept = cr.zone_out()
    pts_plane = cp.zone_out()
    pts_zone = cz.zone_out()
    out_ng, ground = [], []
    for i in range(k):
        out_ng += [kept[i][0], pts_zone[k + i][0]]
        ground.append(pts_plane[2 * i][0])
    return np.concatenate(out_ng), np.concatenate(ground), res


for rings, az in ((64, 2048), (128, 4096)):
    cloud = roi_cloud(7, rings, az)
    n = len(cloud)
    zones = zones_of(PARTS)
    cms = tuple(CloudMerger(max_sensors=1, max_points_per_sensor=n, max_batch_points=n, max_batch_frames=8) for _ in range(3))
    cms[0].set_zones(zones)
    buf = cms[0].upload(cloud)
    scratch = cms[2].upload(np.zeros((n, 4), np.float32))
    for _ in range(3):
        g_ng, g_g, res = gpu_chain(cms, buf, n, scratch)
    torch.cuda.synchronize()
    steps = 10
    t0 = time.perf_counter()
    for _ in range(steps):
        g_ng, g_g, res = gpu_chain(cms, buf, n, scratch)
    gpu_ms = (time.perf_counter() - t0) * 1e3 / steps
    t0 = time.perf_counter()
    c_ng, c_g = cpu_chain(cloud, zones)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    same = g_ng.tobytes() == c_ng.tobytes() and g_g.tobytes() == c_g.tobytes()
    print(json.dumps({"op": "proceedX after getROI", "roi_points": n, "ground": len(g_g), "no_ground": len(g_ng),
                      "ransac_iterations": [r["iterations"] for r in res], "gpu_ms": round(gpu_ms, 3),
                      "cpu_port_ms": round(cpu_ms, 2), "bit_identical_to_cpu": bool(same),
                      "note": "GPU time includes the D2H of both result clouds and the python glue between the stages"}))
    for c in cms:
        c.close()

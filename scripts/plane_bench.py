"""RANSAC ground plane on one GPU (the pcl::SACSegmentation + ExtractIndices block of removeGround(), reference
pc_preprocessing_main.cpp:95-117: 1000 iterations max, threshold 0.3 m, probability 0.99, optimize on) on synthetic ground
zones already in device memory, next to the CPU restatement of PCL's loop (oracle, one core) on the same cloud.
One JSON line per zone size. The call blocks (the host applies PCL's stopping rule), so the time is wall clock."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cloud_merger_b200 import CloudMerger
from oracle import cm_oracle_py as oracle  # the CPU baseline beside the GPU number (bench-side use only)


def ground_scene(seed, n, ground_frac=0.7):
    rng = np.random.default_rng(seed)
    g = int(n * ground_frac)
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = rng.uniform(-30, 30, n)
    pts[:, 1] = rng.uniform(-10, 10, n)
    pts[:g, 2] = -1.8 + 0.02 * pts[:g, 0] + rng.normal(0, 0.03, g)
    pts[g:, 2] = rng.uniform(-1.5, 1.0, n - g)
    return pts[rng.permutation(n)]


THR, PROB = float(np.float32(0.3)), float(np.float32(0.99))
for n, frac, prob, label in ((20000, 0.7, PROB, "zone"), (200000, 0.7, PROB, "whole ROI"), (200000, 0.15, 0.999999, "weak plane: runs to the cap")):
    cloud = ground_scene(1, n, frac)
    cm = CloudMerger(max_sensors=1, max_points_per_sensor=n, max_batch_points=n)
    buf = cm.upload(cloud)
    for _ in range(3):
        r = cm.dev_plane_ransac(buf.ptr, n, THR, prob)
    torch.cuda.synchronize()
    steps = 20
    t0 = time.perf_counter()
    for _ in range(steps):
        r = cm.dev_plane_ransac(buf.ptr, n, THR, prob)
    ms = (time.perf_counter() - t0) * 1e3 / steps
    t0 = time.perf_counter()
    w = oracle.plane_ransac(cloud, THR, prob)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    same = (r["iterations"], r["best_count"], r["n_inliers"]) == (w["iterations"], w["best_count"], len(w["inliers"]))
    print(json.dumps({"op": "plane_ransac", "case": label, "points": n, "iterations": r["iterations"], "inliers": r["n_inliers"],
                      "gpu_ms": round(ms, 4), "cpu_port_ms": round(cpu_ms, 3), "same_result_as_cpu": bool(same)}))
    cm.close()

# five ground zones of one sensor cloud in one call (proceedX runs five removeGround in a row) vs five single calls
sizes = (30000, 11000, 15000, 8000, 20000)
clouds = [ground_scene(10 + i, n, 0.75) for i, n in enumerate(sizes)]
begin = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
allpts = np.ascontiguousarray(np.concatenate(clouds))
cm = CloudMerger(max_sensors=1, max_points_per_sensor=len(allpts), max_batch_points=len(allpts))
buf = cm.upload(allpts)
for _ in range(3):
    cm.dev_plane_ransac_multi(buf.ptr, begin, THR, PROB)
steps = 20
t0 = time.perf_counter()
for _ in range(steps):
    res = cm.dev_plane_ransac_multi(buf.ptr, begin, THR, PROB)
multi_ms = (time.perf_counter() - t0) * 1e3 / steps
t0 = time.perf_counter()
for _ in range(steps):
    for k in range(len(sizes)):
        cm.dev_plane_ransac(buf.ptr + int(begin[k]) * 16, sizes[k], THR, PROB)
single_ms = (time.perf_counter() - t0) * 1e3 / steps
t0 = time.perf_counter()
want = [oracle.plane_ransac(c, THR, PROB) for c in clouds]
cpu_ms = (time.perf_counter() - t0) * 1e3
same = all(r["iterations"] == w["iterations"] and r["n_inliers"] == len(w["inliers"]) for r, w in zip(res, want))
print(json.dumps({"op": "plane_ransac_multi", "case": "5 ground zones of one sensor cloud", "points": int(begin[-1]),
                  "iterations": [r["iterations"] for r in res], "gpu_ms_one_call": round(multi_ms, 4),
                  "gpu_ms_five_calls": round(single_ms, 4), "cpu_port_ms": round(cpu_ms, 3), "same_result_as_cpu": bool(same)}))
cm.close()

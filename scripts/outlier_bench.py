"""Radius outlier removal throughput on one GPU (outlierRemoval() of the reference: radius 0.15 m, min_neighbor 1) on
lidar-shaped ROI clouds already in device memory. One JSON line per size: Mpoints/s and the CUDA-event time of the whole
operation (bounding box, cell keys, radix sort, neighbour count, compaction)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cloud_merger_b200 import ROI_PASSES, CloudMerger, synth

r = float(np.float32(0.15))
for frames in (1, 16):
    S, rings, az = 4, 128, 1024
    clouds = []
    for f in range(frames):
        for s in range(S):
            c = synth.lidar_cloud(2000, s, f, rings, az)
            m = synth.extrinsic(s, S)[:3]
            xyz = (c[:, :3].astype(np.float64) @ m[:, :3].T.astype(np.float64) + m[:, 3]).astype(np.float32)
            keep = np.ones(len(c), bool)
            for axis, lo, hi, _ in ROI_PASSES:
                keep &= (xyz[:, axis] >= lo) & (xyz[:, axis] <= hi)
            clouds.append(np.column_stack([xyz[keep], c[keep, 3]]).astype(np.float32))
    cloud = np.ascontiguousarray(np.concatenate(clouds))
    n = len(cloud)
    cm = CloudMerger(max_sensors=1, max_points_per_sensor=n, max_batch_points=n)
    buf = cm.upload(cloud)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        cm.dev_radius_outlier(buf.ptr, n, r, 1, stream=stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        cm.dev_radius_outlier(buf.ptr, n, r, 1, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    kept = len(cm.radius_outlier_out()[1])
    print(json.dumps({"op": "radius_outlier", "radius_m": 0.15, "min_neighbors": 1, "points": n, "kept": kept,
                      "ms": round(ms, 4), "mpoints_per_s": round(n / ms / 1e3, 1),
                      "note": "includes one host round trip for the key width (unbounded grid) per call"}))
    cm.close()

"""Zone slicing throughput on one GPU: the ten chains of proceedFront over a ROI-shaped cloud already in device memory.
Prints one JSON line: Mpoints/s, the three launches' CUDA-event time and algorithmic GB/s against the measured HBM peak."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from cloud_merger_b200 import CloudMerger
from helpers import reference_front_zones

from cloud_merger_b200 import ROI_PASSES, synth

mode = sys.argv[1] if len(sys.argv) > 1 else "lidar"
steps = 10
if mode == "uniform":  # every zone present in every warp: the worst case for the per-zone ballots
    n = 1 << 24
    rng = np.random.default_rng(5)
    cloud = np.column_stack([rng.uniform(-15, 60, n), rng.uniform(-5, 5, n), rng.uniform(-0.5, 3, n),
                             rng.uniform(0, 255, n)]).astype(np.float32)
else:  # 32 frames x 4 sensors of the cfg2 lidar pattern, transformed and ROI-cropped: what getROI hands to getCloudPart
    S, rings, az = 4, 128, 1024
    parts = []
    for f in range(32):
        for s in range(S):
            c = synth.lidar_cloud(2000, s, f, rings, az)
            m = synth.extrinsic(s, S)[:3]
            xyz = (c[:, :3].astype(np.float64) @ m[:, :3].T.astype(np.float64) + m[:, 3]).astype(np.float32)
            keep = np.ones(len(c), bool)
            for axis, lo, hi, _ in ROI_PASSES:
                keep &= (xyz[:, axis] >= lo) & (xyz[:, axis] <= hi)
            parts.append(np.column_stack([xyz[keep], c[keep, 3]]).astype(np.float32))
    cloud = np.ascontiguousarray(np.concatenate(parts))
    n = len(cloud)
zones = reference_front_zones()
cm = CloudMerger(max_sensors=1, max_points_per_sensor=n)
cm.set_zones(zones)
buf = cm.upload(cloud)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    cm.dev_zone_split(buf.ptr, n, stream=stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    cm.dev_zone_split(buf.ptr, n, stream=stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
out = cm.zone_out()
total = sum(len(s) for _, s in out)
peak = 6534.1
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
algo = n * (16 + 2 + 16 + 2) + total * 20
print(json.dumps({"op": "zone_split", "input": mode, "zones": len(zones), "points": n, "points_out": total, "ms": round(ms, 4),
                  "mpoints_per_s": round(n / ms / 1e3, 1), "algorithmic_bytes": algo,
                  "achieved_gbs": round(algo / ms / 1e6, 1), "frac_of_measured_hbm_peak": round(algo / ms / 1e6 / peak, 4),
                  "cpu_equivalent": "15 pcl::PassThrough runs + copies per cloud (getCloudPart x5, two z windows each)"}))
cm.close()

"""One main-loop iteration of the reference's built node (pcl_preprocessing: per sensor transform -> getROI -> five x
windows with z split, RANSAC plane, ExtractIndices, outlierRemoval -> fuse -> VoxelGrid) device-resident on one GPU
(cloud_merger_b200/node.py) next to the CPU restatement of the same sequence (oracle, one core). One JSON line per size;
wall clock including the upload of the raw clouds and the download of the three published clouds."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from cloud_merger_b200 import synth
from cloud_merger_b200.node import NodeParams, PreprocessingNode
from oracle import cm_oracle_py as oracle  # CPU side of the comparison (bench-side use only)
from test_gpu_node import oracle_frame

for S, rings, az in ((4, 64, 1024), (4, 128, 2048)):
    p = NodeParams()
    mats = [synth.extrinsic(s, S) for s in range(S)]
    clouds = [synth.lidar_cloud(6000, s, 0, rings, az) for s in range(S)]
    ms = {}
    for concurrent in (False, True):   # sensors one after the other / one host thread + CUDA stream per sensor
        node = PreprocessingNode(S, rings * az, p, concurrent=concurrent)
        for s in range(S):
            node.set_extrinsic(s, mats[s])
        for _ in range(3):
            got = node.frame(clouds)
        steps = 10
        t0 = time.perf_counter()
        for _ in range(steps):
            got = node.frame(clouds)
        ms[concurrent] = (time.perf_counter() - t0) * 1e3 / steps
        if not concurrent:
            node.close()
    gpu_ms = ms[True]
    t0 = time.perf_counter()
    ng, g, vg = oracle_frame(oracle, clouds, mats, p)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    same = got["no_ground"].tobytes() == ng.tobytes() and got["ground"].tobytes() == g.tobytes() and \
        got["voxel"].tobytes() == vg["centroid"].tobytes()
    print(json.dumps({"op": "pcl_preprocessing main-loop iteration", "sensors": S, "points_in": S * rings * az,
                      "roi_points": got["roi_points"], "no_ground": got["n_no_ground"], "ground": got["n_ground"],
                      "voxels": got["n_voxels"], "gpu_ms": round(gpu_ms, 3), "gpu_ms_sensors_in_sequence": round(ms[False], 3),
                      "cpu_port_ms": round(cpu_ms, 1),
                      "bit_identical_to_cpu": bool(same)}))
    node.close()

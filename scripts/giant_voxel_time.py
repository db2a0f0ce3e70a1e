#!/usr/bin/env python
"""How long the VoxelGrid takes when ONE voxel holds most of the cloud (the centroid pass finishes such a run with the whole
CTA: finish_tail_run in cm_voxel.cu). usage: python scripts/giant_voxel_time.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cloud_merger_b200 import CloudMerger, synth  # noqa: E402

rng = np.random.default_rng(31)
for dense in (700001, 4000000):
    x = synth.uniform_cloud(91, 300000, extent=(60.0, 60.0, 6.0))
    blob = np.zeros((dense, 4), np.float32)
    blob[:, :3] = rng.uniform(0.01, 0.09, size=(dense, 3)).astype(np.float32)
    x = np.concatenate([x, blob])[rng.permutation(len(x) + dense)]
    with CloudMerger(max_batch_points=len(x)) as cm:
        cm.set_voxel(0.1, 2, True)
        cm.set_profiling(True)
        buf = cm.upload(x)
        for _ in range(3):
            cm.dev_voxelgrid(buf.ptr, len(x))
            st = cm.stats()
        print("one voxel of %d points in a cloud of %d: VoxelGrid %.3f ms, centroid stage %.3f ms"
              % (dense, len(x), st.gpu_ms, cm.stage_ms("centroid")))

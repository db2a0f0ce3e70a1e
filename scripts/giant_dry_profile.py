#!/usr/bin/env python
"""One rank's device-side work of cm_giant_voxelgrid up to the exchange, without a communicator ("dry" object: world virtual
ranks, the collectives are identities): bounds, plan, histogram, splitters, mask, count, scan, grouping. For a kernel launch
list of the partition plan on ONE GPU (ncu never runs under a multi-rank launch):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/giant_dry.csv \
      python scripts/giant_dry_profile.py --points 12500000 --world 8
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cloud_merger_b200 import CloudMerger, GiantCloud, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=12_500_000)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--leaf", type=float, default=0.02)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    import torch
    block = torch.from_numpy(np.ascontiguousarray(synth.map_cloud(4, a.points))).cuda()
    with CloudMerger(max_batch_points=a.points) as cm:
        cm.set_voxel(a.leaf, 1, True)
        g = GiantCloud(cm, 1, a.world, None)
        for _ in range(a.iters):
            info = g.voxelgrid(block.data_ptr(), a.points, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
        g.close()
    print({k: info[k] for k in ("points_sent_away", "key_bits", "send_begin")})


if __name__ == "__main__":
    main()

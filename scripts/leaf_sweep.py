#!/usr/bin/env python
"""BASELINE config 5: VoxelGrid leaf-size sweep 0.01-1.0 m on 16 Mi uniform points (200 x 200 x 10 m): dense-to-sparse
occupancy, 32- and 64-bit keys, 3-5 radix passes. One JSON line per leaf. --check compares with the oracle (slow)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloud_merger_b200 import CloudMerger, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1 << 24)
    ap.add_argument("--leaves", default="0.01,0.02,0.05,0.1,0.2,0.5,1.0")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    peak = 6534.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    x = synth.uniform_cloud(5000, a.points)
    with CloudMerger(max_batch_points=a.points) as cm:
        buf = cm.upload(x)
        cm.set_profiling(True)
        for leaf in [float(v) for v in a.leaves.split(",")]:
            cm.set_voxel(leaf, 1, True)
            ms = []
            for it in range(a.iters + 2):
                cm.dev_voxelgrid(buf.ptr, a.points)
                st = cm.stats()
                if it >= 2:
                    ms.append(st.gpu_ms)
            t = float(np.median(ms))
            M, V, P, k = int(st.survivors), int(st.voxels_out), int(st.sort_passes), int(st.key_bytes)
            b_vg = M * (16 + 2 * k + 2 * P * (k + 4) + (k + 4) + 16) + 20 * V
            b_io = 16 * M + 20 * V
            line = {"workload": "cfg5 leaf sweep", "points": M, "leaf_m": leaf, "voxels": V, "V_over_M": round(V / M, 4),
                    "key_bytes": k, "key_bits": int(st.key_bits), "passes": P, "ms": round(t, 4),
                    "mpoints_per_s": round(M / t / 1e3, 1), "B_vg_frac": round(b_vg / (t * 1e-3) / 1e9 / peak, 4),
                    "B_io_frac": round(b_io / (t * 1e-3) / 1e9 / peak, 4),
                    "pcl_domain": "64-bit oracle only (PCL 1.8.1 refuses: dx*dy*dz > INT32_MAX)" if st.pcl_overflow else "inside PCL's int32 domain"}
            if a.check:
                from oracle import cm_oracle_py as oracle
                o = oracle.voxelgrid(x, [leaf] * 3, 1, True, force64=True)
                out = cm.fetch_batch_outputs(want_sorted=False)
                ok = (out["voxel_idx"].astype(np.int64) == o["idx"]).all() and (out["voxel_count"] == o["count"]).all() and \
                     (out["voxel_xyzi"].view(np.uint32) == o["centroid"].view(np.uint32)).all()
                line["check"] = "ok" if ok else "MISMATCH"
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

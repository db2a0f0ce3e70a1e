"""transform_crop (K1) throughput per PointCloud2 record layout on one GPU: the same cfg2 batch packed as 16-byte xyzi,
32-byte pcl::PointXYZI / Velodyne-Melodic, 22-byte Velodyne and 18-byte Livox records, device-resident.
One JSON line per layout: CUDA-event time of the stage, algorithmic GB/s (n * point_step + 20 * survivors) vs the measured peak."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from cloud_merger_b200 import (LAYOUT_LIVOX18, LAYOUT_PACKED16, LAYOUT_PCL32, LAYOUT_VELODYNE22, CloudMerger, make_layout,
                               synth)

F = int(sys.argv[1]) if len(sys.argv) > 1 else 32
c = synth.CONFIGS["cfg2"]
S, n = c["sensors"], c["rings"] * c["azimuth"]
peak = 6534.1
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for name, L in (("packed16", LAYOUT_PACKED16), ("pcl32", LAYOUT_PCL32), ("velodyne22", LAYOUT_VELODYNE22), ("livox18", LAYOUT_LIVOX18)):
    step = L["point_step"]
    cm = CloudMerger(max_sensors=S, max_batch_points=F * S * n, max_batch_frames=F, max_point_step=step)
    for s in range(S):
        cm.set_extrinsic(s, synth.extrinsic(s, S))
    cm.set_crop(c["passes"])
    cm.set_profiling(True)
    layout = make_layout(step, L["off_x"], L["off_y"], L["off_z"], L["off_intensity"], 1)
    items = []
    for f in range(F):
        for s in range(S):
            raw = synth.pack_cloud(synth.lidar_cloud(2000, s, f, c["rings"], c["azimuth"]), step, L["off_x"], L["off_y"], L["off_z"], L["off_intensity"])
            items.append((cm.upload(raw).ptr, n, layout, s, f))
    segs = cm.make_segments(items)
    ms = []
    for i in range(8):
        cm.dev_transform_crop(segs)
        cm.sync()
        if i >= 3:
            ms.append(cm.stage_ms("transform_crop"))
    st = cm.stats()
    t = float(np.median(ms))
    algo = F * S * n * step + 20 * int(st.survivors)
    print(json.dumps({"op": "transform_crop", "layout": name, "point_step": step, "points": F * S * n, "survivors": int(st.survivors),
                      "ms": round(t, 4), "gpoints_per_s": round(F * S * n / t / 1e6, 2), "algorithmic_bytes": algo,
                      "achieved_gbs": round(algo / t / 1e6, 1), "frac_of_measured_hbm_peak": round(algo / t / 1e6 / peak, 4),
                      "note": "stage time includes the one-CTA tile scan (K1-only entry)"}))
    cm.close()

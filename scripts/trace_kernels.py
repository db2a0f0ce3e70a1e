"""Debug: per-phase clock64 breakdown of the transform_crop kernel and one radix pass (CM_TRACE=1)."""
import ctypes as C
import os
import sys

os.environ["CM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from cloud_merger_b200 import CloudMerger, make_layout, synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
c = synth.CONFIGS["cfg2"]
S, n = c["sensors"], c["rings"] * c["azimuth"]
cm = CloudMerger(max_sensors=S, max_batch_points=F * S * n, max_batch_frames=F)
for s in range(S):
    cm.set_extrinsic(s, synth.extrinsic(s, S))
cm.set_crop(c["passes"]); cm.set_voxel(c["leaf"], c["min_points"], True)
items = []
for f in range(F):
    for s in range(S):
        b = cm.upload(synth.lidar_cloud(2000, s, f, c["rings"], c["azimuth"]))
        items.append((b.ptr, n, make_layout(), s, f))
segs = cm.make_segments(items)
for _ in range(4):
    cm.run_batch(segs); cm.sync()
st = cm.stats()
print("survivors", st.survivors, "voxels", st.voxels_out, "gpu_ms", st.gpu_ms)
for which, name, labels in ((0, "transform_crop", ["desc+load", "xform+rank", "sync1", "rec+minmax", "(unused)", "stores"]),
                            (1, "onesweep pass", ["setup+load issue", "match+rank+sync", "counts published", "scans+sync", "placement", "row wait+sync", "scatter"])):
    nn = C.c_int64()
    cm._check(cm._lib.cm_debug_trace(cm._h, which, None, 0, C.byref(nn)))
    buf = np.zeros(nn.value, np.uint64)
    cm._check(cm._lib.cm_debug_trace(cm._h, which, buf.ctypes.data_as(C.c_void_p), nn.value, C.byref(nn)))
    t = buf.reshape(-1, 8).astype(np.int64)
    t = t[t[:, len(labels) - 1] > 0]
    print("%s: %d tiles traced" % (name, len(t)))
    if which == 1:
        # stamps 0..4 (front) and 5, 6, 7 (back) are relative to the iteration in which the tile's front ran; 7 = back begins
        t = t[t[:, 6] > 0]
        cols = [("setup+load issue", t[:, 0]), ("match+rank+sync", t[:, 1] - t[:, 0]), ("counts published", t[:, 2] - t[:, 1]),
                ("scans+sync", t[:, 3] - t[:, 2]), ("placement", t[:, 4] - t[:, 3]), ("(next tile's front)", t[:, 7] - t[:, 4]),
                ("row wait+sync", t[:, 5] - t[:, 7]), ("scatter", t[:, 6] - t[:, 5])]
        for lab, d in cols:
            print("   %-20s median %7d  p90 %7d  mean %7d cycles" % (lab, np.median(d), np.percentile(d, 90), d.mean()))
        continue
    prev = np.zeros(len(t), np.int64)
    for i, lab in enumerate(labels):
        if lab == "(unused)":
            continue
        d = t[:, i] - prev
        prev = t[:, i]
        print("   %-18s median %7d  p90 %7d  mean %7d cycles" % (lab, np.median(d), np.percentile(d, 90), d.mean()))
    print("   %-18s median %7d  p90 %7d" % ("TOTAL", np.median(t[:, len(labels) - 1]), np.percentile(t[:, len(labels) - 1], 90)))
cm.close()

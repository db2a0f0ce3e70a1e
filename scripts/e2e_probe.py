"""Where does the host path spend its time? Wall time per C-ABI call type over a stream of cfg2 frames (pinned inputs)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from cloud_merger_b200 import CloudMerger, host_alloc, make_layout, synth

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
F = 61
c = synth.CONFIGS["cfg2"]
S, n = c["sensors"], c["rings"] * c["azimuth"]
cm = CloudMerger(max_sensors=S, max_points_per_sensor=n, frames_in_flight=depth + 1)
for s in range(S):
    cm.set_extrinsic(s, synth.extrinsic(s, S))
cm.set_crop(c["passes"]); cm.set_voxel(c["leaf"], c["min_points"], True)
layout = make_layout()
arena, _addr = host_alloc(F * S * n * 16)
host = arena.view(np.float32).reshape(F, S, n, 4)
for f in range(F):
    for s in range(S):
        host[f, s] = synth.lidar_cloud(2000, s, f, c["rings"], c["azimuth"])
ring = [cm.make_frame_buffers(S * n, want_survivors=False, pinned=True) for _ in range(depth + 1)]
t_sub = t_mrg = t_wait = 0.0
gpu_ms = []
mode = sys.argv[2] if len(sys.argv) > 2 else "full"
submit = [cm.prepared_frame_submit(list(range(S)), [_addr + (f * S + s) * n * 16 for s in range(S)], [n] * S, [layout] * S, stamp=f) for f in range(F)]


def step():
    global t_sub, t_mrg, t_wait
    pending = []
    for f in range(F):
        t0 = time.perf_counter()
        if mode == "single":
            for s in range(S):
                cm.submit_cloud(s, host[f, s], n, layout, stamp=f, pinned=True)
        else:
            submit[f]()
        t1 = time.perf_counter()
        pending.append((cm.merge_frame_async(), ring[f % (depth + 1)][0]))
        t2 = time.perf_counter()
        if len(pending) >= depth:
            t, o = pending.pop(0)
            cm.wait_frame_into(t, o)
            gpu_ms.append(cm.stats().gpu_ms)
        t3 = time.perf_counter()
        t_sub += t1 - t0; t_mrg += t2 - t1; t_wait += t3 - t2
    while pending:
        t, o = pending.pop(0)
        cm.wait_frame_into(t, o)


step()
t_sub = t_mrg = t_wait = 0.0
t0 = time.perf_counter()
for _ in range(3):
    step()
dt = time.perf_counter() - t0
k = 3 * F
print("mode %s, gpu_ms/frame median %.3f" % (mode, float(np.median(gpu_ms))))
print("depth %d: %.1f us/frame total; submit x4 %.1f, merge_async %.1f, wait %.1f us; %.0f Mpts/s" % (
    depth, dt / k * 1e6, t_sub / k * 1e6, t_mrg / k * 1e6, t_wait / k * 1e6, S * n * k / dt / 1e6))
cm.close()

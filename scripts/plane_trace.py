import os, sys
os.environ["CM_PLANE_TRACE"]="1"
sys.path.insert(0, "/root/repo")
import numpy as np
from cloud_merger_b200 import CloudMerger
sys.path.insert(0, "/root/repo/scripts")
def ground_scene(seed, n, ground_frac=0.7):
    rng = np.random.default_rng(seed)
    g = int(n * ground_frac)
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = rng.uniform(-30, 30, n); pts[:, 1] = rng.uniform(-10, 10, n)
    pts[:g, 2] = -1.8 + 0.02 * pts[:g, 0] + rng.normal(0, 0.03, g); pts[g:, 2] = rng.uniform(-1.5, 1.0, n - g)
    return pts[rng.permutation(n)]
THR, PROB = float(np.float32(0.3)), float(np.float32(0.99))
c = ground_scene(1, 20000)
cm = CloudMerger(max_sensors=1, max_points_per_sensor=200000, max_batch_points=200000)
buf = cm.upload(c)
for _ in range(4): cm.dev_plane_ransac(buf.ptr, len(c), THR, PROB)
sizes = (30000, 11000, 15000, 8000, 20000)
clouds = [ground_scene(10 + i, n, 0.75) for i, n in enumerate(sizes)]
begin = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
buf2 = cm.upload(np.ascontiguousarray(np.concatenate(clouds)))
for _ in range(4): cm.dev_plane_ransac_multi(buf2.ptr, begin, THR, PROB)
big = ground_scene(1, 200000)
buf3 = cm.upload(big)
for _ in range(3): cm.dev_plane_ransac(buf3.ptr, len(big), THR, PROB)

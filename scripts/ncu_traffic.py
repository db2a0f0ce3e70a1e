"""Writes profiles/traffic.json from an `ncu --set full` report of bench.py: per kernel, the average
dram__bytes_read.sum + dram__bytes_write.sum per launch (what bench.py reports as roofline.traffic).
usage: python scripts/ncu_traffic.py <report.ncu-rep> <workload> <frames_per_step>"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

rep, workload, frames = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


acc = defaultdict(list)
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    for k in ("k_onesweep_pass", "k_transform_crop", "k_voxel_centroid", "k_voxel_key_hist", "k_compact_voxels"):
        if k in name:
            b = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
                to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
            t_us = float(r[ix["gpu__time_duration.sum"]].replace(",", "")) * {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1.0)
            acc[k].append((b, t_us))
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d["%s/%d" % (workload, frames)] = {
    k: {"bytes_per_launch": int(sum(b for b, _ in v) / len(v)), "launches_profiled": len(v),
        "avg_us_under_ncu": round(sum(t for _, t in v) / len(v), 2),
        "source": "ncu --set full --clock-control none, report %s" % os.path.basename(rep)} for k, v in acc.items()}
json.dump(d, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(d["%s/%d" % (workload, frames)], indent=1))

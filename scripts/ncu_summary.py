"""Summarise ncu outputs brought back from the GPU box into small text files under profiles/ (run here, no GPU needed).
usage: python scripts/ncu_summary.py <launches.csv> <report.ncu-rep> <out_prefix>"""
import csv
import subprocess
import sys
from collections import defaultdict


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = defaultdict(list)
    for r in rows:
        agg[r["Kernel Name"].split("(")[0][-48:]].append((float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
    total = sum(v[0] for vs in agg.values() for v in vs)
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("%-50s %6s %12s %8s  %s\n" % ("kernel", "n", "avg_us", "share", "grid x block (last)"))
        for k, vs in sorted(agg.items(), key=lambda kv: -sum(v[0] for v in kv[1])):
            s = sum(v[0] for v in vs)
            f.write("%-50s %6d %12.2f %7.1f%%  %s x %s\n" % (k, len(vs), s / len(vs) / 1e3, 100 * s / total, vs[-1][1], vs[-1][2]))
    print(open(out).read())


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none (per launch)\n")
        for r in rows[2:]:
            f.write("\n== %s\n" % r[idx["Kernel Name"]][:100])
            for w in want:
                if w in idx:
                    f.write("   %-62s %16s %s\n" % (w, r[idx[w]][:16], units[idx[w]]))
            top = sorted(((float(r[idx[h]].replace(",", "") or 0), h) for h in stalls), reverse=True)[:5]
            f.write("   top stalls (warps per issue): " + ", ".join("%s %.2f" % (
                h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, h in top) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    launches(sys.argv[1], sys.argv[3] + "_launches.txt")
    full(sys.argv[2], sys.argv[3] + "_kernels.txt")

"""Per-source-line aggregation of an ncu source page (SASS view): executed warp instructions and stall samples per line.
usage: python scripts/ncu_lines.py <report.ncu-rep> <kernel regex> <object file with -lineinfo> [launch index] [mangled-name regex]
Joins `ncu --page source --csv` (per SASS instruction) with `nvdisasm -g` line info by instruction offset."""
import csv
import re
import subprocess
import sys
import tempfile
import os
from collections import defaultdict


def line_map(obj, kernel_pat):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = None
    for i, l in enumerate(txt):
        if ".section" in l and ".text." in l and re.search(kernel_pat, l):
            start = i
            break
    cur = None
    m = {}
    for l in txt[start + 1:]:
        if l.lstrip().startswith(".section"):
            break
        mm = re.search(r'//## File ".*?([^/"]+)", line (\d+)(.*)', l)
        if mm:
            if "inlined at" in mm.group(3) and cur is not None and False:
                continue
            cur = (mm.group(1), int(mm.group(2)))
            continue
        mi = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if mi:
            m[int(mi.group(1), 16)] = (cur, mi.group(2).strip())
    return m


def main():
    rep, pat, obj = sys.argv[1], sys.argv[2], sys.argv[3]
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                         capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(raw.splitlines()):
        if row and row[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
            continue
        if cur is not None:
            cur.append(row)
    rows = blocks[which]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    lm = line_map(obj, sys.argv[5] if len(sys.argv) > 5 else pat)
    base = None
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in rows[1:]:
        try:
            a = int(r[ix["Address"]], 16)
        except ValueError:
            continue
        if base is None:
            base = a
        off = a - base
        line = lm.get(off, (("?", 0), ""))[0]
        ex = int(r[ix["Instructions Executed"]] or 0)
        sa = int(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
        ni = int(r[ix["Warp Stall Sampling (Not-issued Samples)"]] or 0)
        for t in (agg[line], tot):
            t[0] += ex; t[1] += sa; t[2] += ni
    print("total warp instructions %d, samples %d (not issued %d)" % tuple(tot))
    print("%-28s %12s %6s %9s %6s" % ("line", "warp_inst", "%", "samples", "%"))
    for line, v in sorted(agg.items(), key=lambda kv: (str(kv[0][0]), kv[0][1] if kv[0] else 0)):
        if v[0] * 200 < tot[0] and v[1] * 200 < tot[1]:
            continue
        print("%-28s %12d %5.1f%% %9d %5.1f%%" % ("%s:%d" % line if line else "?", v[0], 100.0 * v[0] / max(tot[0], 1), v[1], 100.0 * v[1] / max(tot[1], 1)))


if __name__ == "__main__":
    main()

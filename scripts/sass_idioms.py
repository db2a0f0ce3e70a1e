#!/usr/bin/env python
"""Counts the SASS idioms that show which hardware paths the kernels of libcloud_merger_gpu.so use (cuobjdump -sass, no GPU
needed): TMA bulk copies (UBLKCP), warp reductions in hardware (REDUX), hardware-aggregated shared-memory increments
(ATOMS.POPC), 16-byte streaming loads (LDG.E.128), votes (VOTE), fused multiply-adds (FFMA -- must be ZERO in the kernels
whose arithmetic has to match the PCL CPU build bit for bit), nanosleep back-off (NANOSLEEP), grid-wide tickets (ATOMG).
Writes a table to stdout; the round's copy lives under profiles/."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cloud_merger_b200", "libcloud_merger_gpu.so")
# column name -> regular expression on the opcode (with its modifiers)
IDIOMS = [("UBLKCP", r"^UBLKCP"), ("REDUX", r"^C?REDUX"), ("ATOMS.POPC", r"^ATOMS\.POPC"), ("LDG.128", r"^LDG\..*128"),
          ("STG.128", r"^STG\..*128"), ("VOTE", r"^VOTEU?\b"), ("FFMA", r"^FFMA"), ("FMUL", r"^FMUL"), ("FADD", r"^FADD"),
          ("F2I", r"^F2I"), ("NANOSLEEP", r"^NANOSLEEP"), ("ATOMG", r"^ATOMG|^RED\."), ("BAR", r"^BAR\."), ("SYNCS", r"^SYNCS")]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, counts, sizes = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(cm::.*$|\(.*$", "", kern).replace("void cm::", "").replace("cm::", "")
        counts.setdefault(kern, collections.Counter())
        continue
    if kern is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[kern]["_n"] += 1
    for idiom, rx in IDIOMS:
        if re.search(rx, op):
            counts[kern][idiom] += 1
print("# SASS idioms of %s" % os.path.relpath(LIB, ROOT))
print("# cubin architectures: %s" % ", ".join(arch))
print("# %-58s %6s " % ("kernel", "instr") + " ".join("%10s" % i for i, _ in IDIOMS))
tot = collections.Counter()
for k, c in counts.items():
    print("  %-58s %6d " % (k[:58], c["_n"]) + " ".join("%10d" % c[i] for i, _ in IDIOMS))
    tot.update(c)
print("  %-58s %6d " % ("TOTAL", tot["_n"]) + " ".join("%10d" % tot[i] for i, _ in IDIOMS))
bad = [k for k, c in counts.items() if c["FFMA"] and ("transform_crop" in k or "key_hist" in k or "route" in k or "giant" in k or "zone" in k)]
print("# FFMA in the kernels whose float arithmetic must match PCL's unfused CPU build (transform, voxel / route keys, zones): %s"
      % ("NONE" if not bad else ", ".join(bad)))

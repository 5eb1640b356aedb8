#!/usr/bin/env python
"""Warp-stall samples of one kernel of an .ncu-rep (captured with --set full --import-source on), by source region and by
opcode.  The SASS addresses of the report are joined with nvdisasm's line table of the SAME build's object file.

    python tools/ncu_stall_regions.py <report.ncu-rep> <object.o> <mangled kernel name>
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kernel = sys.argv[1:4]
REASONS = ("selected", "wait", "math", "short_sb", "not_selected", "dispatch", "no_inst", "branch_resolving", "long_sb", "barrier")

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kernel + ":"))
end = start + 1
while end < len(dis) and not dis[end].startswith("\t.section"):
    end += 1
line_of, cur = {}, None
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = cur

rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, KeyError):
        return 0.0


base = int(data[0][ix["Address"]], 16)
total = sum(val(r, "# Samples") for r in data)
insts = sum(val(r, "Instructions Executed") for r in data)
by_region = collections.defaultdict(collections.Counter)
by_op = collections.defaultdict(collections.Counter)
for r in data:
    where = line_of.get(int(r[ix["Address"]], 16) - base)
    region = "?" if where is None else (where[0] if where[0].endswith(".hpp") else f"{where[0]}:{where[1] // 50 * 50}+")
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]].strip())
    op = m.group(2) if m else "?"
    for tgt in (by_region[region], by_op[op]):
        tgt["samples"] += val(r, "# Samples")
        tgt["inst"] += val(r, "Instructions Executed")
        for s in REASONS:
            tgt[s] += val(r, "stall_" + s)
print(f"{kernel}\n{total:.0f} samples, {insts:.4g} warp instructions; columns: % of all samples")
for title, table in (("source region (file:line/50)", by_region), ("opcode", by_op)):
    print(f"\n{title:34s} {'samp%':>6s} {'inst%':>6s} | " + " ".join(f"{s[:7]:>7s}" for s in REASONS))
    for k, c in sorted(table.items(), key=lambda x: -x[1]["samples"])[:18]:
        print(f"{k:34s} {100 * c['samples'] / total:6.1f} {100 * c['inst'] / insts:6.1f} | " + " ".join(f"{100 * c[s] / total:7.1f}" for s in REASONS))

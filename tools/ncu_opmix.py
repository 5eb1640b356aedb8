"""Per-opcode executed-instruction mix and stall samples from `ncu --page source --csv` of one kernel.
Usage: ncu -i x.ncu-rep --page source --csv > src.csv ; python tools/ncu_opmix.py src.csv [warp_sites]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ops, samples = Counter(), Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iE:
        continue
    src = r[iS].strip()
    parts = src.split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "LDL", "STL")) else "")
    n = int(r[iE] or 0)
    ops[op] += n
    samples[op] += int(r[iSm] or 0)
    tot += n
print(f"total warp-instructions {tot:.3e}  per unit {tot / units:.1f}")
ts = sum(samples.values())
for op, n in ops.most_common(28):
    print(f"{op:14s} {n / units:9.1f} /unit  {100 * n / tot:5.1f}%   stall samples {100 * samples[op] / max(ts, 1):5.1f}%")

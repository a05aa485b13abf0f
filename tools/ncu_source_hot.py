#!/usr/bin/env python
"""Per-instruction view of an `ncu --page source --csv` dump: sample share by SASS region and by
opcode, plus the hottest instructions.  usage: ncu_source_hot.py src.csv [n_top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in body)
execd = sum(int(r[ix["Instructions Executed"]]) for r in body)
print(f"instructions {len(body)}  samples {tot}  warp-instructions executed {execd}")
by_op, by_op_exec = Counter(), Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stall_tot = Counter()
for r in body:
    op = r[ix["Source"]].split()[0]
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    by_op[op] += int(r[ix["# Samples"]])
    by_op_exec[op] += int(r[ix["Instructions Executed"]])
    for h in stall_cols:
        stall_tot[h] += int(r[ix[h]] or 0)
print("samples by opcode:", ", ".join(f"{o}={100*v/tot:.1f}%" for o, v in by_op.most_common(14)))
print("executed by opcode:", ", ".join(f"{o}={100*v/execd:.1f}%" for o, v in by_op_exec.most_common(14)))
print("stall reasons:", ", ".join(f"{h[6:]}={100*v/tot:.1f}%" for h, v in stall_tot.most_common(10)))
# cumulative profile along the address axis in 20 buckets
n = len(body)
B = 24
for b in range(B):
    seg = body[b * n // B:(b + 1) * n // B]
    s = sum(int(r[ix["# Samples"]]) for r in seg)
    e = sum(int(r[ix["Instructions Executed"]]) for r in seg)
    ops = Counter(r[ix["Source"]].split()[0].split(".")[0] for r in seg)
    print(f"  [{b * n // B:5d},{(b + 1) * n // B:5d}) samples {100*s/tot:5.1f}%  executed {100*e/execd:5.1f}%  {ops.most_common(4)}")
top = sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    st = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
    print(f"  {100*int(r[ix['# Samples']])/tot:5.2f}%  {r[ix['Source']].strip()[:70]:70s} {st}")

#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export: executed instructions by opcode, stall samples, hot spots.
usage: ncu -i rep.ncu-rep --page source --csv > x.csv ; python tools/ncu_sass_summary.py x.csv [n_units]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = rows[hdr_i + 1:]
ops, stall = Counter(), Counter()
tot = tot_thr = 0
for r in data:
    if len(r) < len(hdr):
        continue
    sass = r[col["Source"]].strip()
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    op = op.split(".")[0]
    ex = int(r[col["Instructions Executed"]] or 0)
    thr = int(r[col["Thread Instructions Executed"]] or 0)
    ops[op] += ex
    stall[op] += int(r[col["Warp Stall Sampling (All Samples)"]] or 0)
    tot += ex
    tot_thr += thr
print("SASS instructions: %d   warp-instructions executed: %d   thread-instructions: %d" % (len(data), tot, tot_thr))
if units:
    print("per unit: %.2f warp-instr*32, %.2f thread-instr" % (tot * 32 / units, tot_thr / units))
print("\nopcode        executed     %%   stall-samples")
for op, n in ops.most_common(28):
    print("%-12s %10d %5.1f %8d" % (op, n, 100.0 * n / tot, stall[op]))
# shared memory conflicts
if "L1 Wavefronts Shared" in col:
    w = sum(int(r[col["L1 Wavefronts Shared"]] or 0) for r in data if len(r) >= len(hdr))
    wi = sum(int(r[col["L1 Wavefronts Shared Ideal"]] or 0) for r in data if len(r) >= len(hdr))
    print("\nshared wavefronts %d ideal %d (x%.2f)" % (w, wi, w / max(wi, 1)))

#!/usr/bin/env python
"""Top SASS instructions of an `ncu --page source --csv` export by stall samples, with executed counts and dominant stall reason."""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = [r for r in csv.reader(open(path, errors="replace")) if r and not r[0].startswith("==")]
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    if r == hdr:      # the export holds one section per kernel launch: keep the first
        break
    data.append(r)
iS, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
total = sum(int(r[iS]) for r in data)
print(rows[0][1][:120])
agg = {h: sum(int(r[i]) for r in data) for i, h in stall_cols}
print(f"{len(data)} instructions, {total} samples; stall mix: " + ", ".join(f"{k[6:]} {100 * v / total:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
order = sorted(range(len(data)), key=lambda k: -int(data[k][iS]))[:top]
for k in sorted(order):
    r = data[k]
    st = sorted(((int(r[i]), h) for i, h in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {100 * int(r[iS]) / total:5.1f}% exec={int(r[iE]):>9d}  {r[1].strip()[:70]:70s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")

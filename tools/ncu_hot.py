"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples with their dominant reasons."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]
sect = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
lo = starts[sect]
hi = starts[sect + 1] if sect + 1 < len(starts) else len(rows)
print("sections:", len(starts), "| showing", sect, rows[lo][1][:90])
hdr = rows[lo + 1]
data = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
i_src, i_samp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[i_samp] or 0) for r in data)
print("total samples", tot)
agg = {}
for _, h in stall_cols:
    agg[h] = 0
for r in data:
    for i, h in stall_cols:
        agg[h] += int(r[i] or 0)
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
order = sorted(range(len(data)), key=lambda k: -int(data[k][i_samp] or 0))[:top]
for k in sorted(order):
    r = data[k]
    reasons = sorted(((int(r[i] or 0), h) for i, h in stall_cols if int(r[i] or 0)), reverse=True)[:3]
    print(f"{k:5d} {int(r[i_samp]):6d} ({100*int(r[i_samp])/tot:4.1f}%) exec {r[i_exec]:>8s}  {r[i_src].strip()[:70]:70s} {reasons}")

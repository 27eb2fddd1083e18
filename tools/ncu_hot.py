"""Top stall locations of a kernel from an .ncu-rep source page (SASS view).
Usage: python tools/ncu_hot.py file.ncu-rep [topN]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
print(rows[0][1][:120], "total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ranked = sorted(enumerate(data), key=lambda x: -int(x[1][idx["# Samples"]] or 0))[:top]
for pos, r in ranked:
    n = int(r[idx["# Samples"]] or 0)
    st = sorted(((int(r[idx[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{pos:5d} {n:5d} {100.0*n/tot:5.1f}%  {r[idx['Source']].strip()[:70]:70s} {st}")

"""Kernel-wide stall-reason totals from the SASS source page of an .ncu-rep.  Usage: python tools/ncu_stalls.py file.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; data = rows[2:]
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in cols}
s = sum(tot.values())
print(rows[0][1][:100], "samples", s)
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
    print(f"  {k:28s} {v:7d} {100.0*v/s:5.1f}%")

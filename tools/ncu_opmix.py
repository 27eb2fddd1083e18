"""Dynamic opcode mix of a kernel from an .ncu-rep (SASS source page): warp-level instructions executed per opcode."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
mix = collections.Counter(); tot = 0
for r in rows[2:]:
    src = r[idx["Source"]].strip().split()
    if not src: continue
    op = src[1] if src[0].startswith("@") else src[0]
    op = op.split(".")[0]
    n = int(r[idx["Instructions Executed"]] or 0)
    mix[op] += n; tot += n
print(rows[0][1][:100], "total warp instr", tot)
for op, n in mix.most_common(28):
    print(f"{op:10s} {n:12d} {100.0*n/tot:5.1f}%")

"""Registers / stack (spill) bytes of every kernel in libggp.so:  python tools/resusage.py [substring ...]"""
import re, subprocess, sys
so = "generalizedgrosspitaevskii.jl_b200/libggp.so"
out = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
names = []
cur = None
rows = []
for line in out.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and cur:
        rows.append((cur, int(m.group(1)), int(m.group(2))))
        cur = None
dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
for (raw, reg, stack), d in zip(rows, dem):
    d = re.sub(r"\(ggp::\w+<.*?>\)$", "", d).replace("void ggp::", "").replace("(int)", "")
    if all(s in d for s in sys.argv[1:]):
        print(f"{d:60s} REG {reg:4d} STACK {stack:5d}")

"""A/B check of a kernel variant: run the same problem in a child process with and without an environment
switch and compare the final fields bit for bit.  python tools/ab_check.py ENVVAR [N] [dtype] [nsteps]"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ggp_b200 as G
    import problems as P
    N, dt, ns, out = int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), sys.argv[5]
    pb = P.kerr2d(G, N=N, dtype=np.dtype(dt).type, nsteps=ns)
    prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    _, sol = G.solve(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1)
    np.save(out, sol[0][-1])
    sys.exit(0)
var = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
dt = sys.argv[3] if len(sys.argv) > 3 else "complex64"
ns = int(sys.argv[4]) if len(sys.argv) > 4 else 20
res = []
for on in (False, True):
    env = dict(os.environ)
    if on:
        env[var] = "1"
    out = f"/tmp/ab_{int(on)}.npy"
    subprocess.run([sys.executable, __file__, "--child", str(N), dt, str(ns), out], env=env, check=True)
    res.append(np.load(out))
d = np.linalg.norm((res[0] - res[1]).ravel()) / np.linalg.norm(res[0].ravel())
print(f"{var}: N={N} {dt} {ns} steps: rel L2 difference {d:.3e}  identical={np.array_equal(res[0], res[1])}  norm={np.linalg.norm(res[0].ravel()):.6e}")

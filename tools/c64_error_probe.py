"""Where does the ComplexF32 error of the GPU path come from?  Compare c64 runs with c128 runs of the same
problem (both on the GPU), for the full step, dispersion only, and nonlinearity only."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ggp_b200 as G
import ggp_oracle as O
import problems as P

def run(ns, dtype, nsteps, disp=True, nl=True, N=256):
    pb = P.kerr2d(ns, N=N, dtype=dtype, nsteps=nsteps)
    kw = dict(pb["kwargs"])
    if not disp: kw.pop("dispersion")
    if not nl: kw.pop("nonlinearity")
    prob = ns.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **kw)
    _, sol = ns.solve(prob, ns.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1)
    return sol[0][-1].astype(np.complex128)

def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(b)

for nsteps in (100, 1000):
    for tag, kw in (("full", {}), ("dispersion only", dict(nl=False)), ("nonlinearity only", dict(disp=False))):
        g32 = run(G, np.complex64, nsteps, **kw)
        g64 = run(G, np.complex128, nsteps, **kw)
        line = f"{nsteps:5d} steps {tag:18s} gpu32-gpu64 {rel(g32, g64):.3e}  norm drift gpu32 {np.linalg.norm(g32)/np.linalg.norm(run(G, np.complex64, 0, **kw) if False else g64) - 1:+.2e}"
        if nsteps <= 1000:
            o32 = run(O, np.complex64, nsteps, **kw)
            line += f"  oracle32-gpu64 {rel(o32, g64):.3e}  gpu32-oracle32 {rel(g32, o32):.3e}"
        print(line, flush=True)

#!/usr/bin/env python
"""BASELINE config C2 over its FULL tspan on the GPU (2048^2 ComplexF32, dt = 1e-3, 10 000 steps) against the CPU
oracle runs of tests/golden/make_c2_full_tspan.py: the committed subsample fixture (every 16th point, both oracle
precisions, at 20 / 100 / 1000 / 2000 / 5000 / 10 000 steps) and -- when gpurun_ship/c2_full_tspan_*.npy travelled
with the snapshot -- the whole final fields.  Prints / writes the three mutual distances north_star asks for:
ours - fp32 oracle, ours - fp64 oracle, fp32 oracle - fp64 oracle.   python tools/c2_full_tspan_gpu.py [out.json]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ggp_b200 as G  # noqa: E402
import problems as P  # noqa: E402


def rel(a, b):
    a = np.asarray(a, dtype=np.complex128)
    b = np.asarray(b, dtype=np.complex128)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def run(fix, ship=None):
    n, steps, stride = int(fix["n"]), int(fix["steps"]), int(fix["stride"])
    marks = [int(m) for m in fix["marks"]]
    pb = P.kerr2d(G, N=n, dtype=np.complex64, nsteps=steps)
    prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    it = G.init(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1)
    assert it.steps_per_save == steps
    rows, done = [], 0
    final = None
    for m in marks:
        it.advance(m - done)
        done = m
        u = it.fetch()[0].copy()
        sub = u[::stride, ::stride]
        rows.append(dict(steps=m, ours_vs_o32=rel(sub, fix[f"sub32_{m}"]), ours_vs_o64=rel(sub, fix[f"sub64_{m}"]),
                         o32_vs_o64_subsample=rel(fix[f"sub32_{m}"], fix[f"sub64_{m}"]),
                         o32_vs_o64_full_field=float(fix["o32_vs_o64"][marks.index(m)]),
                         norm_ours=float(np.linalg.norm(u.astype(np.complex128).ravel())),
                         norm_o64=float(fix["norm64"][marks.index(m)]), norm_o32=float(fix["norm32"][marks.index(m)])))
        final = u
    it.close()
    out = dict(workload=f"C2: 2-D scalar Kerr GPE {n}^2 ComplexF32, dt=1e-3, {steps} steps (full tspan)",
               oracle=str(fix["how"]), subsample_stride=stride, checkpoints=rows)
    if ship is not None:
        o64, o32 = ship
        out["full_field_final"] = dict(ours_vs_o32=rel(final, o32), ours_vs_o64=rel(final, o64), o32_vs_o64=rel(o32, o64))
    return out


if __name__ == "__main__":
    fix = np.load(os.path.join(ROOT, "tests", "golden", "c2_full_tspan_v1.npz"))
    ship = None
    p64 = os.path.join(ROOT, "gpurun_ship", "c2_full_tspan_c128.npy")
    p32 = os.path.join(ROOT, "gpurun_ship", "c2_full_tspan_c64.npy")
    if os.path.exists(p64) and os.path.exists(p32):
        a, b = np.load(p64), np.load(p32)
        if a.shape == (int(fix["n"]),) * 2:
            ship = (a, b)
    out = run(fix, ship)
    print(json.dumps(out, indent=1))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)

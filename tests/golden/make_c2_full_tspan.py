#!/usr/bin/env python
"""CPU oracle for BASELINE config C2 over its FULL tspan: 2-D scalar Kerr GPE 2048^2, dt = 1e-3, 10 000 steps
(north_star: "Deterministic runs must match ... after the full tspan"; SURVEY §8(d) parity gates).

Test infrastructure: runs the CPU restatement of the reference's step (oracle/ggp_fast_cpu.py = the oracle's
tables and step order on torch's CPU kernels, checked against oracle/ggp_oracle.py step for step in
tests/test_fast_cpu.py; the line-by-line NumPy oracle itself with --numpy) in ComplexF64 and ComplexF32 and writes

  tests/golden/c2_full_tspan_v1.npz   committed: every 16th point of the field in both precisions (128 x 128) at each checkpoint,
                                       norms, the fp32-vs-fp64 oracle distance over the whole field, checkpoints of
                                       those distances along the run
  gpurun_ship/c2_full_tspan_{c64,c128}.npy   NOT committed (git-ignored, travels with gpurun): the whole final
                                       fields for the one-off full-field comparison of tools/c2_full_tspan_gpu.py

Usage: python tests/golden/make_c2_full_tspan.py [--steps 10000] [--threads 6] [--n 2048] [--numpy]
Takes ~1 h of CPU on 6 threads (fp64 ~0.15 s/step, fp32 ~0.07 s/step with MKL).
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ggp_fast_cpu as F  # noqa: E402
import ggp_oracle as O  # noqa: E402
import problems as P  # noqa: E402


def rel(a, b):
    a = np.asarray(a, dtype=np.complex128)
    b = np.asarray(b, dtype=np.complex128)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def run(dtype, n, steps, threads, marks, use_numpy):
    pb = P.kerr2d(O, N=n, dtype=dtype, nsteps=steps)
    prob = O.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    if use_numpy:
        it = O.StrangSplittingIterator(prob, pb["tspan"], dt=pb["dt"], nsaves=1, fft_workers=threads)
    else:
        it = F.FastStrang(prob, pb["tspan"], dt=pb["dt"], nsaves=1, threads=threads)
    assert it.steps_per_save == steps, (it.steps_per_save, steps)
    snaps = {}
    t = it.ts[0]
    t0 = time.time()
    for s in range(1, steps + 1):
        t = t + it.dt
        it.step(t, it.dt)
        if s in marks:
            snaps[s] = np.array(it.u[0] if use_numpy else it.state())
            print(f"  {np.dtype(dtype).name} step {s}: {time.time() - t0:.0f} s", flush=True)
    return snaps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--threads", type=int, default=6)
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--numpy", action="store_true", help="use the line-by-line NumPy oracle instead of the torch port")
    ap.add_argument("--out", default=os.path.join(HERE, "c2_full_tspan_v1.npz"))
    a = ap.parse_args()
    marks = sorted({a.steps} | {m for m in (20, 100, 1000, 2000, 5000) if m < a.steps})
    s64 = run(np.complex128, a.n, a.steps, a.threads, marks, a.numpy)
    s32 = run(np.complex64, a.n, a.steps, a.threads, marks, a.numpy)
    ship = os.path.join(ROOT, "gpurun_ship")
    os.makedirs(ship, exist_ok=True)
    np.save(os.path.join(ship, "c2_full_tspan_c128.npy"), s64[a.steps])
    np.save(os.path.join(ship, "c2_full_tspan_c64.npy"), s32[a.steps])
    out = dict(n=a.n, steps=a.steps, stride=16, marks=np.array(marks),
               how="numpy oracle (pocketfft)" if a.numpy else "oracle/ggp_fast_cpu.py (torch CPU kernels, MKL FFT)",
               o32_vs_o64=np.array([rel(s32[m], s64[m]) for m in marks]),
               norm64=np.array([np.linalg.norm(s64[m].ravel()) for m in marks]),
               norm32=np.array([np.linalg.norm(s32[m].astype(np.complex128).ravel()) for m in marks]))
    for m in marks:
        out[f"sub64_{m}"] = np.ascontiguousarray(s64[m][::16, ::16])
        out[f"sub32_{m}"] = np.ascontiguousarray(s32[m][::16, ::16])
    np.savez_compressed(a.out, **out)
    print("wrote", a.out, "o32-o64:", dict(zip(marks, out["o32_vs_o64"])))


if __name__ == "__main__":
    main()

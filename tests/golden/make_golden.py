"""Generate tests/golden/golden_v1.npz from the CPU oracle (oracle/ggp_oracle.py).

The reference ships no golden vectors and cannot be executed in this image (no Julia), so these
fixtures pin the ORACLE (itself pinned by the reference's known-answer tests,
tests/test_oracle_known_answers.py) for regression, and give the GPU tests a committed target that
does not depend on re-running the oracle.   Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (factory, kwargs, noise_seed)
    "kerr2d_c128": ("kerr2d", dict(N=32, dtype="complex128", nsteps=8, L=16.0), None),
    "kerr2d_c64": ("kerr2d", dict(N=32, dtype="complex64", nsteps=8, L=16.0), None),
    "quick_start_kerr": ("quick_start", dict(N=32, kerr=True), None),
    "exciton_polariton": ("exciton_polariton", dict(N=16, nsaves=4, tspan=(0, 1.5625)), None),
    "exciton_polariton_time_pump": ("exciton_polariton", dict(N=16, nsaves=2, tspan=(0, 1.6), dt=0.05, time_pump=True), None),
    "bistability": ("bistability", dict(n=64, nsaves=4, tspan=(0, 25.78125)), None),
    "windowed_ft_noise": ("windowed_ft", dict(ntraj=8), 11),
    "truncated_wigner_2d_noise": ("truncated_wigner", dict(ntraj=4, N=16, ndim=2, tspan=(0, 0.5)), 5),
    "kerr3d_c128": ("kerr3d", dict(N=8, dtype="complex128", nsteps=4, L=8.0), None),
    # field- and position-dependent noise amplitudes (docs/src/stochastic_simulations.md:62-86)
    "noise_field_1d": ("noise_forms", dict(form="field", ndim=1, M=1, N=32, ntraj=3), 7),
    "noise_profile_q2_2d_two_comp": ("noise_forms", dict(form="both", ndim=2, M=2, N=16, ntraj=2), 9),
    # the generic plan (tests/problems.py: generic): axes of any length, three components with a 3x3 matrix dispersion,
    # an SMatrix nonlinearity meeting an SVector potential (the matrix-vector product of src/kernels.jl:9,44)
    "generic_np2_2d": ("generic", dict(case="np2_2d", dtype="complex128"), None),
    "generic_rabi_vv": ("generic", dict(case="rabi_vv", dtype="complex128"), None),
    "generic_m3_matdisp": ("generic", dict(case="m3_matdisp", dtype="complex128"), None),
}


def noise_source_for(seed):
    rng = np.random.default_rng(seed)

    def src(shape, dtype):
        if np.issubdtype(dtype, np.complexfloating):
            return ((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)).astype(dtype)
        return rng.standard_normal(shape).astype(dtype)
    return src


def build(ns, name):
    import problems as P
    fac, kw, seed = CASES[name]
    kw = dict(kw)
    if "dtype" in kw:
        kw["dtype"] = np.dtype(kw["dtype"]).type
    return getattr(P, fac)(ns, **kw), seed


def run_oracle(name, record=None):
    import ggp_oracle as O
    pb, seed = build(O, name)
    prob = O.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    kw = {}
    if seed is not None:
        kw = dict(noise_source=noise_source_for(seed), record_noise=record)
    ts, sol = O.solve(prob, O.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"],
                      save_start=pb.get("save_start", True), **kw)
    return ts, sol


def export_inputs():
    """Inputs for tests/golden/make_golden_ref.jl (the REAL reference, run by whoever has Julia): every case's
    initial fields and, for the stochastic cases, the noise arrays in the reference's draw order, as plain .npy."""
    import ggp_oracle as O
    dst = os.path.join(HERE, "ref_inputs_v1")
    os.makedirs(dst, exist_ok=True)
    for name in CASES:
        pb, seed = build(O, name)
        for c, x in enumerate(pb["u0"]):
            np.save(os.path.join(dst, f"{name}__u0_{c}.npy"), np.ascontiguousarray(x))
        if seed is not None:
            rec = []
            run_oracle(name, record=rec)
            for k, z in enumerate(rec):
                np.save(os.path.join(dst, f"{name}__xi_{k}.npy"), np.ascontiguousarray(z))
        print("exported", name)
    print("wrote", dst)


if __name__ == "__main__":
    if "--export-inputs" in sys.argv:
        export_inputs()
        sys.exit(0)
    out = {}
    for name in CASES:
        ts, sol = run_oracle(name)
        out[name + "/ts"] = ts
        for c, s in enumerate(sol):
            out[f"{name}/u{c}"] = s[-1]          # final save only (keeps the fixture small)
        print(name, [s.shape for s in sol], float(np.abs(sol[0][-1]).max()))
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")

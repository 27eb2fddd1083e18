"""GPU parity of the generic plan (csrc/generic_plan.cuh) against the CPU oracle: FFT axes of any length (Bluestein),
three and four components, SMatrix-valued nonlinearities with every kind of potential table -- and the generic plan
against the fused kernels on problems both can run (GGP_FORCE_GENERIC=1).

Tolerances: relative L2 over the saved solution <= 1e-10 (ComplexF64), <= 1e-4 (ComplexF32), as BASELINE.json."""
import os

import numpy as np
import pytest

import ggp_oracle as O
import problems as P
from test_gpu_parity import run_both, rel_l2, TOL

pytestmark = pytest.mark.gpu

CASES = ["np2_1d", "prime_1d", "np2_2d", "np2_mixed", "np2_3d", "np2_long", "m3_vec", "m3_matdisp", "m4_mat",
         "rabi", "rabi_vs", "rabi_vv", "rabi_vm", "m3_np2_noise"]
NOISY = {"prime_1d", "m3_np2_noise"}


@pytest.fixture(scope="module")
def G():
    import ggp_b200
    ggp_b200.load()
    assert ggp_b200.lib.load().ggp_device_count() >= 1
    return ggp_b200


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
@pytest.mark.parametrize("case", CASES)
def test_generic_plan_matches_oracle(G, case, dtype):
    g, o = run_both(G, P.generic, noise_seed=5 if case in NOISY else None, case=case, dtype=dtype)
    assert np.abs(o[0]).max() > 0
    err = rel_l2(g, o)
    print(f"\n{case} {np.dtype(dtype).name}: rel L2 {err:.3e}")
    assert err <= TOL[np.dtype(dtype)]


def test_rabi_oscillation_known_answer(G):
    """A constant SMatrix nonlinearity Omega*sigma_x without dispersion is a Rabi rotation:
    u1(t) = cos(Omega t) u1(0) - i sin(Omega t) u2(0)  -- analytic, independent of the oracle."""
    om, T, n = 0.7, 1.0, 24
    rng = np.random.default_rng(0)
    a = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex128)
    b = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex128)
    prob = G.GrossPitaevskiiProblem((a, b), (1.0,), nonlinearity=lambda u, p: G.SMatrix([[0 * G.abs2(u[0]), om + 0 * G.abs2(u[0])],
                                                                                        [om + 0 * G.abs2(u[0]), 0 * G.abs2(u[0])]]))
    ts, sol = G.solve(prob, G.StrangSplitting(), (0.0, T), dt=0.01, nsaves=1, show_progress=False)
    u1 = np.cos(om * T) * a - 1j * np.sin(om * T) * b
    u2 = np.cos(om * T) * b - 1j * np.sin(om * T) * a
    assert np.abs(sol[0][-1] - u1).max() < 1e-12 and np.abs(sol[1][-1] - u2).max() < 1e-12


@pytest.mark.parametrize("name", ["quick_start", "bistability", "exciton_polariton", "kerr3d", "noise_field"])
def test_generic_plan_matches_fused_kernels(G, name, monkeypatch):
    """Same problem through the fused kernels and (GGP_FORCE_GENERIC=1) through the generic plan."""
    fac = {
        "quick_start": lambda ns: P.quick_start(ns, N=64),
        "bistability": lambda ns: P.bistability(ns, nsaves=4, tspan=(0, 3300 * 4 / 512)),
        "exciton_polariton": lambda ns: P.exciton_polariton(ns, N=64, nsaves=4, tspan=(0, 2.0)),
        "kerr3d": lambda ns: P.kerr3d(ns, N=16, dtype=np.complex128),
        "noise_field": lambda ns: P.noise_forms(ns, form="both", M=2, ndim=2),
    }[name]

    def run():
        pb = fac(G)
        prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
        return G.solve(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"], show_progress=False,
                       rng=99)[1]

    fused = run()
    monkeypatch.setenv("GGP_FORCE_GENERIC", "1")
    generic = run()
    assert np.abs(fused[0][-1]).max() > 0
    err = rel_l2(generic, fused)
    print(f"\n{name}: generic vs fused rel L2 {err:.3e}")
    # same Philox stream (noise_field), same tables: only the rounding of the transforms differs
    assert err <= 1e-11


def test_axis_beyond_the_built_sizes_is_rejected(G):
    u0 = np.ones(5000, dtype=np.complex128)
    prob = G.GrossPitaevskiiProblem((u0,), (1.0,), dispersion=lambda ks, p: ks[0] ** 2)
    with pytest.raises(G.GgpError if hasattr(G, "GgpError") else Exception):
        G.solve(prob, G.StrangSplitting(), (0.0, 0.1), dt=0.01, nsaves=1, show_progress=False)

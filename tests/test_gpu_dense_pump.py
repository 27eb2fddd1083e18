"""Dense (non-separable) time-dependent pumps, GGP_PUMP_DENSE / ggp_step_dense: the reference re-evaluates any
`pump(r, param, t)` on the whole direct grid at every half-step (evaluate_pump!, /root/reference/src/misc.jl:29-42,
called from src/strang_splitting.jl:81); the backend recognises S(r) a(t) and otherwise does exactly that, uploading
the profiles.  Compared with the oracle (which calls the closure like the reference) on moving pumps."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

import ggp_oracle as O
import problems as P

pytestmark = pytest.mark.gpu
TOL = {np.dtype(np.complex128): 1e-10, np.dtype(np.complex64): 1e-4}


@pytest.fixture(scope="module")
def G():
    import ggp_b200
    ggp_b200.load()
    assert ggp_b200.lib.load().ggp_device_count() >= 1
    return ggp_b200


def rel_l2(a, b):
    num = sum(np.linalg.norm((x.astype(np.complex128) - y.astype(np.complex128)).ravel()) ** 2 for x, y in zip(a, b))
    den = sum(np.linalg.norm(y.astype(np.complex128).ravel()) ** 2 for y in b)
    return float(np.sqrt(num / den))


def moving_1d(ns, dtype=np.complex128, n=256):
    """test/bistability_cycle.jl's cavity with a Gaussian pump spot that travels across it."""
    pb = P.bistability(ns, n=n, nsaves=4, tspan=(0, 20.0))
    real = np.float32 if dtype == np.complex64 else np.float64

    def pump(x, p, t):
        return np.exp(-(x[0] - 60.0 - 6.0 * t) ** 2 / 40.0 ** 2) * (0.3 + 0.02 * t)

    kw = dict(pb["kwargs"])
    kw["pump"] = pump
    u0 = tuple(x.astype(dtype) for x in pb["u0"])
    return dict(u0=u0, lengths=tuple(real(v) for v in pb["lengths"]), kwargs=kw, tspan=(real(0), real(20.0)),
                dt=real(0.05), nsaves=4)


def moving_2d_two_component(ns, dtype=np.complex128, N=64):
    """test/exciton_polariton_test.jl's system with the photon pump spot on a circular orbit (SVector pump)."""
    pb = P.exciton_polariton(ns, N=N, nsaves=3, tspan=(0, 3.0), dt=0.05, dtype=dtype)

    def pump(r, p, t):
        cx, cy = p.L / 2 + 40 * np.cos(0.8 * t), p.L / 2 + 40 * np.sin(0.8 * t)
        return ns.SVector(p.A * np.exp(-((r[0] - cx) ** 2 + (r[1] - cy) ** 2) / 60.0 ** 2), 0)

    kw = dict(pb["kwargs"])
    kw["pump"] = pump
    return dict(u0=pb["u0"], lengths=pb["lengths"], kwargs=kw, tspan=pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"])


def solve(ns, pb):
    prob = ns.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    return ns.solve(prob, ns.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"])[1]


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_moving_pump_1d(G, dtype):
    g, o = solve(G, moving_1d(G, dtype)), solve(O, moving_1d(O, dtype))
    assert np.abs(o[0][-1]).max() > 1e-3
    assert rel_l2(g, o) <= TOL[np.dtype(dtype)]


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_moving_pump_2d_two_components(G, dtype):
    g, o = solve(G, moving_2d_two_component(G, dtype)), solve(O, moving_2d_two_component(O, dtype))
    assert np.abs(o[0][-1]).max() > 1e-3 and np.abs(o[1][-1]).max() > 0
    assert rel_l2(g, o) <= TOL[np.dtype(dtype)]


def test_dense_route_equals_separable_route(G):
    """A separable pump pushed through the dense route (GGP_PUMP_FORCE_DENSE=1) reproduces the S(r) a(t) route."""
    pb = P.bistability(G, nsaves=3, tspan=(0, 3300 * 3 / 512))
    a = solve(G, pb)
    os.environ["GGP_PUMP_FORCE_DENSE"] = "1"
    try:
        prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
        it = G.init(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"])
        assert it.pump_model.dense
        b = G.solve_(it)[1]
        it.close()
    finally:
        del os.environ["GGP_PUMP_FORCE_DENSE"]
    assert np.abs(a[0][-1]).max() > 0
    assert rel_l2(b, a) <= 1e-12

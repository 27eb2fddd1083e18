"""oracle/ggp_fast_cpu.py (torch / MKL restatement used ONLY to time a strong CPU baseline in bench.py) against the
line-by-line oracle, step for step."""
import numpy as np
import pytest

import ggp_fast_cpu as F
import ggp_oracle as O
import problems as P


def _rel(a, b):
    a = np.asarray(a, dtype=np.complex128)
    b = np.asarray(b, dtype=np.complex128)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


@pytest.mark.parametrize("factory,kw,tol", [
    (P.kerr2d, dict(N=64, dtype=np.complex64, nsteps=20), 2e-5),
    (P.kerr2d, dict(N=64, dtype=np.complex128, nsteps=20), 1e-12),
    (P.kerr3d, dict(N=16, dtype=np.complex128, nsteps=6), 1e-12),
    (P.quick_start, dict(N=32, kerr=True), 1e-12),
])
def test_fast_cpu_matches_oracle(factory, kw, tol):
    pb = factory(O, **kw)
    prob = O.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    assert F.supported(prob)
    ref = O.StrangSplittingIterator(prob, pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"])
    fast = F.FastStrang(prob, pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"], threads=2)
    assert fast.dt == ref.dt and fast.steps_per_save == ref.steps_per_save
    t = ref.ts[0]
    for _ in range(min(12, ref.steps_per_save * pb["nsaves"])):
        t = t + ref.dt
        ref.step(t, ref.dt)
        fast.step()
    assert fast.state().dtype == ref.u[0].dtype
    assert _rel(fast.state(), ref.u[0]) <= tol


def test_fast_cpu_declines_what_it_does_not_cover():
    for pb in (P.exciton_polariton(O, N=16, nsaves=1, tspan=(0, 1)),           # two components, pump
               P.windowed_ft(O, ntraj=4),                                        # noise
               P.bistability(O, n=64, nsaves=1, tspan=(0, 10))):                 # pump
        prob = O.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
        assert not F.supported(prob)
    pb = P.kerr2d(O, N=32, dtype=np.complex64, nsteps=4)
    pb["kwargs"]["nonlinearity"] = lambda u, p: O.abs2(u[0]) ** 2                # not c + g|u|^2
    prob = O.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    assert not F.supported(prob)

"""Strided axes of 4096 and 8192 points against the oracle: the geometries only the large-grid benchmarks use
(factorised twiddles `w_N^j = A[j>>6] B[j&63]`, 256 / 512 threads per line, 16-byte tiles) on grids small enough for the
CPU oracle -- a long strided axis next to a short contiguous one.  Separable dispersion (two-factor exp_D with D_line
staged in shared memory) and a coupled one (full table)."""
import numpy as np
import pytest

import ggp_oracle as O

pytestmark = pytest.mark.gpu
TOL = {np.dtype(np.complex128): 1e-10, np.dtype(np.complex64): 1e-4}


@pytest.fixture(scope="module")
def G():
    import ggp_b200
    ggp_b200.load()
    assert ggp_b200.lib.load().ggp_device_count() >= 1
    return ggp_b200


def _problem(ns, shape, dtype, coupled):
    rng = np.random.default_rng(17)
    real = np.float32 if dtype == np.complex64 else np.float64
    n2, n1 = shape
    y = np.arange(n2)[:, None] / n2
    x = np.arange(n1)[None, :] / n1
    env = np.exp(-40 * (y - 0.5) ** 2) * (1 + 0.3 * np.cos(2 * np.pi * x))
    u0 = (env * (1 + 0.05 * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)))).astype(dtype)

    def dispersion(ks, p):
        d = (ks[0] * ks[0] + ks[1] * ks[1]) / 2
        return d + real(0.01) * ks[0] * ks[1] if coupled else d

    def nonlinearity(u, p):
        return real(0.5) * ns.abs2(u[0])

    return dict(u0=(u0,), lengths=(real(8.0), real(64.0)), kwargs=dict(dispersion=dispersion, nonlinearity=nonlinearity))


@pytest.mark.parametrize("coupled", [False, True])
@pytest.mark.parametrize("shape", [(4096, 32), (8192, 16)])
@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_long_strided_axis(G, dtype, shape, coupled):
    outs = []
    for ns in (G, O):
        pb = _problem(ns, shape, dtype, coupled)
        prob = ns.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
        real = np.float32 if dtype == np.complex64 else np.float64
        outs.append(ns.solve(prob, ns.StrangSplitting(), (real(0), real(0.004)), dt=real(0.001), nsaves=2)[1])
    g, o = outs
    num = np.linalg.norm((g[0].astype(np.complex128) - o[0].astype(np.complex128)).ravel())
    den = np.linalg.norm(o[0].astype(np.complex128).ravel())
    assert num / den <= TOL[np.dtype(dtype)]
    assert np.linalg.norm((o[0][-1] - o[0][0]).ravel()) / den > 1e-4          # the field really evolved

"""Pin the CPU oracle against the reference's own known-answer tests.

The reference ships no golden vectors and cannot run here (no Julia); what it does ship is five
test sets whose oracles are analytic.  Each one is re-expressed here with the reference's
tolerance, so that oracle/ggp_oracle.py is a checked reading of src/*.jl before the CUDA path is
compared with it.
"""
import numpy as np
import pytest
import scipy.linalg

import ggp_oracle as O
import problems as P


def _solve(pb, **kw):
    prob = O.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    return O.solve(prob, O.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"],
                   save_start=pb.get("save_start", True), **kw)


def test_resolve_fixed_timestepping_cases():
    # SURVEY Q3 worked examples (src/fixed_time_stepping.jl:14-24)
    dt, ts, sps = O.resolve_fixed_timestepping(0.01, (0, 1), 64)
    assert sps == 2 and dt == 0.0078125 and ts.dtype == np.float64 and len(ts) == 65
    dt, ts, sps = O.resolve_fixed_timestepping(0.01, (0, 0.4), 64)
    assert sps == 1 and dt == 0.4 / 64
    dt, ts, sps = O.resolve_fixed_timestepping(1e-1, (0, 100), 256)
    assert sps == 4 and dt == 0.09765625
    dt, ts, sps = O.resolve_fixed_timestepping(0.05, (0, 3300), 512)
    assert sps == 129 and sps * 512 == 66048
    dt, ts, sps = O.resolve_fixed_timestepping(np.float32(1e-3), (np.float32(0), np.float32(1)), 1)
    assert ts.dtype == np.float32 and np.asarray(dt).dtype == np.float32


def test_grids():
    # src/problem.jl:129-139, AbstractFFTs.fftfreq (negative Nyquist for even n)
    x = O.direct_grid_1d(8, 128)
    assert x[0] == 0 and np.isclose(x[1], 8 / 128) and len(x) == 128
    k = O.reciprocal_grid_1d(8, 8)
    assert np.allclose(k, 2 * np.pi / 8 * np.array([0, 1, 2, 3, -4, -3, -2, -1]))
    k = O.reciprocal_grid_1d(np.float32(8), 8)
    assert k.dtype == np.float32


def test_exp2x2_matches_scipy_expm():
    rng = np.random.default_rng(0)
    for _ in range(50):
        A = rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2))
        m11, m21, m12, m22 = O.exp2x2(A[0, 0], A[1, 0], A[0, 1], A[1, 1])
        ref = scipy.linalg.expm(A)
        got = np.array([[m11, m12], [m21, m22]])
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-13
    # degenerate branch (z -> 0)
    m11, m21, m12, m22 = O.exp2x2(0.3 + 0.1j, 0.0, 0.0, 0.3 + 0.1j)
    assert np.isclose(m11, np.exp(0.3 + 0.1j)) and np.isclose(m22, np.exp(0.3 + 0.1j)) and m12 == 0


def _lg_like(rs, l):
    """A Laguerre-Gauss-like vortex mode (p=0): (x + i sgn(l) y)^|l| exp(-r²) (closed form; the
    reference uses StructuredLight.lg, un-vendored).  Any smooth localised field serves: the
    check is propagation, not the mode shape."""
    X, Y = np.meshgrid(rs, rs, indexing="xy")
    return ((X + 1j * np.sign(l) * Y) ** abs(l)) * np.exp(-(X ** 2 + Y ** 2))


def test_free_propagation_exact_and_wrappers():
    """test/free_propagation.jl: Strang steps with only dispersion == exact spectral propagator;
    the scalar / SVector{1} / SMatrix{1,1} dispersion forms give the same numbers (:18-28)."""
    rng = np.random.default_rng(1234)
    for _ in range(2):
        N = 64
        L = int(rng.integers(4, 11))
        dt = 0.3 * rng.random()
        nsaves = int(rng.integers(50, 201))
        rs = -L / 2 + np.arange(N) * (L / N)
        u0 = _lg_like(rs, int(rng.integers(1, 6))) + _lg_like(rs, int(rng.integers(1, 6)))
        tspan = (0.0, nsaves * dt)
        d1 = lambda ks, p: (ks[0] ** 2 + ks[1] ** 2) / 2
        d2 = lambda ks, p: O.SVector(d1(ks, p))
        d3 = lambda ks, p: O.SMatrix([[d1(ks, p)]])
        k = O.reciprocal_grid_1d(L, N)
        K2 = k[None, :] ** 2 + k[:, None] ** 2
        sols = []
        for d in (d1, d2, d3):
            prob = O.GrossPitaevskiiProblem((u0,), (L, L), dispersion=d)
            ts, sol = O.solve(prob, O.StrangSplitting(), tspan, dt=dt, nsaves=nsaves)
            sols.append(sol[0])
        exact = np.stack([np.fft.ifft2(np.exp(-1j * K2 * t / 2) * np.fft.fft2(u0)) for t in ts])
        for s in sols:
            assert np.linalg.norm(s - exact) <= np.sqrt(np.finfo(float).eps) * np.linalg.norm(exact)
        assert np.array_equal(sols[0], sols[1]) and np.array_equal(sols[0], sols[2])


def test_kerr_self_convergence_and_norm():
    """test/kerr_propagation.jl replaced (StructuredLight.kerr_propagation is un-vendored) by
    second-order self-convergence under dt-halving and norm conservation."""
    N, L, g = 64, 10, 0.7
    rs = -L / 2 + np.arange(N) * (L / N)
    u0 = _lg_like(rs, 1) + _lg_like(rs, -2)
    disp = lambda ks, p: (ks[0] ** 2 + ks[1] ** 2) / 2
    nl1 = lambda psi, p: p.g * O.abs2(psi) / 2
    nl2 = lambda psi, p: O.SVector(nl1(psi, p)[0])
    nl3 = lambda psi, p: O.SMatrix([[nl1(psi, p)[0]]])
    from types import SimpleNamespace
    param = SimpleNamespace(g=g)
    T = 0.2
    out = {}
    for nsteps in (50, 100, 200):
        prob = O.GrossPitaevskiiProblem((u0,), (L, L), dispersion=disp, nonlinearity=nl1, param=param)
        ts, sol = O.solve(prob, O.StrangSplitting(), (0.0, T), dt=T / nsteps, nsaves=1)
        out[nsteps] = sol[0][-1]
        assert abs(np.linalg.norm(sol[0][-1]) / np.linalg.norm(u0) - 1) < 1e-12
    e1 = np.linalg.norm(out[50] - out[100])
    e2 = np.linalg.norm(out[100] - out[200])
    assert 3.5 < e1 / e2 < 4.5  # second order
    ref = out[50]
    for nl in (nl2, nl3):
        prob = O.GrossPitaevskiiProblem((u0,), (L, L), dispersion=disp, nonlinearity=nl, param=param)
        ts, sol = O.solve(prob, O.StrangSplitting(), (0.0, T), dt=T / 50, nsaves=1)
        assert np.allclose(sol[0][-1], ref, rtol=0, atol=1e-14 * np.abs(ref).max())


@pytest.mark.slow
def test_bistability_cycle_known_answer():
    """test/bistability_cycle.jl:56-65: mean |bistability_curve(max|ψ|²) − I(t)| over saves 140:400 ≤ 3e-3."""
    pb = P.bistability(O)
    ts, sol = _solve(pb)
    Is = np.array([pb["I"](t, pb["tspan"][-1], pb["Imax"]) for t in ts])
    n = np.max(O.abs2(sol[0]), axis=-1)
    pred = n * (pb["gamma"] ** 2 / 4 + (pb["g"] * n - pb["delta"]) ** 2)
    err = np.abs(pred - Is)
    assert err[139:400].sum() / len(Is) <= 3e-3          # Julia 140:400 is 1-based inclusive


def test_bistability_wrappers_short():
    """test/bistability_cycle.jl:67-71 on a short prefix: all 27 wrappings reproduce the first."""
    kw = dict(nsaves=4, tspan=(0, 3300 * 4 / 512))
    base = _solve(P.bistability(O, **kw))[1][0]
    assert np.abs(base).max() > 0
    for wd in range(3):
        for wn in range(3):
            for wp in range(3):
                s = _solve(P.bistability(O, wrap=(wd, wn, wp), **kw))[1][0]
                assert np.allclose(s, base, rtol=0, atol=1e-13 * np.abs(base).max())


def test_exciton_polariton_known_answer():
    """test/exciton_polariton_test.jl:48-53: two steady-state relations < 3e-2."""
    pb = P.exciton_polariton(O)
    ts, sol = _solve(pb)
    p = pb["param"]
    N = 128
    nx = O.abs2(sol[1])[-1, N // 2 - 1, N // 2 - 1]    # Julia [N÷2, N÷2, end] is 1-based
    nc = O.abs2(sol[0])[-1, N // 2 - 1, N // 2 - 1]
    r1 = abs(abs(p.Wr - (p.dx + 1j * p.gx / 2 - p.g * nx) * (p.dc + 1j * p.gc / 2) / p.Wr) ** 2 * nx / p.A ** 2 - 1)
    r2 = abs(abs(p.dx + 1j * p.gx / 2 - p.g * nx) ** 2 * nx / p.Wr ** 2 / nc - 1)
    assert r1 < 3e-2 and r2 < 3e-2


def _window(x, par):
    x0, w = par
    return np.exp(-(x - x0) ** 2 / w ** 2)


def windowed_correlation(sol, rs, par1, par2):
    """test/windowed_ft.jl:31-49.  sol: NumPy (ntraj, N)."""
    s1 = sol * _window(rs, par1)[None, :]
    s2 = sol * _window(rs, par2)[None, :]
    f1 = np.fft.ifftshift(np.fft.fft(np.fft.fftshift(s1, axes=1), axis=1), axes=1)
    f2 = np.fft.ifftshift(np.fft.fft(np.fft.fftshift(s2, axes=1), axis=1), axes=1)
    # g1[i, j] = dot(f2[j, :], f1[i, :]) / length(sol); Julia dot conjugates the first argument
    return np.einsum("tj,ti->ij", np.conj(f2), f1) / sol.size


def analytic_commutation(L, N, par1, par2):
    """test/windowed_ft.jl:51-59."""
    rs = -L / 2 + np.arange(N) * (L / N)
    ks = -np.pi / L + np.arange(N) * (2 * np.pi / L)
    w = _window(rs, par1) * np.conj(_window(rs, par2))
    ph = np.exp(1j * (ks[:, None, None] - ks[None, :, None]) * rs[None, None, :])
    return (ph * w[None, None, :]).sum(-1) / (rs[-1] - rs[0]) / 2


def test_windowed_ft_known_answer():
    """test/windowed_ft.jl:92-99: vacuum commutator of windowed Fourier modes, rtol 7e-2."""
    pb = P.windowed_ft(O)
    rng = np.random.default_rng(1234)

    def noise_source(shape, dtype):
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)

    ts, sol = _solve(pb, noise_source=noise_source)
    s = sol[0][0]
    L, N = pb["L"], pb["N"]
    rs = -L / 2 + np.arange(N) * (L / N)
    worst = 0.0
    for x0 in (-1, 0, 1):
        for w in (3, 4, 5):
            corr = windowed_correlation(s, rs, (x0, w), (-x0, w))
            an = analytic_commutation(L, N, (x0, w), (-x0, w))
            rel = np.linalg.norm(corr - an) / max(np.linalg.norm(corr), np.linalg.norm(an))
            worst = max(worst, rel)
    assert worst < 7e-2, worst


def test_pump_times_quirk_q1():
    """pump evaluated at t0, then t_{n+1}+dt/2 and t_{n+1}+dt (one dt late)."""
    t0, tt = O.pump_times((0, 1), 0.25, 1)
    assert t0 == 0 and np.allclose(tt, [[0.375, 0.5], [0.625, 0.75], [0.875, 1.0], [1.125, 1.25]])


def test_noise_point_quirk_q2():
    """kernels.jl:27,41: the `point` handed to the noise amplitude function indexes EVERY grid axis with K[1]."""
    xs, ys = np.arange(4) * 0.5, np.arange(6) * 0.25
    px, py = O._noise_points((xs, ys))
    assert px.shape == (1, 4) and py.shape == (1, 4)
    assert np.array_equal(px.ravel(), xs) and np.array_equal(py.ravel(), ys[:4])      # y[K1], not y[K2]
    with pytest.raises(IndexError):
        O._noise_points((ys, xs))                                                     # n1 = 6 > n2 = 4: BoundsError
    import problems as P
    pb = P.noise_forms(O, form="profile", ndim=2, M=1, N=8, ntraj=2)
    prob = O.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    seen = []
    inner = prob.position_noise_func
    prob.position_noise_func = lambda u, r, p: (seen.append([np.asarray(x).shape for x in r]), inner(u, r, p))[1]
    rng = np.random.default_rng(0)
    src = lambda shape, dtype: (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype)
    O.solve(prob, O.StrangSplitting(), (0, 0.1), dt=0.05, nsaves=1, noise_source=src)
    assert seen and all(s == [(1, 8), (1, 8)] for s in seen)


def test_field_dependent_noise_uses_pre_update_field():
    """kernels.jl:40-49: `fields` is read once, before the update; the noise amplitude eta(fields, point, param) and the
    nonlinear phase are both evaluated on it:  result = cis(-dt G(u)) u - i sqrt(dt) eta(u) xi."""
    rng = np.random.default_rng(2)
    u = (rng.standard_normal((3, 8)) + 1j * rng.standard_normal((3, 8)))
    xi = (rng.standard_normal((3, 8)) + 1j * rng.standard_normal((3, 8)))
    g, alpha, c, dt = 0.7, 0.3, 0.2, 0.05
    out = O.muladd([u], O.multiplicativeIdentity, O.additiveIdentity, O.additiveIdentity, dt,
                   lambda f, p: g * O.abs2(f[0]), lambda f, r, p: c + alpha * abs(f[0]), [xi], None,
                   (np.zeros((1, 8)),))
    want = np.exp(-1j * dt * g * np.abs(u) ** 2) * u - 1j * np.sqrt(dt) * (c + alpha * np.abs(u)) * xi
    assert np.allclose(out[0], want, rtol=1e-14, atol=1e-15)
    post = np.exp(-1j * dt * g * np.abs(u) ** 2) * u
    wrong = post - 1j * np.sqrt(dt) * (c + alpha * np.abs(post + 0.1)) * xi          # any post-update evaluation differs
    assert not np.allclose(out[0], wrong)

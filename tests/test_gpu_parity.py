"""GPU parity tests: the CUDA path (through the C ABI, via the host mirror of the reference's
interface) against the CPU oracle on identical inputs.

Tolerances are BASELINE.json's: relative L2 error over the whole saved solution
  <= 1e-10 for ComplexF64,  <= 1e-4 for ComplexF32.
"""
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
    return float(np.sqrt(num / den)) if den > 0 else float(np.sqrt(num))


def run_both(G, factory, noise_seed=None, **kw):
    pbo = factory(O, **kw)
    pbg = factory(G, **kw)
    prob_o = O.GrossPitaevskiiProblem(pbo["u0"], pbo["lengths"], **pbo["kwargs"])
    rec = []
    okw = {}
    if noise_seed is not None:
        rng = np.random.default_rng(noise_seed)

        def noise_source(shape, dtype):
            if np.issubdtype(dtype, np.complexfloating):
                return ((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)).astype(dtype)
            return rng.standard_normal(shape).astype(dtype)

        okw = dict(noise_source=noise_source, record_noise=rec)
    ts_o, sol_o = O.solve(prob_o, O.StrangSplitting(), pbo["tspan"], dt=pbo["dt"], nsaves=pbo["nsaves"],
                          save_start=pbo.get("save_start", True), **okw)
    prob_g = G.GrossPitaevskiiProblem(pbg["u0"], pbg["lengths"], **pbg["kwargs"])
    ts_g, sol_g = G.solve(prob_g, G.StrangSplitting(), pbg["tspan"], dt=pbg["dt"], nsaves=pbg["nsaves"],
                          save_start=pbg.get("save_start", True), show_progress=False,
                          noise_buffers=rec if noise_seed is not None else None)
    assert np.array_equal(ts_o, ts_g)
    assert all(a.shape == b.shape and a.dtype == b.dtype for a, b in zip(sol_o, sol_g))
    return sol_g, sol_o


@pytest.mark.parametrize("kerr", [False, True])
def test_c1_quick_start(G, kerr):
    """BASELINE config C1 (examples/quick_start.jl), both runs, ComplexF64."""
    g, o = run_both(G, P.quick_start, kerr=kerr)
    assert rel_l2(g, o) <= 1e-10


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_c2_shape_kerr2d(G, dtype):
    """BASELINE config C2's problem at 256^2, 100 steps."""
    g, o = run_both(G, P.kerr2d, N=256, dtype=dtype, nsteps=100)
    assert rel_l2(g, o) <= TOL[np.dtype(dtype)]
    # distance to the fp64 oracle is reported alongside (SURVEY hard part 5)
    if dtype == np.complex64:
        pb = P.kerr2d(O, N=256, dtype=np.complex128, nsteps=100)
        prob = O.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
        _, ref64 = O.solve(prob, O.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1)
        assert rel_l2(g, ref64) <= 1e-4


def test_c2_c64_1000_step_prefix(G):
    """SURVEY hard part 5: the ComplexF32 gate (<= 1e-4) on a 1000-step prefix of C2's problem (256^2),
    against the fp32 oracle AND the fp64 oracle; the three mutual distances are printed."""
    g, o32 = run_both(G, P.kerr2d, N=256, dtype=np.complex64, nsteps=1000)
    pb = P.kerr2d(O, N=256, dtype=np.complex128, nsteps=1000)
    prob = O.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    _, o64 = O.solve(prob, O.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1)
    d_g32, d_g64, d_3264 = rel_l2(g, o32), rel_l2(g, o64), rel_l2(o32, o64)
    print(f"\nc64 1000 steps: ours-fp32oracle {d_g32:.3e}  ours-fp64oracle {d_g64:.3e}  fp32oracle-fp64oracle {d_3264:.3e}")
    assert d_g32 <= 1e-4 and d_g64 <= 1e-4


def test_bistability_prefix_and_wrappers(G):
    """test/bistability_cycle.jl on a prefix of the run (time-dependent separable pump, lossy
    dispersion, Kerr) + the 27 scalar/SVector/SMatrix{1,1} wrappings (:67-71)."""
    kw = dict(nsaves=8, tspan=(0, 3300 * 8 / 512))
    g, o = run_both(G, P.bistability, **kw)
    assert np.abs(o[0]).max() > 0
    assert rel_l2(g, o) <= 1e-10
    for wd in range(3):
        for wn in range(3):
            for wp in range(3):
                pb = P.bistability(G, wrap=(wd, wn, wp), **kw)
                prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
                _, s = G.solve(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"])
                assert np.array_equal(s[0], g[0])


@pytest.mark.slow
def test_bistability_full_known_answer(G):
    """The reference's own criterion on the full 66 048-step run (test/bistability_cycle.jl:56-65)."""
    pb = P.bistability(G)
    prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    ts, sol = G.solve(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"])
    Is = np.array([pb["I"](t, pb["tspan"][-1], pb["Imax"]) for t in ts])
    n = np.max(G.abs2(sol[0]), axis=-1)
    pred = n * (pb["gamma"] ** 2 / 4 + (pb["g"] * n - pb["delta"]) ** 2)
    assert np.abs(pred - Is)[139:400].sum() / len(Is) <= 3e-3


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_exciton_polariton_parity(G, dtype):
    """test/exciton_polariton_test.jl problem (M = 2, 2x2 matrix-exponential dispersion table,
    SVector nonlinearity, SVector pump), 64 steps."""
    g, o = run_both(G, P.exciton_polariton, nsaves=16, tspan=(0, 6.25), dtype=dtype)
    assert np.abs(o[0]).max() > 0
    assert rel_l2(g, o) <= TOL[np.dtype(dtype)]


def test_exciton_polariton_known_answer(G):
    """test/exciton_polariton_test.jl:48-53 on the GPU path."""
    pb = P.exciton_polariton(G)
    prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    ts, sol = G.solve(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"])
    p, N = pb["param"], 128
    nx = G.abs2(sol[1])[-1, N // 2 - 1, N // 2 - 1]
    nc = G.abs2(sol[0])[-1, N // 2 - 1, N // 2 - 1]
    r1 = abs(abs(p.Wr - (p.dx + 1j * p.gx / 2 - p.g * nx) * (p.dc + 1j * p.gc / 2) / p.Wr) ** 2 * nx / p.A ** 2 - 1)
    r2 = abs(abs(p.dx + 1j * p.gx / 2 - p.g * nx) ** 2 * nx / p.Wr ** 2 / nc - 1)
    assert r1 < 3e-2 and r2 < 3e-2


def test_c3_time_dependent_pump_em(G):
    """BASELINE config C3 shape at 64^2: exciton-polariton with the bistability envelope."""
    g, o = run_both(G, P.exciton_polariton, N=64, nsaves=4, tspan=(0, 3.2), dt=0.05, time_pump=True)
    assert np.abs(o[0]).max() > 0
    assert rel_l2(g, o) <= 1e-10


def test_windowed_ft_host_noise_bit_for_bit_dynamics(G):
    """Stochastic run fed the oracle's noise buffer (test mode): deterministic gate."""
    g, o = run_both(G, P.windowed_ft, noise_seed=7, ntraj=500)
    assert rel_l2(g, o) <= 1e-10


def test_truncated_wigner_host_noise_2d(G):
    """BASELINE config C4's problem at 32^2 x 16 trajectories, host-fed noise."""
    g, o = run_both(G, P.truncated_wigner, noise_seed=3, ntraj=16, N=32, ndim=2, tspan=(0, 0.5))
    assert rel_l2(g, o) <= 1e-10


@pytest.mark.parametrize("form,ndim,M,dtype,real_proto", [
    ("field", 1, 1, np.complex128, False),
    ("profile", 1, 1, np.complex128, False),
    ("both", 1, 1, np.complex64, False),
    ("both", 2, 2, np.complex128, False),      # SVector amplitudes, profile at the reference's Q2 points (x[K1], y[K1])
    ("field", 2, 1, np.complex128, True),      # real noise prototype
])
def test_field_and_position_dependent_noise_host_fed(G, form, ndim, M, dtype, real_proto):
    """eta_i(u, r) = P(r) (e_i + sum_j a_ij |u_j|) (docs/src/stochastic_simulations.md:62-86) against the
    oracle, which calls the user's closure with the reference's own `point` (src/kernels.jl:27,41), on identical
    host-fed noise."""
    g, o = run_both(G, P.noise_forms, noise_seed=11, form=form, ndim=ndim, M=M, dtype=dtype, real_proto=real_proto)
    assert rel_l2(g, o) <= TOL[np.dtype(dtype)]
    # the noise really depends on the field / position: switching the form off changes the answer
    g0, _ = run_both(G, P.noise_forms, noise_seed=11, form="field", ndim=ndim, M=M, dtype=dtype, real_proto=real_proto) \
        if form != "field" else (None, None)
    if g0 is not None:
        assert rel_l2(g, g0) > 1e-3


def test_windowed_ft_philox_known_answer(G):
    """test/windowed_ft.jl:92-99 with the in-kernel Philox stream (rtol 7e-2 at 10^4 trajectories)."""
    from test_oracle_known_answers import analytic_commutation, windowed_correlation
    pb = P.windowed_ft(G)
    prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    ts, sol = G.solve(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1, save_start=False, rng=1234)
    s = sol[0][0]
    L, N = pb["L"], pb["N"]
    rs = -L / 2 + np.arange(N) * (L / N)
    worst = 0.0
    for x0 in (-1, 0, 1):
        for w in (3, 4, 5):
            corr = windowed_correlation(s, rs, (x0, w), (-x0, w))
            an = analytic_commutation(L, N, (x0, w), (-x0, w))
            worst = max(worst, np.linalg.norm(corr - an) / max(np.linalg.norm(corr), np.linalg.norm(an)))
    assert worst < 7e-2, worst
    # sharding invariance of the counter-based stream: two half-ensembles == the full ensemble
    halves = []
    for off in (0, 5000):
        u0 = (pb["u0"][0][off:off + 5000],)
        kw = dict(pb["kwargs"])
        kw["noise_prototype"] = tuple(np.empty_like(x) for x in u0)
        pr = G.GrossPitaevskiiProblem(u0, pb["lengths"], **kw)
        it = G.init(pr, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1, save_start=False, rng=1234,
                    batch_offset=off)
        _, sh = G.solve_(it)
        it.close()
        halves.append(sh[0][0])
    assert np.array_equal(np.concatenate(halves, axis=0), s)


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_c5_shape_kerr3d(G, dtype):
    """BASELINE config C5's problem at 32^3 (and a non-cubic grid)."""
    g, o = run_both(G, P.kerr3d, N=32, dtype=dtype, nsteps=10)
    assert rel_l2(g, o) <= TOL[np.dtype(dtype)]


def _free_problem(ns, shape, lengths, dtype, seed=0, potential=None, M=1, kind="scalar"):
    rng = np.random.default_rng(seed)
    u0 = tuple((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dtype) for _ in range(M))

    def dispersion(ks, p):
        s = 0
        for k in ks:
            s = s + k * k
        s = s / 2
        if kind == "scalar":
            return s
        if kind == "diag":
            return ns.SVector(s, 0.5 * s - 0.1j)
        return ns.SMatrix([[s - 0.05j, 0.3], [0.3, 0.7 * s - 0.02j]])

    kw = dict(dispersion=dispersion)
    if potential is not None:
        kw["potential"] = potential
    return dict(u0=u0, lengths=lengths, kwargs=kw)


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_every_line_length_1d(G, n, dtype):
    """Every instantiated FFT size, ragged batch (37 lines: not a multiple of lines-per-CTA)."""
    outs = []
    for ns in (G, O):
        pb = _free_problem(ns, (37, n), (10.0,), dtype)
        prob = ns.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
        outs.append(ns.solve(prob, ns.StrangSplitting(), (0.0, 0.03), dt=0.01, nsaves=1)[1])
    assert rel_l2(*outs) <= (1e-12 if dtype == np.complex128 else 2e-6)


@pytest.mark.parametrize("shape", [(16, 512), (512, 16), (4, 8), (2, 64, 32), (1024, 64)])
@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_rectangular_grids_2d(G, shape, dtype):
    """Non-square grids, tiny fast axes (W clamps to n1), batch dims."""
    nd = 2
    outs = []
    for ns in (G, O):
        pb = _free_problem(ns, shape, (7.0, 11.0), dtype, seed=1)
        prob = ns.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
        outs.append(ns.solve(prob, ns.StrangSplitting(), (0.0, 0.02), dt=0.01, nsaves=2)[1])
    assert rel_l2(*outs) <= (1e-12 if dtype == np.complex128 else 2e-6)


@pytest.mark.parametrize("shape", [(8, 16, 64), (64, 8, 16), (2, 16, 16, 16)])
def test_rectangular_grids_3d(G, shape):
    outs = []
    for ns in (G, O):
        pb = _free_problem(ns, shape, (7.0, 11.0, 5.0), np.complex128, seed=2)
        prob = ns.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
        outs.append(ns.solve(prob, ns.StrangSplitting(), (0.0, 0.02), dt=0.01, nsaves=2)[1])
    assert rel_l2(*outs) <= 1e-12


@pytest.mark.parametrize("dkind", ["diag", "full"])
@pytest.mark.parametrize("vkind", [None, "scalar", "diag", "full"])
def test_table_kinds_two_components(G, dkind, vkind):
    """Every legal exp_D / exp_V shape combination for M = 2 (src/kernels.jl:9-15); `potential` is
    never exercised by the reference's own tests (SURVEY §4 coverage gaps)."""
    outs = []
    for ns in (G, O):
        pot = None
        if vkind == "scalar":
            pot = lambda r, p: 0.3 * (r[0] - 3.0) ** 2 + 0.1 * r[1] - 0.02j
        elif vkind == "diag":
            pot = lambda r, p, ns=ns: ns.SVector(0.3 * (r[0] - 3.0) ** 2, 0.2 * r[1] - 0.05j)
        elif vkind == "full":
            pot = lambda r, p, ns=ns: ns.SMatrix([[0.3 * r[0], 0.1 + 0.05j], [0.1 - 0.05j, 0.2 * r[1] - 0.03j]])
        pb = _free_problem(ns, (32, 64), (6.0, 6.0), np.complex128, seed=3, potential=pot, M=2, kind=dkind)
        prob = ns.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
        outs.append(ns.solve(prob, ns.StrangSplitting(), (0.0, 0.05), dt=0.01, nsaves=1)[1])
    assert rel_l2(*outs) <= 1e-12


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
@pytest.mark.parametrize("shape,kind", [((8, 16, 64), "full"), ((2, 16, 16, 32), "diag"), ((8, 4096), "full"), ((4096, 8), "full"),
                                        ((3, 64, 128), "scalar"), ((8192, 4), "diag"), ((4, 8192), "scalar")])
def test_two_components_geometries(G, shape, kind, dtype):
    """The component-parallel kernels (csrc/kernels_cp.cuh) over what the C3-shaped tests do not reach: 3-D grids (the
    forward-only / inverse-only passes of the middle axis), long lines with factorised twiddles on either axis, batch
    dims, a scalar (separable) dispersion shared by both components, and a Kerr + cross-Kerr nonlinearity with a
    per-component pump so that the exchange of the half-step matters."""
    nd = 3 if len(shape) >= 3 and shape[-3] > 3 and kind != "scalar" else 2
    if shape == (2, 16, 16, 32):
        nd = 3
    lengths = (7.0, 11.0, 5.0)[:nd]
    outs = []
    for ns in (G, O):
        pb = _free_problem(ns, shape, lengths, dtype, seed=5, M=2, kind=kind)
        pb["kwargs"]["nonlinearity"] = lambda u, p, ns=ns: ns.SVector(0.4 * ns.abs2(u[0]) + 0.2 * ns.abs2(u[1]), 0.3 * ns.abs2(u[1]) - 0.01j)
        # "full": the second component's pump does not follow the first one's envelope -> dense pump route (PW_DENSE);
        # otherwise S(r) a(t) (PW_DET)
        sep = 0.0 if kind == "full" else 1.0
        pb["kwargs"]["pump"] = lambda r, p, t, ns=ns: ns.SVector(0.3 * np.exp(-(r[0] - 3.0) ** 2) * (1 + t),
                                                                 (0.1 + 0 * r[0]) * (1 + sep * t))
        prob = ns.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
        outs.append(ns.solve(prob, ns.StrangSplitting(), (0.0, 0.03), dt=0.01, nsaves=1)[1])
    assert rel_l2(*outs) <= (1e-12 if dtype == np.complex128 else 2e-6)


def test_no_dispersion_and_complex_nonlinearity(G):
    """Absent dispersion (AdditiveIdentity -> no FFT at all) and a complex (lossy) nonlinearity."""
    outs = []
    for ns in (G, O):
        rng = np.random.default_rng(5)
        u0 = ((rng.standard_normal((64, 64)) + 1j * rng.standard_normal((64, 64))),)
        nl = lambda psi, p, ns=ns: (0.7 - 0.2j) * ns.abs2(psi[0]) + 0.1
        prob = ns.GrossPitaevskiiProblem(u0, (5.0, 5.0), nonlinearity=nl)
        outs.append(ns.solve(prob, ns.StrangSplitting(), (0.0, 0.05), dt=0.01, nsaves=1)[1])
    assert rel_l2(*outs) <= 1e-12


def test_error_behaviour(G):
    """Unsupported inputs fail loudly (no CPU fallback): an axis beyond the built transform sizes (any length up to
    4096 runs on the generic plan, tests/test_gpu_generic.py), unregistered closure."""
    u0 = (np.zeros((2, 4100), dtype=np.complex128),)
    disp = lambda ks, p: (ks[0] ** 2 + ks[1] ** 2) / 2
    prob = G.GrossPitaevskiiProblem(u0, (5.0, 5.0), dispersion=disp)
    with pytest.raises(G.GgpError):
        G.solve(prob, G.StrangSplitting(), (0.0, 0.05), dt=0.01, nsaves=1)
    u0 = (np.zeros((64, 64), dtype=np.complex128),)
    bad_nl = lambda psi, p: G.abs2(psi[0]) ** 2
    prob = G.GrossPitaevskiiProblem(u0, (5.0, 5.0), dispersion=disp, nonlinearity=bad_nl)
    with pytest.raises(G.UnsupportedForm):
        G.solve(prob, G.StrangSplitting(), (0.0, 0.05), dt=0.01, nsaves=1)
    with pytest.raises(AssertionError):   # src/problem.jl:112-113
        G.GrossPitaevskiiProblem((np.zeros((4, 4)), np.zeros((4, 8))), (1.0, 1.0))


def test_full_size_properties_c2(G):
    """BASELINE config C2 at its full size (2048^2, ComplexF32) through size-independent properties:
    norm conservation of the unitary Kerr flow, and time reversal (forward dt then backward dt)."""
    pb = P.kerr2d(G, N=2048, dtype=np.complex64, nsteps=20)
    prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    ts, sol = G.solve(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1)
    u0, u1 = sol[0][0].astype(np.complex128), sol[0][1].astype(np.complex128)
    assert abs(np.linalg.norm(u1) / np.linalg.norm(u0) - 1) < 2e-5
    assert np.linalg.norm(u1 - u0) / np.linalg.norm(u0) > 1e-3          # it did move
    # backward in time from u1: the Strang step is symmetric, so (-dt) undoes (+dt) up to rounding
    prob_b = G.GrossPitaevskiiProblem((sol[0][1],), pb["lengths"], **pb["kwargs"])
    _, solb = G.solve(prob_b, G.StrangSplitting(), (pb["tspan"][1], pb["tspan"][0]), dt=-pb["dt"], nsaves=1)
    ub = solb[0][1].astype(np.complex128)
    assert np.linalg.norm(ub - u0) / np.linalg.norm(u0) < 5e-5


def test_observables(G):
    pb = P.truncated_wigner(G, ntraj=8, N=64, ndim=2, tspan=(0, 0.1))
    prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    it = G.init(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1, save_start=False, rng=1)
    u = pb["u0"][0]
    dens = it.observe(G.lib.OBS_DENSITY)
    assert np.allclose(dens[0], (np.abs(u) ** 2).sum(0), rtol=1e-12)
    nk = it.observe(G.lib.OBS_MOMENTUM)
    ref = (np.abs(np.fft.fft2(u)) ** 2).sum(0) / u[0].size ** 2
    assert np.allclose(nk[0], ref, rtol=1e-10, atol=1e-12 * ref.max())
    nrm = it.observe(G.lib.OBS_NORM)
    assert np.isclose(nrm[0], (np.abs(u) ** 2).sum(), rtol=1e-12)
    # the state is untouched by observing
    assert np.array_equal(it.fetch()[0], u)
    it.close()


@pytest.mark.parametrize("ntraj,N,dtype", [(37, 64, np.complex128), (600, 256, np.complex128), (50, 32, np.complex64)])
def test_g2_momentum_observable(G, ntraj, N, dtype):
    """the trajectory sum of examples/truncated_wigner.jl:143-154 (`G2`) on the device, against the
    example's own host formulas -- raw second moment, then g2(k,k') with the Wigner-ordering corrections of `f`."""
    pb = P.truncated_wigner(G, ntraj=ntraj, N=N, ndim=1, dtype=dtype, tspan=(0, 0.5))
    prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    it = G.init(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1, save_start=False, rng=3)
    it.advance(4)
    u = np.array(it.fetch()[0]).astype(np.complex128)                    # (ntraj, N)
    raw = it.observe(G.lib.OBS_G2_MOMENTUM)[0]
    nk = it.observe(G.lib.OBS_MOMENTUM)[0]
    assert np.array_equal(np.array(it.fetch()[0]).astype(np.complex128), u)        # observing leaves the state alone
    it.close()
    inten = np.abs(np.fft.fft(u, axis=-1) / N) ** 2                        # |ft_sol|^2, ft_sol = fft(sol, 1) / N  (:110)
    ref_raw = inten.T @ inten
    tol = 1e-10 if dtype == np.complex128 else 2e-5
    assert np.allclose(raw, ref_raw, rtol=tol, atol=tol * ref_raw.max())
    assert np.allclose(nk, inten.sum(0), rtol=tol, atol=tol * inten.sum(0).max())
    # the example's G2 / g2 from the two device observables (commutator = 1/L, :139-156)
    Lbox = float(pb["lengths"][0])
    cm = 1.0 / Lbox
    delta = np.eye(N)
    G2_dev = (raw - (1 + delta) * cm / 2 * (nk[:, None] + nk[None, :] - ntraj * cm / 2)) / ntraj
    G2_ref = np.empty((N, N))
    for m in range(N):
        for n in range(N):
            G2_ref[m, n] = np.mean(inten[:, m] * inten[:, n] - (1 + (m == n)) * cm / 2 * (inten[:, m] + inten[:, n] - cm / 2))
    assert np.allclose(G2_dev, G2_ref, rtol=1e3 * tol, atol=1e3 * tol * np.abs(G2_ref).max())


@pytest.mark.parametrize("M,ntraj,N", [(1, 300, 64), (2, 33, 32)])
def test_windowed_correlation_observable(G, M, ntraj, N):
    """`correlation` of test/windowed_ft.jl:31-49 on the device (ggp_observe_windowed) against the
    host restatement used by the known-answer tests, on the same ensemble."""
    from test_oracle_known_answers import _window, windowed_correlation
    rng = np.random.default_rng(5)
    L = 20.0
    u0 = tuple((rng.standard_normal((ntraj, N)) + 1j * rng.standard_normal((ntraj, N))) for _ in range(M))
    prob = G.GrossPitaevskiiProblem(u0, (L,), dispersion=lambda ks, p: ks[0] ** 2 / 2)
    it = G.init(prob, G.StrangSplitting(), (0, 1.0), dt=0.1, nsaves=1, save_start=False)
    it.advance(3)
    state = [np.array(x) for x in it.fetch()]
    rs = -L / 2 + np.arange(N) * (L / N)
    for par1, par2 in (((-1, 3), (1, 3)), ((0, 5), (0, 4))):
        got = it.observe_windowed(_window(rs, par1), _window(rs, par2))
        for c in range(M):
            ref = windowed_correlation(state[c], rs, par1, par2) * state[c].size
            assert np.allclose(got[c], ref, rtol=1e-10, atol=1e-10 * np.abs(ref).max()), (c, par1, par2)
    assert all(np.array_equal(a, b) for a, b in zip(state, it.fetch()))       # the state is untouched
    it.close()

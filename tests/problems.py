"""Problem definitions shared by the oracle tests and the GPU parity tests.

Each factory takes `ns`, a namespace providing SVector / SMatrix / abs2 (either oracle/ggp_oracle.py
or the product's host module), and returns a dict with the arguments of
`GrossPitaevskiiProblem(u0, lengths; ...)` and of `solve(prob, StrangSplitting(), tspan; dt, nsaves)`.
The closures are the reference's own test / example closures (file:line cited per factory).
"""
from types import SimpleNamespace

import numpy as np


def _sumsq(ks):
    s = 0
    for k in ks:
        s = s + k * k
    return s


def quick_start(ns, kerr=True, N=128, dtype=np.complex128):
    """examples/quick_start.jl:16-17,35,48,63-70,90-101 (BASELINE config C1)."""
    L = 8
    dL = L / N
    rs = np.arange(N) * dL
    X, Y = np.meshgrid(rs, rs, indexing="xy")   # numpy (y, x): x fastest == Julia first index
    u0 = np.exp(-(X - L / 2) ** 2 - (Y - L / 2) ** 2).astype(dtype)

    def dispersion(ks, param):
        return _sumsq(ks) / 2

    def nonlinearity(u, param):
        return param.g * ns.abs2(u[0])

    kw = dict(dispersion=dispersion)
    tspan = (0, 1)
    if kerr:
        kw.update(nonlinearity=nonlinearity, param=SimpleNamespace(g=-6))
        tspan = (0, 0.4)
    return dict(u0=(u0,), lengths=(L, L), kwargs=kw, tspan=tspan, dt=0.01, nsaves=64)


def bistability(ns, wrap=(0, 0, 0), n=256, nsaves=512, tspan=(0, 3300), dt=0.05):
    """test/bistability_cycle.jl:1-55; wrap selects scalar / SVector{1} / SMatrix{1,1} wrappers
    for (dispersion, nonlinearity, pump) as in :28-35,67-71."""
    w0, g, delta, kz, gamma = 1483, 0.01, 0.3, 27, 0.1
    wp = w0 + delta
    L = 256
    Imax, width = 0.6, 50
    tmax = tspan[-1]
    param = SimpleNamespace(tmax=tmax, Imax=Imax, width=width, wp=wp, w0=w0, kz=kz, gamma=gamma, g=g, L=L)

    def dispersion(ks, p):
        return -1j * p.gamma / 2 + p.w0 * (1 + _sumsq(ks) / (2 * p.kz ** 2)) - p.wp

    def nonlinearity(psi, p):
        return p.g * ns.abs2(psi)

    def I(t, tmax, Imax):
        val = -Imax * t * (t - tmax) * 4 / tmax ** 2
        return 0.0 if val < 0 else val

    def pump(x, p, t):
        s = 0
        for xi in x:
            s = s + (xi - p.L / 2) ** 2
        return np.exp(-s / p.width ** 2) * np.sqrt(I(t, p.tmax, p.Imax))

    def wrapv(f):
        return lambda *a: ns.SVector(_first(f(*a)))

    def wrapm(f):
        return lambda *a: ns.SMatrix([[_first(f(*a))]])

    def _first(v):
        return v[0] if isinstance(v, ns.SVector) else v

    fs = []
    for f, w in zip((dispersion, nonlinearity, pump), wrap):
        fs.append(f if w == 0 else (wrapv(f) if w == 1 else wrapm(f)))
    u0 = (np.zeros(n, dtype=np.complex128),)
    return dict(u0=u0, lengths=(L,), kwargs=dict(dispersion=fs[0], nonlinearity=fs[1], pump=fs[2], param=param),
                tspan=tspan, dt=dt, nsaves=nsaves, I=I, delta=delta, g=g, gamma=gamma, Imax=Imax)


def exciton_polariton(ns, N=128, nsaves=256, tspan=(0, 100), dt=1e-1, dtype=np.complex128, time_pump=False):
    """test/exciton_polariton_test.jl:1-46."""
    hbar = 0.654
    Wr = 5.07 / (2 * hbar)
    gx = 0.0015 / hbar
    gc = 0.07 / 0.6571 / hbar
    wx = 1484.44 / hbar
    wc = 1482.76 / hbar
    m = hbar ** 2 / (2 * 2e-1)
    wp = wc
    dx_ = wp - wx
    dc_ = wp - wc
    A, w, g = 2, 100, 1e-2 / hbar
    L = 256
    param = SimpleNamespace(hbar=hbar, m=m, wc=wc, dc=dc_, gc=gc, dx=dx_, gx=gx, Wr=Wr, A=A, w=w, g=g, L=L,
                            tmax=tspan[-1])

    def dispersion(k, p):
        Dcc = p.hbar * _sumsq(k) / (2 * p.m) - p.dc - 1j * p.gc
        Dxx = -p.dx - 1j * p.gx
        Dxc = p.Wr
        return ns.SMatrix([[Dcc, Dxc], [Dxc, Dxx]])

    def nonlinearity(psi, p):
        return ns.SVector(0, p.g * ns.abs2(psi[1]))

    def pump(r, p, t):
        s = 0
        for ri in r:
            s = s + (ri - p.L / 2) ** 2
        amp = 1.0
        if time_pump:  # examples/bistability.jl:71-74 envelope, used by BASELINE config C3
            val = -t * (t - p.tmax) * 4 / p.tmax ** 2
            amp = np.sqrt(val) if val > 0 else 0.0
        return ns.SVector(p.A * np.exp(-s / p.w ** 2) * amp, 0)

    u0 = (np.zeros((N, N), dtype=dtype), np.zeros((N, N), dtype=dtype))
    return dict(u0=u0, lengths=(L, L), kwargs=dict(dispersion=dispersion, nonlinearity=nonlinearity, pump=pump,
                                                   param=param),
                tspan=tspan, dt=dt, nsaves=nsaves, param=param)


def windowed_ft(ns, ntraj=10 ** 4, N=64):
    """test/windowed_ft.jl:23-29,62-90."""
    L = 20
    dL = L / N
    hbar = 0.6582
    gamma = 0.1 / hbar
    m = hbar ** 2 / 2.5
    d0 = 0.49 / hbar
    dt = 4
    param = SimpleNamespace(d0=d0, m=m, gamma=gamma, hbar=hbar, L=L, dL=dL, N=N, dt=dt)

    def dispersion(ks, p):
        return -1j * p.gamma / 2 + p.hbar * _sumsq(ks) / (2 * p.m) - p.d0

    def position_noise_func(psi, r, p):
        return np.sqrt(p.gamma / 2 / p.dL)

    u0 = (np.zeros((ntraj, N), dtype=np.complex128),)
    noise_prototype = tuple(np.empty_like(x) for x in u0)
    return dict(u0=u0, lengths=(L,), kwargs=dict(dispersion=dispersion, param=param,
                                                 position_noise_func=position_noise_func,
                                                 noise_prototype=noise_prototype),
                tspan=(0, 200), dt=dt, nsaves=1, save_start=False, L=L, N=N)


def truncated_wigner(ns, ntraj=256, N=256, ndim=1, dtype=np.complex128, seed=1234, tspan=(0, 200), dt=0.05,
                     traj_range=None):
    """examples/truncated_wigner.jl:33-50,93-99 (1-D as shipped; ndim=2 is BASELINE config C4)."""
    hbar = 0.6582
    gamma = 0.047 / hbar
    m = 1 / 6
    g = 3e-4 / hbar
    delta = 0.49 / hbar
    A = 10
    L = 512
    dx = L / N
    vol = dx ** ndim
    param = SimpleNamespace(hbar=hbar, m=m, delta=delta, gamma=gamma, g=g, A=A, L=L, dx=vol)

    def dispersion(ks, p):
        return p.hbar * _sumsq(ks) / (2 * p.m) - p.delta - 1j * p.gamma / 2

    def pump(x, p, t):
        return p.A

    def nonlinearity(psi, p):
        return p.g * (ns.abs2(psi[0]) - 1 / p.dx)

    def position_noise_func(psi, xs, p):
        return np.sqrt(p.gamma / (2 * p.dx))

    # the ensemble is generated in blocks of 256 trajectories seeded by (seed, block), so that rank r of a
    # sharded run can build exactly its slice [lo, hi) of the same global ensemble
    lo, hi = traj_range if traj_range is not None else (0, ntraj)
    B = 256
    parts = []
    for b in range(lo // B, (hi + B - 1) // B):
        rng = np.random.default_rng([seed, b])
        shape = (B,) + (N,) * ndim
        z = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)
        a, e = max(lo, b * B) - b * B, min(hi, (b + 1) * B) - b * B
        parts.append((z[a:e] / np.sqrt(2 * vol)).astype(dtype))
    u0 = (np.concatenate(parts, axis=0),)
    noise_prototype = tuple(np.empty_like(x) for x in u0)
    return dict(u0=u0, lengths=(L,) * ndim,
                kwargs=dict(dispersion=dispersion, nonlinearity=nonlinearity, pump=pump, param=param,
                            noise_prototype=noise_prototype, position_noise_func=position_noise_func),
                tspan=tspan, dt=dt, nsaves=1, save_start=False, param=param)


def kerr2d(ns, N=256, dtype=np.complex64, L=64.0, g=1.0, dt=1e-3, nsteps=100, seed=1234):
    """BASELINE config C2 shape: 2-D scalar Kerr, D=|k|²/2, G=g|u|², Gaussian + 10 % noise."""
    real = np.float32 if dtype == np.complex64 else np.float64
    Lr = real(L)
    rs = np.arange(N).astype(real) * (Lr / N)
    X, Y = np.meshgrid(rs, rs, indexing="xy")
    rng = np.random.default_rng(seed)
    xi = (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))) / np.sqrt(2)
    u0 = (np.exp(-((X - Lr / 2) ** 2 + (Y - Lr / 2) ** 2) / 16) * (1 + 0.1 * xi)).astype(dtype)

    def dispersion(ks, param):
        return _sumsq(ks) / 2

    def nonlinearity(u, param):
        return param.g * ns.abs2(u[0])

    dtr = real(dt)
    return dict(u0=(u0,), lengths=(Lr, Lr),
                kwargs=dict(dispersion=dispersion, nonlinearity=nonlinearity, param=SimpleNamespace(g=real(g))),
                tspan=(real(0), real(nsteps) * dtr), dt=dtr, nsaves=1)


def kerr3d(ns, N=32, dtype=np.complex64, L=32.0, g=1.0, dt=1e-3, nsteps=10, seed=1234):
    """BASELINE config C5 shape: 3-D Kerr GPE."""
    real = np.float32 if dtype == np.complex64 else np.float64
    Lr = real(L)
    rs = np.arange(N).astype(real) * (Lr / N)
    Z, Y, X = np.meshgrid(rs, rs, rs, indexing="ij")
    rng = np.random.default_rng(seed)
    xi = (rng.standard_normal((N, N, N)) + 1j * rng.standard_normal((N, N, N))) / np.sqrt(2)
    u0 = (np.exp(-((X - Lr / 2) ** 2 + (Y - Lr / 2) ** 2 + (Z - Lr / 2) ** 2) / 16) * (1 + 0.01 * xi)).astype(dtype)

    def dispersion(ks, param):
        return _sumsq(ks) / 2

    def nonlinearity(u, param):
        return param.g * ns.abs2(u[0])

    dtr = real(dt)
    return dict(u0=(u0,), lengths=(Lr, Lr, Lr),
                kwargs=dict(dispersion=dispersion, nonlinearity=nonlinearity, param=SimpleNamespace(g=real(g))),
                tspan=(real(0), real(nsteps) * dtr), dt=dtr, nsaves=1)


def kerr3d_slab(ns, N=512, rank=0, world=1, dtype=np.complex64, L=32.0, g=1.0, dt=1e-3, nsteps=100, seed=1234):
    """BASELINE config C5 built slab by slab: rank's z-planes [rank*N/world, (rank+1)*N/world) of the
    N^3 Gaussian + 1 % noise field; the noise is seeded per z-plane so every sharding sees the same global field."""
    real = np.float32 if dtype == np.complex64 else np.float64
    Lr = real(L)
    rs = np.arange(N).astype(real) * (Lr / N)
    nz = N // world
    z0 = rank * nz
    Y, X = np.meshgrid(rs, rs, indexing="ij")
    gxy = np.exp(-((X - Lr / 2) ** 2 + (Y - Lr / 2) ** 2) / 16).astype(real)
    u0 = np.empty((nz, N, N), dtype=dtype)

    def plane(k):
        rng = np.random.default_rng([seed, z0 + k])
        xi = rng.standard_normal((N, N, 2), dtype=np.float32)
        gz = real(np.exp(-((rs[z0 + k] - Lr / 2) ** 2) / 16))
        u0[k] = (gxy * gz) * (1 + real(0.01 / np.sqrt(2)) * (xi[..., 0] + 1j * xi[..., 1]))

    if nz * N * N >= 1 << 24:       # large slabs: planes are independent (one generator each), NumPy releases the GIL
        import os
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max(1, min(16, len(os.sched_getaffinity(0))))) as ex:
            list(ex.map(plane, range(nz)))
    else:
        for k in range(nz):
            plane(k)

    def dispersion(ks, param):
        return _sumsq(ks) / 2

    def nonlinearity(u, param):
        return param.g * ns.abs2(u[0])

    dtr = real(dt)
    return dict(u0=(u0,), lengths=(Lr, Lr, Lr),
                kwargs=dict(dispersion=dispersion, nonlinearity=nonlinearity, param=SimpleNamespace(g=real(g))),
                tspan=(real(0), real(nsteps) * dtr), dt=dtr, nsaves=1)


def noise_forms(ns, form="field", ndim=1, M=1, N=32, ntraj=3, dtype=np.complex128, real_proto=False):
    """Noise amplitudes of docs/src/stochastic_simulations.md:62-86 beyond the constant one:
    `field`  alpha*abs(u[1]) + const,  `profile`  beta*exp(-sum(abs2, r)/sigma2),  `both`  profile x (const + |u|)
    with an SVector return for M = 2.  Damped free dispersion + Kerr term so that every stage of the half-step runs."""
    L = 10.0
    rng = np.random.default_rng(42)
    shape = (ntraj,) + (N,) * ndim
    u0 = tuple(((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / 2).astype(dtype) for _ in range(M))
    param = SimpleNamespace(alpha=0.3, beta=0.7, sigma2=30.0, c=0.2, g=0.5, gamma=0.1)

    def dispersion(ks, p):
        return _sumsq(ks) / 2 - 1j * p.gamma / 2

    def nonlinearity(u, p):
        return p.g * ns.abs2(u[0])

    def prof(r, p):
        s = 0
        for ri in r:
            s = s + ri * ri
        return p.beta * np.exp(-s / p.sigma2)

    if form == "field":
        def eta(u, r, p):
            return p.c + p.alpha * abs(u[0])
    elif form == "profile":
        def eta(u, r, p):
            return prof(r, p)
    else:
        def eta(u, r, p):
            if M == 1:
                return prof(r, p) * (p.c + p.alpha * abs(u[0]))
            return ns.SVector(prof(r, p) * (p.c + p.alpha * abs(u[1])), prof(r, p) * (2 * p.c + 0.5 * p.alpha * abs(u[0])))
    real_t = np.float32 if dtype == np.complex64 else np.float64
    proto = tuple(np.empty(x.shape, dtype=real_t if real_proto else dtype) for x in u0)
    return dict(u0=u0, lengths=(L,) * ndim,
                kwargs=dict(dispersion=dispersion, nonlinearity=nonlinearity, param=param,
                            noise_prototype=proto, position_noise_func=eta),
                tspan=(0, 0.5), dt=0.05, nsaves=2, save_start=True)


def generic(ns, case="np2_1d", dtype=np.complex128, nsteps=6):
    """Problems that run on the generic plan (generic_plan.cuh): FFT axes of any length (the reference plans FFTW
    transforms of whatever size u0 has, src/misc.jl:53-58), more than two components (NTuple{M}, src/kernels.jl:31-35)
    and SMatrix-valued nonlinearities (src/kernels.jl:22-25,44).  sizes are Julia order (n1 fastest); NumPy arrays are
    (batch, ..., n2, n1)."""
    real = np.float32 if dtype == np.complex64 else np.float64
    rng = np.random.default_rng(7)
    cases = {
        # case: (sizes, M, batch)
        "np2_1d": ((100,), 1, ()), "prime_1d": ((127,), 1, (3,)), "np2_2d": ((96, 80), 1, ()),
        "np2_mixed": ((128, 100), 1, (2,)), "np2_3d": ((24, 20, 18), 1, ()), "np2_long": ((1500,), 1, ()),
        "m3_vec": ((32, 32), 3, ()), "m3_matdisp": ((64,), 3, (2,)), "m4_mat": ((48,), 4, ()),
        "rabi": ((32, 32), 2, ()), "rabi_vs": ((32, 32), 2, ()), "rabi_vv": ((32, 32), 2, ()), "rabi_vm": ((32, 32), 2, ()),
        "m3_np2_noise": ((30, 28), 3, (2,)),
    }
    sizes, M, batch = cases[case]
    nd = len(sizes)
    L = tuple(real(6.0 + 2 * a) for a in range(nd))
    shape = batch + tuple(reversed(sizes))
    axes = [np.arange(n).astype(real) * (L[a] / n) for a, n in enumerate(sizes)]
    mesh = np.meshgrid(*reversed(axes), indexing="ij")          # slowest axis first
    r2 = sum((m - L[nd - 1 - i] / 2) ** 2 for i, m in enumerate(mesh))
    u0 = tuple(((np.exp(-r2 / (2.0 + c)) * (1 + 0.05 * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))))
                * np.exp(0.3j * c)).astype(dtype) for c in range(M))
    p = SimpleNamespace(g=real(0.8), om=real(0.6), gamma=real(0.05), v=real(0.3))
    kw = dict(param=p)

    def disp_scalar(ks, p):
        return _sumsq(ks) / 2 - 1j * p.gamma / 2

    def pot_scalar(rs, p):
        return p.v * _sumsq(rs) / 10

    def pump_static(rs, p, t):
        return 0.4 * np.exp(-_sumsq(rs) / 8)

    if case in ("np2_1d", "np2_long"):
        kw.update(dispersion=disp_scalar, potential=pot_scalar, nonlinearity=lambda u, p: p.g * ns.abs2(u[0]))
    elif case == "prime_1d":
        kw.update(dispersion=disp_scalar, nonlinearity=lambda u, p: p.g * ns.abs2(u[0]),
                  noise_prototype=tuple(np.empty(x.shape, dtype=dtype) for x in u0),
                  position_noise_func=lambda u, r, p: 0.2 + 0.1 * abs(u[0]))
    elif case in ("np2_2d", "np2_mixed"):
        kw.update(dispersion=disp_scalar, nonlinearity=lambda u, p: p.g * (ns.abs2(u[0]) - 0.1j), pump=pump_static)
    elif case == "np2_3d":
        kw.update(dispersion=disp_scalar, potential=pot_scalar, nonlinearity=lambda u, p: p.g * ns.abs2(u[0]))
    elif case == "m3_vec":
        kw.update(dispersion=lambda ks, p: ns.SVector(_sumsq(ks) / 2, _sumsq(ks) / 3 - 1j * p.gamma, 0.2 + _sumsq(ks) / 4),
                  potential=lambda rs, p: ns.SVector(p.v * _sumsq(rs) / 10, 0.1, -p.v * rs[0]),
                  nonlinearity=lambda u, p: ns.SVector(p.g * ns.abs2(u[0]) + 0.5 * ns.abs2(u[2]), 0.3 * ns.abs2(u[1]) - 0.02j,
                                                       p.g * (ns.abs2(u[0]) + ns.abs2(u[1]) + ns.abs2(u[2]))),
                  pump=lambda rs, p, t: ns.SVector(0.4 * np.exp(-_sumsq(rs) / 8) * (1 + 0.5 * t), 0 * rs[0], 0.1 + 0 * rs[0]))
    elif case == "m3_matdisp":
        kw.update(dispersion=lambda ks, p: ns.SMatrix([[_sumsq(ks) / 2, p.om, 0 * ks[0]],
                                                       [p.om, _sumsq(ks) / 3 - 1j * p.gamma, 0.5 * p.om],
                                                       [0 * ks[0], 0.5 * p.om, 0.1 + 0 * ks[0]]]),
                  nonlinearity=lambda u, p: p.g * (ns.abs2(u[0]) + ns.abs2(u[2])))
    elif case == "m4_mat":
        def nl4(u, p):
            z = 0 * ns.abs2(u[0])
            return ns.SMatrix([[p.g * ns.abs2(u[0]), p.om + z, z, z],
                               [p.om + z, p.g * ns.abs2(u[1]) - 0.05j, 0.3 * p.om + z, z],
                               [z, 0.3 * p.om + z, 0.5 * ns.abs2(u[3]), 0.2j + z],
                               [z, z, -0.2j + z, p.g * ns.abs2(u[2])]])
        kw.update(dispersion=disp_scalar, nonlinearity=nl4)
    elif case.startswith("rabi"):
        def nl2(u, p):
            z = 0 * ns.abs2(u[0])
            return ns.SMatrix([[p.g * ns.abs2(u[0]) + 0.2 * ns.abs2(u[1]), p.om + z],
                               [p.om + z, p.g * ns.abs2(u[1]) - 0.03j]])
        kw.update(dispersion=disp_scalar, nonlinearity=nl2, pump=pump_static)
        if case == "rabi_vs":
            kw.update(potential=pot_scalar)
        elif case == "rabi_vv":     # SMatrix nonlinearity x SVector potential: `_mul` gives the mat-vec product (quirk)
            kw.update(potential=lambda rs, p: ns.SVector(p.v * _sumsq(rs) / 10, 0.2 + 0 * rs[0]))
        elif case == "rabi_vm":
            kw.update(potential=lambda rs, p: ns.SMatrix([[p.v * _sumsq(rs) / 10, 0.1 + 0 * rs[0]],
                                                          [0.1 + 0 * rs[0], 0 * rs[0]]]))
    elif case == "m3_np2_noise":
        kw.update(dispersion=disp_scalar,
                  nonlinearity=lambda u, p: ns.SVector(p.g * ns.abs2(u[0]), p.g * ns.abs2(u[1]), 0.1 * ns.abs2(u[2])),
                  noise_prototype=tuple(np.empty(x.shape, dtype=dtype) for x in u0),
                  position_noise_func=lambda u, r, p: ns.SVector(0.2 + 0.1 * abs(u[1]), 0.1, 0.05 * abs(u[0])))
    dtr = real(0.01)
    return dict(u0=u0, lengths=L, kwargs=kw, tspan=(real(0), real(nsteps) * dtr), dt=dtr, nsaves=2)


def fuzz(ns, seed=0):
    """A random legal problem (deterministic in `seed`, identical for every namespace): dimensions, power-of-two and
    other axis lengths, batch dims, 1-3 components, every table kind, Number / SVector / SMatrix nonlinearities, static /
    separable / non-separable pumps, constant and field-dependent noise.  Used by tests/test_gpu_fuzz.py to reach
    combinations nobody wrote a dedicated test for."""
    rng = np.random.default_rng([20260, seed])
    pick = lambda xs: xs[int(rng.integers(len(xs)))]
    nd = pick([1, 1, 2, 2, 3])
    M = pick([1, 1, 2, 2, 3])
    dtype = pick([np.complex128, np.complex128, np.complex64])
    real = np.float32 if dtype == np.complex64 else np.float64
    pools = {1: [8, 16, 64, 256, 1024, 5, 12, 30, 100, 243], 2: [4, 8, 16, 32, 64, 6, 12, 20, 48], 3: [4, 8, 16, 3, 6, 10]}
    sizes = tuple(pick(pools[nd]) for _ in range(nd))
    batch = pick([(), (), (2,), (3,)])
    L = tuple(real(5.0 + 1.5 * a) for a in range(nd))
    shape = batch + tuple(reversed(sizes))
    u0 = tuple(((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * 0.6).astype(dtype) for _ in range(M))
    c = lambda: real(rng.uniform(0.1, 0.9))
    kw = {}
    # dispersion
    dk = pick(["none", "scalar", "scalar", "diag", "full"]) if M > 1 else pick(["none", "scalar", "scalar", "scalar"])
    d0, d1, d2, dg = c(), c(), c(), c() * real(0.1)
    if dk == "scalar":
        kw["dispersion"] = lambda ks, p: d0 * _sumsq(ks) - 1j * dg
    elif dk == "diag":
        kw["dispersion"] = lambda ks, p: ns.SVector(*[(d0 + real(0.2) * i) * _sumsq(ks) - 1j * dg * i for i in range(M)])
    elif dk == "full":
        kw["dispersion"] = lambda ks, p: ns.SMatrix([[((d0 + real(0.2) * i) * _sumsq(ks) - 1j * dg) if i == j else d1 * real(0.5) + 0 * ks[0]
                                                      for j in range(M)] for i in range(M)])
    # nonlinearity (decides which potential kinds are legal: SVector x SMatrix is a DimensionMismatch, src/kernels.jl:9)
    nk = pick(["none", "number", "vector", "matrix"]) if M > 1 else pick(["none", "number", "number", "vector"])
    g0, g1, gl = c(), c(), c() * real(0.05)
    if nk == "number":
        kw["nonlinearity"] = lambda u, p: g0 * sum(ns.abs2(u[i]) for i in range(M)) - 1j * gl
    elif nk == "vector":
        kw["nonlinearity"] = lambda u, p: ns.SVector(*[g0 * ns.abs2(u[i]) + g1 * ns.abs2(u[(i + 1) % M]) - 1j * gl * i
                                                       for i in range(M)])
    elif nk == "matrix":
        kw["nonlinearity"] = lambda u, p: ns.SMatrix([[(g0 * ns.abs2(u[i]) - 1j * gl) if i == j else g1 + 0 * ns.abs2(u[0])
                                                       for j in range(M)] for i in range(M)])
    vk_pool = ["none", "scalar", "diag", "full"] if M > 1 else ["none", "scalar", "scalar"]
    if nk == "vector" and M > 1:
        vk_pool = ["none", "scalar", "diag"]
    vk = pick(vk_pool)
    v0, v1 = c(), c()
    if vk == "scalar":
        kw["potential"] = lambda rs, p: v0 * _sumsq(rs) / 10 - 1j * real(0.01)
    elif vk == "diag":
        kw["potential"] = lambda rs, p: ns.SVector(*[(v0 + real(0.1) * i) * rs[0] - 1j * real(0.01) * i for i in range(M)])
    elif vk == "full":
        kw["potential"] = lambda rs, p: ns.SMatrix([[(v0 * rs[0] if i == j else v1 * real(0.3) + 0 * rs[0]) for j in range(M)]
                                                    for i in range(M)])
    pk = pick(["none", "static", "separable", "dense", "number"])
    a0 = c()
    if pk == "static":
        kw["pump"] = lambda rs, p, t: ns.SVector(*[a0 * np.exp(-_sumsq(rs) / 6) * (i + 1) for i in range(M)]) if M > 1 \
            else a0 * np.exp(-_sumsq(rs) / 6)
    elif pk == "separable":
        kw["pump"] = lambda rs, p, t: ns.SVector(*[a0 * np.exp(-_sumsq(rs) / 6) * (i + 1) * (1 + 3 * t) for i in range(M)]) if M > 1 \
            else a0 * np.exp(-_sumsq(rs) / 6) * (1 + 3 * t)
    elif pk == "dense":
        kw["pump"] = lambda rs, p, t: ns.SVector(*[a0 * np.exp(-(rs[0] - 2 - 10 * t * (i + 1)) ** 2) for i in range(M)]) if M > 1 \
            else a0 * np.exp(-(rs[0] - 2 - 10 * t) ** 2)
    elif pk == "number":
        kw["pump"] = lambda rs, p, t: a0 * (1 + t) + 0 * rs[0]
    qk = pick(["none", "none", "const", "field"])
    e0, e1 = c() * real(0.3), c() * real(0.2)
    noisy = qk != "none"
    if noisy:
        real_proto = bool(rng.integers(2))
        kw["noise_prototype"] = tuple(np.empty(x.shape, dtype=real if real_proto else dtype) for x in u0)
        if qk == "const":
            kw["position_noise_func"] = (lambda u, r, p: ns.SVector(*[e0 * (i + 1) for i in range(M)])) if M > 1 else (lambda u, r, p: e0)
        else:
            kw["position_noise_func"] = (lambda u, r, p: ns.SVector(*[e0 + e1 * abs(u[(i + 1) % M]) for i in range(M)])) if M > 1 \
                else (lambda u, r, p: e0 + e1 * abs(u[0]))
    dtr = real(0.01)
    desc = f"nd={nd} sizes={sizes} batch={batch} M={M} {np.dtype(dtype).name} D={dk} G={nk} V={vk} F={pk} noise={qk}"
    return dict(u0=u0, lengths=L, kwargs=kw, tspan=(real(0), real(4) * dtr), dt=dtr, nsaves=2, noisy=noisy, desc=desc)

"""Host side of the B200 backend: the reference's user surface
(`GrossPitaevskiiProblem`, `StrangSplitting`, `init`, `step!`, `solve!`, `solve`) above the C ABI.

It mirrors, in Python, exactly what the Julia shim in `julia/` does (there is no Julia in this
image): everything that involves the user's closures stays on the host --

  * grids                              src/problem.jl:129-139
  * resolve_fixed_timestepping         src/fixed_time_stepping.jl:14-24
  * exp tables via get_exponential     src/misc.jl:12-20   (arbitrary closures, exact parity)
  * closure recognition: nonlinearity -> c_i + sum_j g_ij |u_j|^2, pump -> S(r) a(t), noise -> const
    (SURVEY §8a registered forms); an unrecognised closure raises -- there is NO CPU fallback
  * the pump-amplitude schedule at the reference's (one-dt-late) times, SURVEY Q1
  * `ts` accumulation and the save loop  src/fixed_time_stepping.jl:38-50

and every grid-sized floating-point operation of the time loop runs in libggp.so.

Array convention (as in oracle/): Julia (n1, ..., nd, batch...) column-major == NumPy C-order
(batch..., nd, ..., n1).  Closures get coordinates in Julia order (`ks[0]` along n1).
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import _lib as L

# ------------------------------------------------------------------------------------------------
# StaticArrays stand-ins and the identity singletons of src/kernels.jl:1-7
# ------------------------------------------------------------------------------------------------


class SVector:
    def __init__(self, *items):
        if len(items) == 1 and isinstance(items[0], (list, tuple)):
            items = tuple(items[0])
        self.items = list(items)

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]

    def __iter__(self):
        return iter(self.items)

    def _map(self, o, op):
        if isinstance(o, SVector):
            return SVector([op(a, b) for a, b in zip(self.items, o.items)])
        return SVector([op(a, o) for a in self.items])

    def __mul__(self, o):
        return self._map(o, lambda a, b: a * b)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._map(o, lambda a, b: a / b)

    def __add__(self, o):
        return self._map(o, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, o):
        return self._map(o, lambda a, b: a - b)

    def __neg__(self):
        return SVector([-a for a in self.items])


class SMatrix:
    def __init__(self, rows):
        self.rows = [list(r) for r in rows]
        assert all(len(r) == len(self.rows) for r in self.rows), "square matrices only"

    @property
    def n(self):
        return len(self.rows)

    def __getitem__(self, ij):
        return self.rows[ij[0]][ij[1]]


def abs2(x):
    if isinstance(x, SVector):
        return SVector([abs2(a) for a in x.items])
    if isinstance(x, (tuple, list)):
        return SVector([abs2(a) for a in x])
    x = np.asarray(x)
    return x.real * x.real + x.imag * x.imag


class _Identity:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        return self

    def __repr__(self):
        return self._name


additiveIdentity = _Identity("additiveIdentity")
multiplicativeIdentity = _Identity("multiplicativeIdentity")


def _absent(f):
    return f is None or f is additiveIdentity


class UnsupportedForm(ValueError):
    """A closure that is legal in the reference but outside the registered device forms."""


# ------------------------------------------------------------------------------------------------
# problem, grids, time stepping
# ------------------------------------------------------------------------------------------------


def _jl(x):
    """Julia literal semantics for bare Python numbers (Float64 / Int are strong types)."""
    if isinstance(x, (np.floating, np.integer)):
        return x
    if isinstance(x, (int,)):
        return int(x)
    return np.float64(x)


class GrossPitaevskiiProblem:
    """src/problem.jl:97-122."""

    def __init__(self, u0, lengths, *, dispersion=additiveIdentity, potential=additiveIdentity,
                 nonlinearity=additiveIdentity, pump=additiveIdentity, position_noise_func=additiveIdentity,
                 momentum_noise_func=additiveIdentity, noise_prototype=additiveIdentity, param=None):
        u0 = tuple(np.asarray(x) for x in u0)
        lengths = tuple(lengths)
        assert all(x.ndim >= len(lengths) for x in u0)      # src/problem.jl:112
        assert all(x.shape == u0[0].shape for x in u0)      # src/problem.jl:113

        def cplx(x):                                        # complex.(u0), src/problem.jl:114
            if np.iscomplexobj(x):
                return x
            return x.astype(np.complex64 if x.dtype == np.float32 else np.complex128)

        self.u0 = tuple(cplx(x) for x in u0)
        ls = [_jl(l) for l in lengths]
        lt = np.result_type(*[np.asarray(l).dtype for l in ls])
        self.lengths = tuple(lt.type(l) for l in ls)        # promote(lengths...), :115
        self.dispersion, self.potential = dispersion, potential
        self.nonlinearity, self.pump = nonlinearity, pump
        self.position_noise_func, self.momentum_noise_func = position_noise_func, momentum_noise_func
        self.noise_prototype = noise_prototype
        self.param = param

    def __repr__(self):
        return f"{len(self.lengths)}D GrossPitaevskiiProblem"

    @property
    def ndim(self):
        return len(self.lengths)

    @property
    def sizes(self):
        s = self.u0[0].shape
        return tuple(reversed(s[len(s) - self.ndim:]))      # (n1, ..., nd)


def direct_grid(prob):
    """src/problem.jl:129-133: x_j = (j-1) L/N."""
    out = []
    for Lk, n in zip(prob.lengths, prob.sizes):
        step = Lk / n if isinstance(Lk, np.floating) else np.float64(Lk) / n
        out.append(np.arange(n).astype(np.asarray(step).dtype) * step)
    return tuple(out)


def reciprocal_grid(prob):
    """src/problem.jl:135-139 with AbstractFFTs.fftfreq (negative Nyquist bin)."""
    out = []
    for Lk, n in zip(prob.lengths, prob.sizes):
        ratio = n / Lk if isinstance(Lk, np.floating) else np.float64(n) / np.float64(Lk)
        ft = np.asarray(ratio).dtype
        fs = ft.type(2 * math.pi) * ft.type(n) / ft.type(Lk)
        i = np.arange(n)
        idx = np.where(i < ((n + 1) >> 1), i, i - n).astype(ft)
        out.append(idx * (fs / ft.type(n)))
    return tuple(out)


def _mesh(axes):
    d = len(axes)
    pts = []
    for m, ax in enumerate(axes):
        shape = [1] * d
        shape[d - 1 - m] = len(ax)
        pts.append(ax.reshape(shape))
    return tuple(pts)


def resolve_fixed_timestepping(dt, tspan, nsaves):
    """src/fixed_time_stepping.jl:14-24 (dt is rewritten)."""
    dt = _jl(dt)
    t0, t1 = _jl(tspan[0]), _jl(tspan[-1])
    T = np.result_type(*[np.asarray(v).dtype for v in (dt, t0, t1)])
    if not np.issubdtype(T, np.floating):
        T = np.dtype(np.float64)
    ts = np.empty(nsaves + 1, dtype=T)
    ts[0] = t0
    dT = T.type(T.type(t1) - T.type(t0)) / nsaves
    steps_per_save = int(math.ceil(dT / dt))
    return dT / steps_per_save, ts, steps_per_save


class StrangSplitting:
    """src/strang_splitting.jl:7."""


# ------------------------------------------------------------------------------------------------
# tables: get_exponential (src/misc.jl:12-20), evaluated on the host with the user's closure
# ------------------------------------------------------------------------------------------------


def _cis(z):
    z = np.asarray(z)
    if np.iscomplexobj(z):
        return np.exp(-z.imag) * (np.cos(z.real) + 1j * np.sin(z.real))
    return np.cos(z) + 1j * np.sin(z)


def _expm2(a, c, b, d):
    """exp of [[a, b], [c, d]] per grid point: the closed form StaticArrays uses for 2x2
    (sqrt of the discriminant, two expm1, series branch for a vanishing discriminant)."""
    a, b, c, d = np.broadcast_arrays(*(np.asarray(v, dtype=np.result_type(a, b, c, d, np.complex64))
                                       for v in (a, b, c, d)))
    z = np.sqrt((a - d) * (a - d) + 4 * b * c)
    e = np.expm1((a + d - z) / 2)
    f = np.expm1((a + d + z) / 2)
    eps = np.finfo(z.real.dtype).eps
    tiny = (z.real ** 2 + z.imag ** 2) < eps * eps
    g = np.where(tiny, np.exp((a + d) / 2) * (1 + z * z / 24), (f - e) / np.where(tiny, 1, z))
    return (g * (a - d) + f + e) / 2 + 1, g * c, g * b, (-g * (a - d) + f + e) / 2 + 1  # m11 m21 m12 m22


def _expm_batched(A):
    """exp of a stack of small matrices (..., M, M), M > 2: scaling and squaring with a Pade approximant
    (scipy.linalg.expm, the algorithm family StaticArrays / LinearAlgebra use for sizes beyond 2x2)."""
    import scipy.linalg
    A = np.asarray(A, dtype=complex)
    flat = A.reshape((-1,) + A.shape[-2:])
    # identical points (constant couplings on a large grid) are exponentiated once
    uniq, inv = np.unique(flat.reshape(flat.shape[0], -1), axis=0, return_inverse=True)
    out = np.stack([scipy.linalg.expm(m.reshape(A.shape[-2:])) for m in uniq])
    return out[np.asarray(inv).reshape(-1)].reshape(A.shape)


def exp_table(f, grid, param, dt, M):
    """Returns (kind, AoS array of shape (npoints, ncols) complex) for cis(-dt * f(point, param))."""
    if _absent(f):
        return L.TABLE_NONE, None
    shape = tuple(len(g) for g in reversed(grid))
    val = f(_mesh(grid), param)
    if isinstance(val, SMatrix) and val.n == 1:
        val = val[0, 0] if M == 1 else val
    if isinstance(val, SVector) and len(val) == 1 and M == 1:
        val = val[0]
    if isinstance(val, SVector):
        if len(val) != M:
            raise ValueError("SVector table length must equal the number of components")
        cols = [np.broadcast_to(_cis(-dt * np.asarray(v)), shape) for v in val]
        kind = L.TABLE_DIAG
    elif isinstance(val, SMatrix):
        if val.n != M:
            raise ValueError("SMatrix table must be M x M")
        if M == 2:
            m11, m21, m12, m22 = _expm2(*(1j * (-dt * np.asarray(val[i, j])) for (i, j) in ((0, 0), (1, 0), (0, 1), (1, 1))))
            cols = [np.broadcast_to(m, shape) for m in (m11, m21, m12, m22)]  # column-major like SMatrix
        else:
            # -dt * D is formed in the closure's own precision first (Float32 dt x Float32 D in a ComplexF32 problem,
            # as `-δt * f(x, param)` does in src/misc.jl:15); only the exponential runs in double
            E = _expm_batched(np.stack([np.stack([np.broadcast_to(np.asarray(1j * (-dt * np.asarray(val[i, j])), dtype=complex), shape)
                                                  for j in range(M)], axis=-1) for i in range(M)], axis=-2))
            cols = [E[..., i, j] for j in range(M) for i in range(M)]          # column-major: (1,1) (2,1) ... (M,M)
        kind = L.TABLE_FULL
    else:
        cols = [np.broadcast_to(_cis(-dt * np.asarray(val)), shape)]
        kind = L.TABLE_SCALAR
    table = np.stack([np.asarray(c).reshape(-1) for c in cols], axis=1)
    return kind, np.ascontiguousarray(table)


def separable_dispersion_tol(f, grid, param, table, rng=None):
    """`disp_sep_tol` of include/ggp.h for a scalar dispersion table of a ComplexF32 problem: if D(k) is a sum over
    axes in Float64 arithmetic (checked on random grid points: the mixed second difference
    D(k) - sum_a D(k_a e_a) + (d-1) D(0) vanishes), return the deviation of the given table from the product of its
    own axis factors (its rounding, eps32 * |phase|) with a 25 % margin; otherwise 0 (library default)."""
    d = len(grid)
    if d < 2:
        return 0.0
    rng = rng or np.random.default_rng(0x5E9)
    g64 = [np.asarray(g, dtype=np.float64) for g in grid]
    idx = [rng.integers(0, len(g), size=2048) for g in g64]
    zero = [np.full(2048, g[0]) for g in g64]           # k = 0 is the first entry of every fftfreq axis
    pts = [g[i] for g, i in zip(g64, idx)]

    def D(p):
        v = f(tuple(p), param)
        if isinstance(v, (SVector, SMatrix)):
            return None
        return np.asarray(v, dtype=np.complex128) + np.zeros(2048)

    full, origin = D(pts), D(zero)
    if full is None or origin is None:
        return 0.0
    acc = full + (d - 1) * origin
    for a in range(d):
        acc = acc - D([pts[b] if b == a else zero[b] for b in range(d)])
    scale = max(1e-300, float(np.abs(full).max()))
    if float(np.abs(acc).max()) > 1e-12 * scale:
        return 0.0
    # deviation of the table from perp (x) line, where line runs along the LAST axis (the strided kernel's)
    shape = tuple(len(g) for g in reversed(grid))           # (n_d, ..., n_1)
    tab = np.asarray(table).reshape(shape[0], -1)
    d0 = tab[0, 0]
    if d0 == 0 or not np.isfinite(d0):
        return 0.0
    line, perp = tab[:, 0] / d0, tab[0, :]
    dev, dmax = 0.0, 0.0
    for l0 in range(0, shape[0], 256):                      # chunked: the table may be hundreds of MB
        blk = tab[l0:l0 + 256]
        dev = max(dev, float(np.abs(blk - line[l0:l0 + 256, None] * perp[None, :]).max()))
        dmax = max(dmax, float(np.abs(blk).max()))
    return 1.25 * dev / dmax + 1e-9 if dmax > 0 else 0.0


def dispersion_axis_factors(f, grid, param, dt):
    """GGP_TABLE_SEP_AXES (include/ggp.h): if the scalar dispersion closure is a sum over axes -- verified in Float64 on
    random grid points, as in `separable_dispersion_tol` -- return the d per-axis factors of
    exp_D(k) = cis(-dt D(k)) = prod_a cis(-dt D_a(k_a)),  D_1(k_1) = D(k_1, 0, ...),  D_a(k_a) = D(0, .., k_a, ..) - D(0),
    each evaluated in Float64 on the axis' own grid values; otherwise None.  Used for grids whose full table would
    not reasonably fit the host (1024^3 ComplexF64 = 16 GiB)."""
    d = len(grid)
    g64 = [np.asarray(g, dtype=np.float64) for g in grid]
    rng = np.random.default_rng(0x5E9)
    nprobe = 2048
    idx = [rng.integers(0, len(g), size=nprobe) for g in g64]
    zero = [np.full(nprobe, g[0]) for g in g64]
    pts = [g[i] for g, i in zip(g64, idx)]

    def D(p, n):
        v = f(tuple(p), param)
        if isinstance(v, (SVector, SMatrix)):
            return None
        return np.asarray(v, dtype=np.complex128) + np.zeros(n)

    full, origin = D(pts, nprobe), D(zero, nprobe)
    if full is None or origin is None:
        return None
    acc = full + (d - 1) * origin
    for a in range(d):
        acc = acc - D([pts[b] if b == a else zero[b] for b in range(d)], nprobe)
    scale = max(1e-300, float(np.abs(full).max()))
    if float(np.abs(acc).max()) > 1e-12 * scale:
        return None
    out = []
    for a in range(d):
        n = len(g64[a])
        line = D([g64[b] if b == a else np.full(n, g64[b][0]) for b in range(d)], n)
        if a > 0:
            line = line - D([np.full(n, g64[b][0]) for b in range(d)], n)
        out.append(np.ascontiguousarray(_cis(-np.float64(dt) * line), dtype=np.complex128))
    return out


# ------------------------------------------------------------------------------------------------
# closure recognition (registered forms)
# ------------------------------------------------------------------------------------------------


def recognise_nonlinearity(f, param, M, rng=None):
    """Fit the closure by probing and verify on held-out samples.  Registered forms:
      Number / SVector   G_i(u)  = c_i  + sum_j g_ij  |u_j|^2    -> ("scalar" | "vector", c[M], g[M][M])
      SMatrix (M x M)    G_ij(u) = C_ij + sum_k g_ijk |u_k|^2    -> ("matrix", C[M][M], g[M][M][M])
    (the matrix form: src/kernels.jl:22-25 -- `cis` of an SMatrix is the matrix exponential, docs
    general_overview.md:77)."""
    rng = rng or np.random.default_rng(0xC0FFEE)
    P = 4 * (M + 1) + 8

    def probe(n):
        amp = rng.uniform(0.2, 2.0, size=(M, n))
        ph = np.exp(2j * np.pi * rng.uniform(size=(M, n)))
        return amp * ph

    def evaluate(u):
        val = f(SVector([u[j] for j in range(M)]), param)
        kind = "scalar"
        if isinstance(val, SMatrix):
            if val.n == 1:
                val = SVector(val[0, 0])
            elif val.n == M:
                rows = [np.broadcast_to(np.asarray(val[i, j], dtype=complex), u.shape[1:])
                        for i in range(M) for j in range(M)]
                return "matrix", np.stack(rows)
            else:
                raise UnsupportedForm("matrix-valued nonlinearity must be M x M")
        if isinstance(val, SVector):
            kind = "vector"
            if len(val) == 1 and isinstance(val[0], SVector):
                val = val[0]
            rows = [np.broadcast_to(np.asarray(v, dtype=complex), u.shape[1:]) for v in val]
        else:
            rows = [np.broadcast_to(np.asarray(val, dtype=complex), u.shape[1:])]
        return kind, np.stack(rows)

    u = probe(P)
    kind, G = evaluate(u)
    if kind == "vector" and G.shape[0] != M:
        raise UnsupportedForm("nonlinearity must return a Number, an SVector of length M or an M x M SMatrix")
    A = np.concatenate([np.ones((1, P)), np.abs(u) ** 2], axis=0).T        # (P, M+1)
    coef, *_ = np.linalg.lstsq(A, G.T, rcond=None)                          # (M+1, rows)
    v = probe(16)
    _, Gv = evaluate(v)
    pred = (np.concatenate([np.ones((1, 16)), np.abs(v) ** 2], axis=0).T @ coef).T
    scale = max(1e-300, np.abs(Gv).max(), np.abs(coef).max())
    if np.abs(pred - Gv).max() > 1e-9 * scale:
        raise UnsupportedForm("nonlinearity is not of a registered form (c_i + sum_j g_ij |u_j|^2, or the same per "
                              "entry of an SMatrix); the B200 backend has no CPU fallback")
    coef = coef.copy()
    coef[np.abs(coef) < 1e-13 * scale] = 0                                  # round-off of the fit
    if kind == "matrix":
        Cm = coef[0].reshape(M, M)
        g = coef[1:].T.reshape(M, M, M)                                     # g[i][j][k]
        return "matrix", Cm, g
    rows = G.shape[0]
    c = np.zeros(M, dtype=complex)
    g = np.zeros((M, M), dtype=complex)
    for i in range(M):
        src = 0 if rows == 1 else i
        c[i] = coef[0, src]
        g[i, :] = coef[1:, src]
    if kind == "vector" and rows == 1 and M == 1:
        kind = "scalar"
    return kind, c, g


class PumpModel:
    """The pump closure on the device: either the registered separable form  F_i(r, t) = S_i(r) a(t)  (S table
    (npoints, ncomp) + amplitude schedule; `dense` False) or, when the closure does not separate, the reference's own
    procedure -- the closure evaluated on the whole direct grid at every half-step time (evaluate_pump!,
    src/misc.jl:34-42) -- handed to the library profile by profile (`dense` True, GGP_PUMP_DENSE).

    Recognition never decides from a handful of sampled times (a rectangular pulse between two samples
    was dropped, a transient second profile was lost): the closure is evaluated at K probe points for EVERY scheduled
    time, F(r_k, t) = a(t) S(r_k) must hold at all of them, and `zero` needs F = 0 at every scheduled time."""

    NPROBE = 12

    def __init__(self, f, prob, tspan, times, grid=None, force_dense=False):
        grid = grid if grid is not None else direct_grid(prob)
        pts = _mesh(grid)
        shape = tuple(len(g) for g in reversed(grid))
        self.f, self.param, self.M = f, prob.param, len(prob.u0)
        self.grid, self.shape = grid, shape
        self._pts = pts

        sched = [tspan[0]] + [t for t in times]
        # (1) full-grid evaluations at a few times pick the reference profile and the probe points
        cand = [tspan[0], tspan[-1], 0.5 * (tspan[0] + tspan[-1])]
        if len(times):
            cand += [times[(k * (len(times) - 1)) // 7] for k in range(8)]
        best, tref = None, None
        for t in cand:
            F = self.on_grid(t)
            if best is None or np.abs(F).max() > np.abs(best).max():
                best, tref = F, t
        self.ncomp = best.shape[1]
        if self.ncomp not in (1, self.M):
            raise UnsupportedForm("pump must return a Number or an SVector of length M")
        npts = best.shape[0]
        mag = np.abs(best).max(axis=1)
        order = np.argsort(-mag, kind="stable")
        probe = [int(order[0])] + [int(order[(k * (npts - 1)) // (self.NPROBE - 1)]) for k in range(1, self.NPROBE)]
        probe += [int((k * (npts - 1)) // (self.NPROBE - 1)) for k in range(self.NPROBE)]      # + a regular spread
        self.probe = np.array(sorted(set(probe)))
        multi = np.unravel_index(self.probe, shape)                                              # (i_d, ..., i_1)
        self._probe_pts = tuple(np.asarray(g)[i] for g, i in zip(grid, reversed(multi)))
        # (2) the closure at the probe points for every scheduled time
        vals = np.stack([self._at_probes(t) for t in sched])                                     # (ntimes, K, ncomp)
        self.zero = bool(np.abs(vals).max() == 0 and np.abs(best).max() == 0)
        self.dense = False
        self.S, self.tref = best.copy(), tref
        self._sched_amp = None
        if self.zero:
            return
        if np.abs(best).max() == 0:                         # all full-grid candidates vanish, but a probe saw the pump
            k = int(np.abs(vals).max(axis=(1, 2)).argmax())
            best, tref = self.on_grid(sched[k]), sched[k]
            self.S, self.tref = best.copy(), tref
        flat = int(np.abs(best).argmax())
        self.pidx, self.cidx = np.unravel_index(flat, best.shape)
        self._ref = best[self.pidx, self.cidx]              # a(tref) == 1 by construction
        multi1 = np.unravel_index(self.pidx, shape)
        self.rpt = tuple(np.asarray(g[i]).reshape(()) for g, i in zip(grid, reversed(multi1)))
        if force_dense or os.environ.get("GGP_PUMP_FORCE_DENSE"):
            self.dense = True
            return
        # (3) separability: F(r_k, t) == a(t) S(r_k) at every probe point and every scheduled time ...
        amps = np.array([self.amp(t) for t in sched])
        Sk = self.S[self.probe]                                                                  # (K, ncomp)
        scale = max(np.abs(vals).max(), np.abs(Sk).max() * np.abs(amps).max(), 1e-300)
        if np.abs(vals - amps[:, None, None] * Sk[None]).max() > 1e-10 * scale:
            self.dense = True
            return
        # ... and on the full grid at the candidate times
        for t in cand:
            F = self.on_grid(t)
            if np.abs(F - self.amp(t) * self.S).max() > 1e-10 * max(np.abs(F).max(), np.abs(self.S).max()):
                self.dense = True
                return
        self._sched_amp = amps

    def on_grid(self, t):
        """The closure on the whole direct grid at time t: (npoints, ncomp) complex (grid_map!, src/misc.jl:34-37)."""
        val = self.f(self._pts, self.param, t)
        if isinstance(val, SMatrix):
            if val.n != 1:
                raise UnsupportedForm("matrix-valued pump")
            val = SVector(val[0, 0])
        if isinstance(val, SVector):
            cols = [np.broadcast_to(np.asarray(v, dtype=complex), self.shape).reshape(-1) for v in val]
        else:
            cols = [np.broadcast_to(np.asarray(val, dtype=complex), self.shape).reshape(-1)]
        return np.stack(cols, axis=1)

    def _at_probes(self, t):
        val = self.f(self._probe_pts, self.param, t)
        if isinstance(val, SMatrix):
            val = SVector(val[0, 0])
        K = len(self.probe)
        if isinstance(val, SVector):
            cols = [np.broadcast_to(np.asarray(v, dtype=complex), (K,)) for v in val]
        else:
            cols = [np.broadcast_to(np.asarray(val, dtype=complex), (K,))]
        return np.stack(cols, axis=1)

    def amp(self, t):
        if self.zero:
            return 0.0 + 0.0j
        val = self.f(self.rpt, self.param, t)
        if isinstance(val, SMatrix):
            val = SVector(val[0, 0])
        if isinstance(val, SVector):
            val = val[self.cidx]
        return complex(np.asarray(val, dtype=complex).reshape(())) / complex(self._ref)


def noise_points(grid):
    """The `point`s the reference hands to the noise amplitude function: every grid axis indexed with the FIRST
    index K[1] (src/kernels.jl:27,41).  Returns an (n1, d) array, or None where Julia would throw a
    BoundsError (n1 longer than another axis)."""
    n1 = len(grid[0])
    if any(len(g) < n1 for g in grid):
        return None
    return np.stack([np.asarray(g[:n1]) for g in grid], axis=1)


def recognise_noise(f, prob, rng=None):
    """Registered form  eta_i(u, r) = P(r) (e_i + sum_j a_ij |u_j|):  constant amplitudes
    (examples/truncated_wigner.jl:96, test/windowed_ft.jl:27-29, SVector docs/src/stochastic_simulations.md:80-86),
    field-dependent `alpha*abs(u[1])` (:68-72) and spatial profiles (:74-78).
    Returns (e[M], a[M][M], P) -- P is None (no position dependence) or the n1 values of the profile at the
    reference's Q2 points, normalised to 1 at the probe point."""
    rng = rng or np.random.default_rng(0xBEEF)
    M = len(prob.u0)
    grid = direct_grid(prob)
    pts = noise_points(grid)

    def evaluate(u, r):
        v = f(SVector([u[j] for j in range(M)]), r, prob.param)
        if isinstance(v, SMatrix):
            raise UnsupportedForm("matrix-valued noise amplitudes are not a registered form")
        if isinstance(v, SVector):
            if len(v) != M:
                raise UnsupportedForm("noise amplitude SVector must have length M")
            return np.array([complex(np.asarray(x).reshape(())) for x in v])
        return np.full(M, complex(np.asarray(v).reshape(())))

    def probe():
        return rng.uniform(0.2, 2.0, size=M) * np.exp(2j * np.pi * rng.uniform(size=M))

    r_mid = tuple(g[len(g) // 2] for g in grid) if pts is None else tuple(pts[len(pts) // 2])
    # reference point: where the amplitude is largest along the Q2 points (a profile may vanish somewhere)
    u_ref = probe()
    k0 = None
    if pts is not None:
        along = np.array([evaluate(u_ref, tuple(pts[k])) for k in range(len(pts))])         # (n1, M)
        k0 = int(np.abs(along).max(axis=1).argmax())
        r0 = tuple(pts[k0])
    else:
        r0 = r_mid
    # field dependence at r0:  least squares on the features (1, |u_1|, ..., |u_M|), checked on held-out probes
    P_ = 4 * (M + 1) + 8
    U = [probe() for _ in range(P_)]
    A = np.array([[1.0] + list(np.abs(u)) for u in U])
    Y = np.array([evaluate(u, r0) for u in U])                                                # (P_, M)
    coef, *_ = np.linalg.lstsq(A, Y, rcond=None)                                              # (M+1, M)
    V = [probe() for _ in range(12)]
    pred = np.array([[1.0] + list(np.abs(u)) for u in V]) @ coef
    truth = np.array([evaluate(u, r0) for u in V])
    scale = max(1e-300, np.abs(truth).max(), np.abs(coef).max())
    if np.abs(pred - truth).max() > 1e-9 * scale:
        raise UnsupportedForm("noise amplitude is not of the registered form P(r) (e_i + sum_j a_ij |u_j|); "
                              "the B200 backend has no CPU fallback")
    coef[np.abs(coef) < 1e-13 * scale] = 0
    e, a = coef[0].copy(), coef[1:].T.copy()                                                  # a[i][j]
    # position dependence
    profile = None
    if pts is None:
        # Julia would index out of bounds as soon as the closure is called; closures that ignore r are fine
        other = evaluate(u_ref, tuple(g[0] for g in grid))
        if np.abs(other - evaluate(u_ref, r0)).max() > 1e-12 * scale:
            raise UnsupportedForm("position-dependent noise with n1 longer than another axis: the reference's "
                                  "`point` (src/kernels.jl:27,41) is out of bounds there")
    else:
        ref = evaluate(u_ref, r0)
        c = int(np.abs(ref).argmax())
        if abs(ref[c]) > 0:
            prof = along[:, c] / ref[c]
            if np.abs(prof - 1).max() > 1e-12:
                u2 = probe()
                base = evaluate(u2, r0)
                for k in (0, len(pts) // 3, len(pts) - 1):
                    if np.abs(evaluate(u2, tuple(pts[k])) - prof[k] * base).max() > 1e-9 * scale:
                        raise UnsupportedForm("noise amplitude does not separate as P(r) x (field part)")
                if np.abs(along - prof[:, None] * ref[None, :]).max() > 1e-9 * scale:
                    raise UnsupportedForm("noise amplitude profile differs between components")
                profile = np.ascontiguousarray(prof, dtype=np.complex128)
    return e, a, profile


# ------------------------------------------------------------------------------------------------
# iterator = StrangSplittingIterator (src/strang_splitting.jl:9-67) with a device plan inside
# ------------------------------------------------------------------------------------------------


def _ptr_array(arrs):
    arr = (C.c_void_p * len(arrs))()
    for i, a in enumerate(arrs):
        arr[i] = a.ctypes.data
    return arr


class _PinnedBlock:
    """Owner of one ggp_host_alloc block; frees it when the NumPy array built on top of it dies."""

    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            if self.ptr:
                self.lib.ggp_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class StrangSplittingIterator:
    def __init__(self, prob, tspan, *, dt, nsaves, save_start=True, rng=None, device=-1,
                 batch_offset=0, stream=None, slab=None, slab_local=False, result_buffers=True):
        """slab=(rank, world): 3-D slab decomposition, one process per GPU.  `prob.u0` is the GLOBAL field
        (each rank keeps z-planes [rank*n3/world, (rank+1)*n3/world)), or already this rank's z-slab if
        slab_local=True.  Results / fetch() are the local z-slab.  Attach a communicator
        (parallel.attach_nccl) before stepping."""
        lib = L.load()
        self.lib = lib
        self.prob = prob
        self.dt, self.ts, self.steps_per_save = resolve_fixed_timestepping(dt, tspan, nsaves)
        self.nsaves, self.save_start = nsaves, bool(save_start)
        self.tspan = tspan
        M = len(prob.u0)
        self.M = M
        dtype = prob.u0[0].dtype
        if dtype not in (np.complex64, np.complex128):
            raise ValueError("fields must be ComplexF32 or ComplexF64")
        self.dtype = dtype
        sizes = prob.sizes
        u0_local = prob.u0
        self.slab = slab
        dg_pump = None
        if slab is not None:
            rank, world = slab
            if prob.ndim != 3 or prob.u0[0].ndim != 3:
                raise ValueError("slab decomposition needs a 3-D problem without batch dims")
            if slab_local:                                   # u0 is this rank's z-slab: rebuild the global sizes
                sizes = (sizes[0], sizes[1], sizes[2] * world)
                gshape = (sizes[2], sizes[1], sizes[0])
                prob_g = GrossPitaevskiiProblem(tuple(np.broadcast_to(np.zeros((), dtype), gshape) for _ in prob.u0),
                                                prob.lengths)
            else:
                prob_g = prob
            n3l, n2l = sizes[2] // world, sizes[1] // world
            z0, y0 = rank * n3l, rank * n2l
            if not slab_local:
                u0_local = tuple(np.ascontiguousarray(x[z0:z0 + n3l]) for x in prob.u0)
            rg_g, dg_g = reciprocal_grid(prob_g), direct_grid(prob_g)
            rg = (rg_g[0], rg_g[1][y0:y0 + n2l], rg_g[2])     # y-slab: the layout of the z pass
            dg = (dg_g[0], dg_g[1], dg_g[2][z0:z0 + n3l])     # z-slab: the resident layout
            dg_pump = dg
            nspatial = int(np.prod(sizes)) // world
            self.nbatch = 1
        else:
            nspatial = int(np.prod(sizes))
            self.nbatch = int(prob.u0[0].size // nspatial)
            rg, dg = reciprocal_grid(prob), direct_grid(prob)
        self.u0_local = u0_local
        self._result_shape = [(nsaves + self.save_start,) + tuple(x.shape) for x in u0_local]
        self.result = None                                   # allocated page-locked once the library is up

        # exp_D: the reference's own table (get_exponential, :53) -- except for grids whose table would not
        # reasonably fit the host, where a dispersion that is a sum over axes is handed over as d per-axis factors
        # (GGP_TABLE_SEP_AXES; GGP_SEP_AXES_MIN = smallest number of points that takes this route, default 2^26)
        daxes = None
        npts_full = int(np.prod(sizes))
        if (M == 1 and prob.ndim >= 2 and not _absent(prob.dispersion)
                and all(int(n_) & (int(n_) - 1) == 0 for n_ in sizes)
                and npts_full >= int(os.environ.get("GGP_SEP_AXES_MIN", 1 << 26))):
            rg_full = reciprocal_grid(prob_g) if slab is not None else rg
            daxes = dispersion_axis_factors(prob.dispersion, rg_full, prob.param, self.dt)
        if daxes is not None:
            dkind, dtab = L.TABLE_SEP_AXES, None
        else:
            dkind, dtab = exp_table(prob.dispersion, rg, prob.param, self.dt, M)        # :53
        vkind, vtab = exp_table(prob.potential, dg, prob.param, self.dt / 2, M)         # :54

        d = L.GgpDesc()
        d.abi_version, d.struct_size = L.GGP_ABI_VERSION, C.sizeof(L.GgpDesc)
        d.ndim, d.ncomp = prob.ndim, M
        for i, n in enumerate(sizes):
            d.n[i] = int(n)
        d.nbatch, d.batch_offset = self.nbatch, int(batch_offset)
        d.precision = L.GGP_C64 if dtype == np.complex64 else L.GGP_C128
        d.table_precision = L.GGP_C128
        d.device = device
        d.stream = stream
        d.dt = float(self.dt)
        if slab is not None:
            d.slab_rank, d.slab_nranks = int(slab[0]), int(slab[1])
        keep = []

        def as_c128(t):
            a = np.ascontiguousarray(t, dtype=np.complex128)
            keep.append(a)
            return a.ctypes.data

        d.disp_kind, d.pot_kind = dkind, vkind
        if dkind == L.TABLE_SCALAR and dtype == np.complex64 and slab is None and not os.environ.get("GGP_NO_SEP_HINT"):
            d.disp_sep_tol = separable_dispersion_tol(prob.dispersion, rg, prob.param, dtab)
        # quirk Q6: ComplexF32 fields stepped with a Float64 dt / lengths hold ComplexF64 tables in the reference
        if dtype == np.complex64 and dtab is not None and np.asarray(dtab).dtype == np.complex128 \
                and not os.environ.get("GGP_NO_Q6"):
            d.mixed_precision_tables = 1
        d.disp_table = as_c128(dtab) if dtab is not None else None
        if daxes is not None:
            for a_, ax_ in enumerate(daxes):
                d.disp_axes[a_] = as_c128(ax_)
        d.pot_table = as_c128(vtab) if vtab is not None else None

        if not _absent(prob.nonlinearity):
            kind, c, g = recognise_nonlinearity(prob.nonlinearity, prob.param, M)
            scalar = kind == "scalar"
            if kind == "matrix" or M > 2:
                # generic plan (ABI 5): coefficient arrays of any M
                d.nl_kind, d.nl_scalar = (L.NL_MATRIX if kind == "matrix" else L.NL_DIAG), int(scalar)
                cc = c if kind != "scalar" else c[:1]
                gg = g if kind != "scalar" else g[:1]
                d.nl_c_ext = as_c128(np.asarray(cc, dtype=np.complex128).reshape(-1))
                d.nl_g_ext = as_c128(np.asarray(gg, dtype=np.complex128).reshape(-1))
            else:
                d.nl_kind, d.nl_scalar = L.NL_DIAG, int(scalar)
                for i in range(M):
                    d.nl_c[i][0], d.nl_c[i][1] = c[i].real, c[i].imag
                    for j in range(M):
                        d.nl_g[i][j][0], d.nl_g[i][j][1] = g[i, j].real, g[i, j].imag
            self.nl = (kind, c, g)

        # pump amplitude schedule at the reference's times
        nsteps = nsaves * self.steps_per_save
        self.pump_model = None
        self.amps = None
        if not _absent(prob.pump):
            t = self.ts[0]
            times = np.empty((nsteps, 2), dtype=self.ts.dtype)
            for i in range(nsteps):
                t = t + self.dt
                times[i, 0] = t + self.dt / 2
                times[i, 1] = t + self.dt
            pm = PumpModel(prob.pump, prob, (self.ts[0], _jl(tspan[-1])), times.reshape(-1), grid=dg_pump)
            self.pump_model = pm
            self.pump_times = times
            if pm.zero:
                pass                        # F = 0 at every scheduled time: no pump term at all
            elif pm.dense:
                # not S(r) a(t): the reference's own procedure, profile by profile (src/misc.jl:34-42)
                d.pump_kind, d.pump_ncomp = L.PUMP_DENSE, pm.ncomp
                d.pump_table = as_c128(pm.on_grid(self.ts[0]))                          # :58
            else:
                d.pump_kind, d.pump_ncomp = L.PUMP_SEPARABLE, pm.ncomp
                d.pump_table = as_c128(pm.S)
                sched = pm._sched_amp                                                   # a(t0), then a at `times`
                a0 = complex(sched[0])                                                  # :58
                d.pump_amp0[0], d.pump_amp0[1] = a0.real, a0.imag
                amps = np.asarray(sched[1:], dtype=np.complex128).reshape(nsteps, 2)
                if np.all(amps == a0):
                    self.amps = None        # static pump: the library repeats pump_amp0
                else:
                    self.amps = np.ascontiguousarray(amps)

        self.noise_real = False
        if not _absent(prob.position_noise_func):
            if _absent(prob.noise_prototype):
                raise ValueError("position_noise_func needs a noise_prototype")
            eta, alpha, profile = recognise_noise(prob.position_noise_func, prob)
            protos = tuple(np.asarray(x) for x in prob.noise_prototype)
            # The reference indexes xi with the leading ndims(xi) indices only (src/kernels.jl:24,27): a prototype
            # without the batch dims would share one noise field between all trajectories, and mixed real / complex
            # prototypes draw differently per component.  The device draws one independent stream per (element,
            # trajectory, component), so only the shapes every shipped example uses are accepted.
            if len(protos) != M or any(x.shape != prob.u0[0].shape for x in protos):
                raise UnsupportedForm("noise_prototype must hold one array per component with the shape of u0 "
                                      "(prototypes without the batch dims share noise between trajectories in the "
                                      "reference; not a registered form)")
            if len({np.iscomplexobj(x) for x in protos}) != 1:
                raise UnsupportedForm("noise_prototype arrays must be all real or all complex")
            proto = protos[0]
            self.noise_real = not np.iscomplexobj(proto)
            field = bool(np.any(alpha != 0)) or profile is not None
            if field and slab is not None:
                raise UnsupportedForm("field-/position-dependent noise with a slab decomposition")
            d.noise_kind, d.noise_real = (L.NOISE_FIELD if field else L.NOISE_CONST), int(self.noise_real)
            if M > 2:
                d.noise_eta_ext = as_c128(np.asarray(eta, dtype=np.complex128).reshape(-1))
                d.noise_alpha_ext = as_c128(np.asarray(alpha, dtype=np.complex128).reshape(-1))
            else:
                for i in range(M):
                    d.noise_eta[i][0], d.noise_eta[i][1] = eta[i].real, eta[i].imag
                    for j in range(M):
                        d.noise_alpha[i][j][0], d.noise_alpha[i][j][1] = alpha[i, j].real, alpha[i, j].imag
            if profile is not None:
                d.noise_profile = as_c128(profile)
            self.noise_form = (eta, alpha, profile)
            if rng is None:
                seed = int.from_bytes(os.urandom(8), "little")
            elif isinstance(rng, (int, np.integer)):
                seed = int(rng)
            else:
                seed = int(rng.integers(0, 2 ** 63))
            d.seed = seed & (2 ** 64 - 1)
            self.seed = d.seed

        self._desc, self._keep = d, keep
        handle = C.c_void_p()
        L.check(lib.ggp_plan_create(C.byref(d), C.byref(handle)))
        self.handle = handle
        # result[..., n] slices (src/strang_splitting.jl:41-43), every slot pre-filled with u0; page-locked so the
        # streaming saves (ggp_save_async) overlap the next save interval.  Julia's trailing save index is the
        # leading NumPy axis over the same memory.
        res = []
        if result_buffers:                  # (benchmarks of multi-GiB states step without a result array)
            for x, shp in zip(u0_local, self._result_shape):
                r = self._pinned_empty(shp, x.dtype)
                r[...] = x
                res.append(r)
        self.result = tuple(res) if result_buffers else None
        self.u = [self._pinned_like(x) for x in u0_local]                               # :48 (pinned staging)
        for dst, src in zip(self.u, u0_local):
            np.copyto(dst, src)
        L.check(lib.ggp_set_state(self.handle, _ptr_array(self.u)))
        self._step_index = 0

    def _pinned_like(self, x):
        return self._pinned_empty(x.shape, x.dtype)

    def _pinned_empty(self, shape, dtype):
        """Page-locked host array (async DMA at full PCIe rate); plain NumPy if that fails.  The block is owned by
        the array: it is released when the last view of it is garbage-collected, NOT by close() -- `solve`
        hands `result` to the caller after the plan is gone."""
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        ptr = self.lib.ggp_host_alloc(nbytes) if nbytes else None
        if not ptr:
            return np.empty(shape, dtype=dtype)
        buf = (C.c_char * nbytes).from_address(ptr)
        buf._ggp_owner = _PinnedBlock(self.lib, ptr)          # np.frombuffer keeps `buf` alive, `buf` keeps the block
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def upload(self, u0=None):
        """Host -> device copy of the state (`u = copy.(prob.u0)`, src/strang_splitting.jl:48)."""
        if u0 is not None:
            for dst, src in zip(self.u, u0):
                np.copyto(dst, src)
        L.check(self.lib.ggp_set_state(self.handle, _ptr_array(self.u)))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ggp_plan_destroy(self.handle)
            self.handle = None
            self.u = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- device stepping ---------------------------------------------------------------------
    def advance(self, nsteps, noise_buffers=None):
        """nsteps x step! on the device (src/fixed_time_stepping.jl:43-47)."""
        amp_ptr = None
        if self.amps is not None:
            seg = np.ascontiguousarray(self.amps[self._step_index:self._step_index + nsteps])
            assert seg.shape[0] == nsteps, "stepping past the end of the pump schedule"
            amp_ptr = seg.ctypes.data
        nptr = None
        bufs = None
        if noise_buffers is not None:
            real_t = np.float32 if self.dtype == np.complex64 else np.float64
            want = real_t if self.noise_real else self.dtype
            bufs = [np.ascontiguousarray(b, dtype=want) for b in noise_buffers]
            assert len(bufs) == 2 * nsteps * self.M, "need 2*nsteps*M noise buffers (step, half-step, component)"
            nptr = _ptr_array(bufs)
        pm = self.pump_model
        if pm is not None and pm.dense and not pm.zero:
            # evaluate_pump! on the host for every half-step of the interval (src/misc.jl:34-42), a few steps at a time
            done = 0
            while done < nsteps:
                k = min(8, nsteps - done)
                i0 = self._step_index + done
                assert i0 + k <= self.pump_times.shape[0], "stepping past the end of the pump schedule"
                profs = [np.ascontiguousarray(pm.on_grid(self.pump_times[i0 + i, h]), dtype=np.complex128)
                         for i in range(k) for h in (0, 1)]
                sub = None if bufs is None else _ptr_array(bufs[2 * done * self.M:2 * (done + k) * self.M])
                L.check(self.lib.ggp_step_dense(self.handle, k, _ptr_array(profs), sub))
                done += k
        else:
            L.check(self.lib.ggp_step(self.handle, nsteps, amp_ptr, nptr))
        self._step_index += nsteps

    def fetch(self):
        L.check(self.lib.ggp_get_state(self.handle, _ptr_array(self.u)))
        return self.u

    def save_async(self, slot):
        """Non-blocking `map(copy!, slice, iter.u)` into result[..., slot] (src/fixed_time_stepping.jl:48): device
        snapshot in stream order, PCIe transfer on a second stream while the next interval steps (ggp_save_async)."""
        L.check(self.lib.ggp_save_async(self.handle, _ptr_array([r[slot] for r in self.result])))

    def save_wait(self):
        L.check(self.lib.ggp_save_wait(self.handle))

    # -- checkpoint / resume ---------------------------------------------------
    def checkpoint(self):
        """Bytes that continue this run bit-identically when given to `restore` of an iterator built from the same
        problem / tspan / dt / nsaves: the library blob (fields, Philox counter word, F_now amplitude) followed by
        the host-side position in the step / pump schedule."""
        n = int(self.lib.ggp_checkpoint_bytes(self.handle))
        if n < 0:
            L.check(n)
        buf = np.empty(n + 8, dtype=np.uint8)
        L.check(self.lib.ggp_checkpoint_save(self.handle, buf.ctypes.data, n))
        buf[n:] = np.frombuffer(np.int64(self._step_index).tobytes(), dtype=np.uint8)
        return buf.tobytes()

    def restore(self, blob):
        buf = np.frombuffer(blob, dtype=np.uint8)
        n = int(self.lib.ggp_checkpoint_bytes(self.handle))
        if buf.size != n + 8:
            raise ValueError("checkpoint size does not match this problem")
        L.check(self.lib.ggp_checkpoint_load(self.handle, buf.ctypes.data, n))
        self._step_index = int(np.frombuffer(buf[n:].tobytes(), dtype=np.int64)[0])

    def observe_windowed(self, w1, w2):
        """sum_traj conj(F2[j]) F1[i], F_a = ifftshift(fft(fftshift(u .* w_a)))  (test/windowed_ft.jl:31-49, before its
        division by length(sol)), per component, computed on the device: complex array (M, N, N)."""
        N = self.u[0].shape[-1]
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(w1, dtype=np.complex128), (N,)))
        b = np.ascontiguousarray(np.broadcast_to(np.asarray(w2, dtype=np.complex128), (N,)))
        out = np.empty((self.M, N, N), dtype=np.complex128)
        L.check(self.lib.ggp_observe_windowed(self.handle, a.ctypes.data, b.ctypes.data, out.ctypes.data))
        return out

    def observe(self, kind):
        shape = self.u[0].shape[self.u[0].ndim - self.prob.ndim:]
        nspatial = int(np.prod(shape))
        n = self.M if kind == L.OBS_NORM else self.M * nspatial
        if kind == L.OBS_G2_MOMENTUM:
            n = self.M * nspatial * nspatial
        out = np.empty(n, dtype=np.float64)
        L.check(self.lib.ggp_observe(self.handle, kind, out.ctypes.data))
        if kind == L.OBS_NORM:
            return out
        if kind == L.OBS_G2_MOMENTUM:                     # [c][m][n] = sum_traj |F_c(m)|^2 |F_c(n)|^2, F = fft(u_c)/N
            return out.reshape(self.M, nspatial, nspatial)
        return out.reshape((self.M,) + tuple(shape))


def init(prob, alg, tspan, *, dt, nsaves, show_progress=True, progress=None, save_start=True,
         workgroup_size=(), rng=None, **backend_kw):
    """CommonSolve.init (src/strang_splitting.jl:32-39).  `workgroup_size` is accepted and ignored
    (block sizes are fixed per kernel); `show_progress` / `progress` are host-side cosmetics."""
    assert isinstance(alg, StrangSplitting)
    return StrangSplittingIterator(prob, tspan, dt=dt, nsaves=nsaves, save_start=save_start, rng=rng, **backend_kw)


def step_(it, t=None, dt=None, noise_buffers=None):
    """CommonSolve.step! (src/strang_splitting.jl:86-90): one Strang step on the device."""
    it.advance(1, noise_buffers)


def solve_(it, noise_buffers=None):
    """CommonSolve.solve! (src/fixed_time_stepping.jl:26-54)."""
    off = 1 if it.save_start else 0
    t = it.ts[0]
    sps, M = it.steps_per_save, it.M
    for n in range(it.nsaves):
        nb = None
        if noise_buffers is not None:
            nb = noise_buffers[n * sps * 2 * M:(n + 1) * sps * 2 * M]
        it.advance(sps, nb)                                  # :43-47 batched into one call
        for _ in range(sps):
            t = t + it.dt                                    # :44 (accumulated in T)
        it.save_async(n + off)                               # :48, overlapped with the next interval
        it.ts[n + 1] = t                                     # :49
    it.save_wait()
    return it.ts[1 - off:], it.result                        # :53


def solve(prob, alg, tspan, *, noise_buffers=None, **kw):
    """solve(prob, StrangSplitting(), tspan; dt, nsaves, ...) (src/fixed_time_stepping.jl:79-81)."""
    it = init(prob, alg, tspan, **kw)
    try:
        return solve_(it, noise_buffers)
    finally:
        it.close()

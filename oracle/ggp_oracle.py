"""CPU oracle for the Strang-splitting time step of GeneralizedGrossPitaevskii.jl.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product path
(generalizedgrosspitaevskii.jl_b200 -> libggp.so) never routes through this file.

It is a NumPy restatement of the reference's algorithm, function by function (file:line into
/root/reference/src):

  resolve_fixed_timestepping  fixed_time_stepping.jl:14-24
  direct_grid / reciprocal_grid  problem.jl:129-139 (+ AbstractFFTs `fftfreq` index rule)
  identity algebra (_mul/_add/_cis)  kernels.jl:1-25
  muladd (the fused point-wise kernel)  kernels.jl:37-54
  get_exponential  misc.jl:12-20
  pump double buffer  misc.jl:22-42
  perform_ft!  misc.jl:53-64   (forward unnormalised, inverse 1/N; == numpy.fft.fftn/ifftn)
  potential_pump_step! / diffusion_step! / step!  strang_splitting.jl:69-90
  init / solve!  strang_splitting.jl:32-67, fixed_time_stepping.jl:26-54

PARITY PINNING.  The reference cannot be executed here (no Julia, no FFTW) and ships no golden
vectors.  The oracle is pinned by the reference's own *known-answer* tests, re-expressed in
tests/test_oracle_known_answers.py (bistability curve, exciton-polariton steady state, windowed-FT
vacuum commutator, exact free propagation, scalar/SVector/SMatrix{1,1} wrapper equivalence).
Bit-level / 1e-10-level agreement with a real Julia run is therefore "parity unpinned" for:
  * the 2x2 matrix exponential (StaticArrays.jl `_exp(::Size{(2,2)})`, compat "1", no Manifest ->
    version unpinned; restated below from the published closed form and cross-checked against
    scipy.linalg.expm),
  * the RNG stream (Random.randn!, not reproducible outside Julia; noise is host-fed instead).

Array convention: a Julia array of size (n1, n2, ..., batch...) (column-major) is held here as a
NumPy C-order array of shape (batch..., ..., n2, n1) over the same memory.  Closures receive the
grid coordinates in Julia order, i.e. `ks[0]` is the coordinate along the fastest axis n1.
"""
from __future__ import annotations

import math
import numpy as np

# --------------------------------------------------------------------------------------------
# StaticArrays stand-ins.  Components are NumPy arrays (whole-grid vectorisation).
# --------------------------------------------------------------------------------------------


class SVector:
    """Stand-in for StaticArrays.SVector whose entries are whole-grid arrays (or scalars)."""

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

    def _bin(self, other, op):
        if isinstance(other, SVector):
            return SVector([op(a, b) for a, b in zip(self.items, other.items)])
        return SVector([op(a, other) for a in self.items])

    def __mul__(self, o):
        return self._bin(o, lambda a, b: a * b)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._bin(o, lambda a, b: a / b)

    def __add__(self, o):
        return self._bin(o, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, o):
        return self._bin(o, lambda a, b: a - b)

    def __neg__(self):
        return SVector([-a for a in self.items])


class SMatrix:
    """Stand-in for StaticArrays.SMatrix; `rows[i][j]` entries are whole-grid arrays (or scalars)."""

    def __init__(self, rows):
        self.rows = [list(r) for r in rows]
        n = len(self.rows)
        assert all(len(r) == n for r in self.rows), "square matrices only"

    @property
    def n(self):
        return len(self.rows)

    def __getitem__(self, ij):
        i, j = ij
        return self.rows[i][j]

    def scale(self, s):
        return SMatrix([[s * e for e in r] for r in self.rows])


def abs2(x):
    """Julia's abs2, broadcasting over SVector like `abs2.(ψ)`."""
    if isinstance(x, SVector):
        return SVector([abs2(a) for a in x.items])
    if isinstance(x, (tuple, list)):
        return SVector([abs2(a) for a in x])
    x = np.asarray(x)
    return x.real * x.real + x.imag * x.imag


# --------------------------------------------------------------------------------------------
# identity algebra -- kernels.jl:1-25
# --------------------------------------------------------------------------------------------


class _AdditiveIdentity:
    def __call__(self, *a, **k):  # kernels.jl:3
        return self

    def __repr__(self):
        return "additiveIdentity"


class _MultiplicativeIdentity:
    def __call__(self, *a, **k):  # kernels.jl:7
        return self

    def __repr__(self):
        return "multiplicativeIdentity"


additiveIdentity = _AdditiveIdentity()
multiplicativeIdentity = _MultiplicativeIdentity()


def _is_add_id(x):
    return isinstance(x, _AdditiveIdentity)


def _is_mul_id(x):
    return isinstance(x, _MultiplicativeIdentity)


def _mul2(x, y):
    """kernels.jl:9-14."""
    if _is_mul_id(x):
        return y
    if _is_mul_id(y):
        return x
    if _is_add_id(y):
        return additiveIdentity
    if _is_add_id(x):
        # Julia would raise MethodError only for number*AdditiveIdentity in the *first* slot;
        # `_mul(x, ::AdditiveIdentity)` is the only defined absorbing rule.  The reference never
        # produces it in the first slot; treat it as absorbing too.
        return additiveIdentity
    if isinstance(x, SVector) and isinstance(y, SVector):
        return SVector([a * b for a, b in zip(x.items, y.items)])  # elementwise, kernels.jl:10
    if isinstance(x, SMatrix) and isinstance(y, SVector):
        n = x.n
        return SVector([sum(x[i, j] * y[j] for j in range(n)) for i in range(n)])
    if isinstance(x, SMatrix) and isinstance(y, SMatrix):
        n = x.n
        return SMatrix([[sum(x[i, k] * y[k, j] for k in range(n)) for j in range(n)] for i in range(n)])
    if isinstance(x, SVector) and isinstance(y, SMatrix):
        raise TypeError("DimensionMismatch: SVector * SMatrix is not defined in the reference (kernels.jl:9)")
    if isinstance(x, SMatrix):
        return x.scale(y)
    if isinstance(y, SMatrix):
        return y.scale(x)
    if isinstance(x, SVector):
        return SVector([a * y for a in x.items])
    if isinstance(y, SVector):
        return SVector([x * b for b in y.items])
    return x * y


def _mul(x, *args):
    """kernels.jl:15 -- right fold."""
    if not args:
        return x
    return _mul2(x, _mul(*args))


def _add2(x, y):
    """kernels.jl:17-20 -- `x .+ y` broadcasts a scalar over an SVector."""
    if _is_add_id(x):
        return y
    if _is_add_id(y):
        return x
    if isinstance(x, SMatrix) and x.n == 1:
        x = SVector([x[0, 0]])
    if isinstance(y, SMatrix) and y.n == 1:
        y = SVector([y[0, 0]])
    if isinstance(x, SVector) and isinstance(y, SVector):
        return SVector([a + b for a, b in zip(x.items, y.items)])
    if isinstance(x, SVector):
        return SVector([a + y for a in x.items])
    if isinstance(y, SVector):
        return SVector([x + b for b in y.items])
    return x + y


def _add(x, *args):
    if not args:
        return x
    return _add2(x, _add(*args))


def _cis_scalar(z):
    """Julia `cis(z)` = exp(im*z); for complex z: exp(-imag z) * (cos(real z) + im sin(real z))."""
    z = np.asarray(z)
    if np.iscomplexobj(z):
        return np.exp(-z.imag) * (np.cos(z.real) + 1j * np.sin(z.real))
    return np.cos(z) + 1j * np.sin(z)


def exp2x2(a, c, b, d):
    """Closed-form 2x2 matrix exponential, following StaticArrays.jl `_exp(::Size{(2,2)}, A)`
    (src/expm.jl; column-major entries a=A11, c=A21, b=A12, d=A22).  Third-party arithmetic that
    is NOT under /root/reference (StaticArrays compat "1", unpinned).  Cross-checked against
    scipy.linalg.expm in tests/test_oracle_known_answers.py."""
    a, b, c, d = (np.asarray(v, dtype=np.result_type(a, b, c, d, np.complex64)) for v in (a, b, c, d))
    a, b, c, d = np.broadcast_arrays(a, b, c, d)
    z = np.sqrt((a - d) * (a - d) + 4 * b * c)
    e = np.expm1((a + d - z) / 2)
    f = np.expm1((a + d + z) / 2)
    eps = np.finfo(z.real.dtype).eps
    small = (z.real * z.real + z.imag * z.imag) < eps * eps
    zs = np.where(small, 1, z)
    g = np.where(small, np.exp((a + d) / 2) * (1 + z * z / 24), (f - e) / zs)
    m11 = (g * (a - d) + f + e) / 2 + 1
    m12 = g * b
    m21 = g * c
    m22 = (-g * (a - d) + f + e) / 2 + 1
    return m11, m21, m12, m22


def _cis(x):
    """kernels.jl:22-25.  SVector -> elementwise; SMatrix -> LinearAlgebra.cis = exp(im*A)."""
    if _is_add_id(x):
        return multiplicativeIdentity
    if isinstance(x, SVector):
        return SVector([_cis_scalar(a) for a in x.items])
    if isinstance(x, SMatrix):
        if x.n == 1:
            return SMatrix([[_cis_scalar(x[0, 0])]])
        if x.n == 2:
            m11, m21, m12, m22 = exp2x2(1j * x[0, 0], 1j * x[1, 0], 1j * x[0, 1], 1j * x[1, 1])
            return SMatrix([[m11, m12], [m21, m22]])
        # M > 2: StaticArrays falls through to the generic scaling-and-squaring Pade algorithm (its `_exp` for
        # sizes beyond 2x2); scipy.linalg.expm is the same published algorithm (Higham 2005 / Al-Mohy & Higham 2009)
        import scipy.linalg
        n = x.n
        ent = [[np.asarray(x[i, j], dtype=complex) for j in range(n)] for i in range(n)]
        shape = np.broadcast_shapes(*(e.shape for r in ent for e in r))
        A = np.stack([np.stack([np.broadcast_to(1j * ent[i][j], shape) for j in range(n)], axis=-1)
                      for i in range(n)], axis=-2)
        flat = A.reshape((-1, n, n))
        E = np.stack([scipy.linalg.expm(m) for m in flat]).reshape(A.shape)
        return SMatrix([[E[..., i, j] for j in range(n)] for i in range(n)])
    return _cis_scalar(x)


def _neg_scale(dt, x):
    """`-δt * f(...)` for every kind (misc.jl:15, kernels.jl:44)."""
    return _mul2(-dt, x)


# --------------------------------------------------------------------------------------------
# grids -- problem.jl:129-139
# --------------------------------------------------------------------------------------------


def _julia_float(x):
    """Julia literal semantics: a bare Python float/int behaves like Float64/Int (strong type)."""
    if isinstance(x, (np.floating, np.integer)):
        return x
    if isinstance(x, (int, np.integer)):
        return int(x)
    return np.float64(x)


def direct_grid_1d(L, n):
    """StepRangeLen(zero(L), L / N, N): x_j = (j-1) * (L/N).  problem.jl:129."""
    L = _julia_float(L)
    step = L / n if isinstance(L, np.floating) else np.float64(L) / n
    return np.arange(n).astype(np.asarray(step).dtype) * step


def reciprocal_grid_1d(L, n):
    """fftfreq(N, oftype(N/L, 2π) * N / L).  problem.jl:135.  AbstractFFTs.Frequencies:
    entry i (0-based) = (i - (i < n_nonnegative ? 0 : n)) * (fs/n), n_nonnegative = (n+1)>>1,
    so for even n the Nyquist bin is negative."""
    L = _julia_float(L)
    ratio = n / L if isinstance(L, np.floating) else np.float64(n) / np.float64(L)
    ftype = np.asarray(ratio).dtype
    fs = ftype.type(2 * math.pi) * ftype.type(n) / ftype.type(L)
    mult = fs / ftype.type(n)
    i = np.arange(n)
    nn = (n + 1) >> 1
    idx = np.where(i < nn, i, i - n).astype(ftype)
    return idx * mult


def _mesh(axes):
    """Coordinates in Julia order (axes[0] fastest) broadcast to NumPy shape (n_d, ..., n_1)."""
    d = len(axes)
    out = []
    for m, ax in enumerate(axes):
        shape = [1] * d
        shape[d - 1 - m] = len(ax)
        out.append(ax.reshape(shape))
    return tuple(out)


def _noise_points(axes):
    """The `point` the reference hands to the noise amplitude function (kernels.jl:41): build_field_at(grid, K)
    indexes EVERY 1-D grid axis with K[1] (`_getindex(A, K) = A[K[1:ndims(A)]...]`, kernels.jl:27) -- quirk Q2:
    point = (x[K1], y[K1], ...), a function of the first (fastest) index only; a BoundsError in Julia if n1 exceeds
    the length of another axis (IndexError here).  Identical to the true mesh in 1-D."""
    d = len(axes)
    n1 = len(axes[0])
    out = []
    for ax in axes:
        if len(ax) < n1:
            raise IndexError("BoundsError: noise `point` indexes every grid axis with K[1] (src/kernels.jl:27,41) "
                             "and n1 exceeds the length of another axis")
        out.append(np.asarray(ax[:n1]).reshape([1] * (d - 1) + [n1]))
    return tuple(out)


# --------------------------------------------------------------------------------------------
# problem container -- problem.jl:97-122
# --------------------------------------------------------------------------------------------


class GrossPitaevskiiProblem:
    def __init__(self, u0, lengths, *, dispersion=additiveIdentity, potential=additiveIdentity,
                 nonlinearity=additiveIdentity, pump=additiveIdentity,
                 position_noise_func=additiveIdentity, momentum_noise_func=additiveIdentity,
                 noise_prototype=additiveIdentity, param=None):
        u0 = tuple(u0)
        lengths = tuple(lengths)
        assert all(x.ndim >= len(lengths) for x in u0)            # problem.jl:112
        assert all(x.shape == u0[0].shape for x in u0)            # problem.jl:113
        self.u0 = tuple(np.asarray(x) if np.iscomplexobj(x) else np.asarray(x).astype(
            np.complex64 if np.asarray(x).dtype == np.float32 else np.complex128) for x in u0)  # :114
        ls = [_julia_float(l) for l in lengths]
        lt = np.result_type(*[np.asarray(l).dtype for l in ls])   # promote(lengths...) :115
        self.lengths = tuple(lt.type(l) for l in ls)
        self.dispersion, self.potential = dispersion, potential
        self.nonlinearity, self.pump = nonlinearity, pump
        self.position_noise_func, self.momentum_noise_func = position_noise_func, momentum_noise_func
        self.noise_prototype = noise_prototype
        self.param = param

    @property
    def ndim(self):
        return len(self.lengths)

    @property
    def spatial_shape_julia(self):
        """(n1, ..., nd), n1 fastest."""
        s = self.u0[0].shape
        return tuple(reversed(s[len(s) - self.ndim:]))

    def direct_grid(self):
        return tuple(direct_grid_1d(L, n) for L, n in zip(self.lengths, self.spatial_shape_julia))

    def reciprocal_grid(self):
        return tuple(reciprocal_grid_1d(L, n) for L, n in zip(self.lengths, self.spatial_shape_julia))


# --------------------------------------------------------------------------------------------
# fixed time stepping -- fixed_time_stepping.jl:14-24
# --------------------------------------------------------------------------------------------


def resolve_fixed_timestepping(dt, tspan, nsaves):
    dt = _julia_float(dt)
    t0, t1 = _julia_float(tspan[0]), _julia_float(tspan[-1])
    T = np.result_type(*[np.asarray(v).dtype for v in (dt, t0, t1)])
    if not np.issubdtype(T, np.floating):
        T = np.dtype(np.float64)                                  # float(promote_type(...))
    ts = np.empty(nsaves + 1, dtype=T)
    ts[0] = t0
    dT = T.type(T.type(t1) - T.type(t0)) / nsaves                 # T(last - first) / nsaves
    q = dT / dt
    steps_per_save = int(math.ceil(q))                            # round(Int, ΔT/dt, RoundUp)
    _dt = dT / steps_per_save
    return _dt, ts, steps_per_save


# --------------------------------------------------------------------------------------------
# tables -- misc.jl:12-20
# --------------------------------------------------------------------------------------------


def get_exponential(f, grid, param, dt):
    """Table of `_cis(-δt * f(point, param))` over the spatial grid (misc.jl:12-20)."""
    if _is_add_id(f):
        return multiplicativeIdentity                             # misc.jl:12
    pts = _mesh(grid)
    shape = tuple(len(g) for g in reversed(grid))
    val = _cis(_neg_scale(dt, f(pts, param)))
    return _broadcast_kind(val, shape)


def _broadcast_kind(val, shape):
    if isinstance(val, SVector):
        return SVector([np.broadcast_to(np.asarray(a), shape).copy() for a in val.items])
    if isinstance(val, SMatrix):
        return SMatrix([[np.broadcast_to(np.asarray(e), shape).copy() for e in r] for r in val.rows])
    if _is_add_id(val) or _is_mul_id(val):
        return val
    return np.broadcast_to(np.asarray(val), shape).copy()


def evaluate_pump(prob, t):
    """grid_map!(dest, prob.pump, direct_grid, param, t) (misc.jl:34-37)."""
    if _is_add_id(prob.pump):
        return additiveIdentity
    grid = prob.direct_grid()
    pts = _mesh(grid)
    shape = tuple(len(g) for g in reversed(grid))
    return _broadcast_kind(prob.pump(pts, prob.param, t), shape)


# --------------------------------------------------------------------------------------------
# the fused point-wise kernel -- kernels.jl:37-54
# --------------------------------------------------------------------------------------------


def muladd(fields, exp_dt, F_next, F_now, dt, nonlinearity, noise_func, xi, param, point):
    """One application of `muladd_kernel!` over the whole grid.

    fields : list of M arrays (batch..., spatial...)
    exp_dt : table kind (broadcast over batch dims because tables index only leading dims,
             kernels.jl:27)
    dt     : δt; `False` in the k-space call (strang_splitting.jl:73) => noise and G vanish.
    """
    dtype = fields[0].dtype
    f = SVector(list(fields))
    if _is_add_id(noise_func) or dt is False:
        noise = additiveIdentity  # √false = 0 multiplies the noise away (kernels.jl:42)
    else:
        sq = np.sqrt(dt)
        noise = _mul(-1j * sq, noise_func(f, point, param), SVector(list(xi)))    # kernels.jl:42
    if dt is False:
        exp_val = exp_dt                                           # cis(-0*G) = 1
        Fn = Fo = additiveIdentity
    else:
        G = nonlinearity(f, param) if not _is_add_id(nonlinearity) else additiveIdentity
        exp_val = _mul(_cis(_neg_scale(dt, G)), exp_dt)            # kernels.jl:44
        Fn = _mul(dt / 2, F_next)                                  # kernels.jl:45
        Fo = _mul(dt / 2, F_now)                                   # kernels.jl:46
    result = _add(_mul(exp_val, _add(f, Fo)), Fn)                  # kernels.jl:48
    result = _add(result, noise)                                   # kernels.jl:49
    if isinstance(result, SMatrix):
        raise TypeError("result must be a vector of fields")
    out = []
    for n in range(len(fields)):                                   # kernels.jl:51-53 (store converts)
        out.append(np.asarray(result[n]).astype(dtype, copy=False) if isinstance(result, SVector)
                   else np.asarray(result).astype(dtype, copy=False))
    shape = fields[0].shape
    return [np.ascontiguousarray(np.broadcast_to(o, shape)) for o in out]


# --------------------------------------------------------------------------------------------
# Strang splitting iterator -- strang_splitting.jl:9-90
# --------------------------------------------------------------------------------------------


class StrangSplitting:
    pass


class _LazyNoisePoints:
    """Q2 points, built on first element access: closures that ignore `r` (every shipped one) never trigger the
    reference's out-of-bounds case for n1 > n2."""

    def __init__(self, axes):
        self.axes, self._pts = axes, None

    def _get(self):
        if self._pts is None:
            self._pts = _noise_points(self.axes)
        return self._pts

    def __getitem__(self, i):
        return self._get()[i]

    def __iter__(self):
        return iter(self._get())

    def __len__(self):
        return len(self.axes)


class StrangSplittingIterator:
    def __init__(self, prob, tspan, *, dt, nsaves, save_start=True, noise_source=None,
                 record_noise=None, fft_workers=None):
        """`init` (strang_splitting.jl:32-67).

        noise_source(shape, dtype) -> ξ array with <|ξ|²>=1 (complex) or <ξ²>=1 (real); called
        once per prototype array per real-space half-step in the reference's draw order
        (misc.jl:44-51; strang_splitting.jl:80).  record_noise: optional list that receives every
        drawn ξ (so a test can feed the identical buffer to the GPU path)."""
        self.prob = prob
        self.dt, self.ts, self.steps_per_save = resolve_fixed_timestepping(dt, tspan, nsaves)  # :45
        self.save_start = bool(save_start)
        self.nsaves = nsaves
        self.u = [x.copy() for x in prob.u0]                        # :48
        self.rg = prob.reciprocal_grid()                            # :51
        self.dg = prob.direct_grid()                                # :52
        self.exp_Ddt = get_exponential(prob.dispersion, self.rg, prob.param, self.dt)       # :53
        self.exp_Vdt = get_exponential(prob.potential, self.dg, prob.param, self.dt / 2)    # :54
        self.pump_next = evaluate_pump(prob, _julia_float(tspan[0]))                         # :58
        self.pump_now = self.pump_next
        self.noise_source = noise_source
        self.record_noise = record_noise
        self.fft_workers = fft_workers
        d = prob.ndim
        self.axes = tuple(range(-d, 0))
        self.result = [np.stack([x] * (nsaves + self.save_start), axis=0) for x in prob.u0]  # :41-43
        self._point_direct = _mesh(self.dg)
        self._point_recip = _mesh(self.rg)
        self._point_noise = None if _is_add_id(prob.position_noise_func) else _LazyNoisePoints(self.dg)

    # misc.jl:44-51
    def _sample_noise(self):
        prob = self.prob
        if _is_add_id(prob.position_noise_func):
            return None
        protos = prob.noise_prototype
        xi = []
        for x in protos:
            z = self.noise_source(x.shape, x.dtype)
            z = np.asarray(z, dtype=x.dtype)
            if self.record_noise is not None:
                self.record_noise.append(z)
            xi.append(z)
        return xi

    # strang_splitting.jl:78-84
    def potential_pump_step(self, t, dt):
        prob = self.prob
        xi = self._sample_noise()                                   # :80
        if not _is_add_id(prob.pump):                               # :81 / misc.jl:39-42
            self.pump_now = self.pump_next
            self.pump_next = evaluate_pump(prob, t)
        self.u = muladd(self.u, self.exp_Vdt, self.pump_next, self.pump_now, dt, prob.nonlinearity,
                        prob.position_noise_func, xi, prob.param,
                        self._point_noise if self._point_noise is not None else self._point_direct)  # :82-83 (Q2)

    # strang_splitting.jl:69-76
    def diffusion_step(self):
        if self.fft_workers:
            import scipy.fft as sfft
            ft = [sfft.fftn(x, axes=self.axes, workers=self.fft_workers) for x in self.u]
        else:
            ft = [np.fft.fftn(x, axes=self.axes) for x in self.u]   # :72 (forward, unnormalised)
        ft = [f.astype(x.dtype, copy=False) for f, x in zip(ft, self.u)]
        ft = muladd(ft, self.exp_Ddt, additiveIdentity, additiveIdentity, False, additiveIdentity,
                    additiveIdentity, None, self.prob.param, self._point_recip)              # :73-74
        if self.fft_workers:
            import scipy.fft as sfft
            out = [sfft.ifftn(f, axes=self.axes, workers=self.fft_workers) for f in ft]
        else:
            out = [np.fft.ifftn(f, axes=self.axes) for f in ft]     # :75 (inverse, 1/N)
        self.u = [o.astype(x.dtype, copy=False) for o, x in zip(out, self.u)]

    # strang_splitting.jl:86-90
    def step(self, t, dt):
        self.potential_pump_step(t + dt / 2, dt / 2)
        self.diffusion_step()
        self.potential_pump_step(t + dt, dt / 2)

    # fixed_time_stepping.jl:26-54
    def solve(self):
        dt = self.dt
        ts = self.ts
        t = ts[0]
        off = 1 if self.save_start else 0
        for n in range(self.nsaves):
            for _ in range(self.steps_per_save):
                t = t + dt                                          # :44 (accumulated in T)
                self.step(t, dt)                                    # :45
            for r, x in zip(self.result, self.u):                   # :48
                r[n + off] = x
            ts[n + 1] = t                                           # :49
        return (ts[1 - off:], tuple(self.result))                   # :53


def solve(prob, alg, tspan, *, dt, nsaves, save_start=True, noise_source=None, record_noise=None,
          fft_workers=None, **_ignored):
    """`solve(prob, StrangSplitting(), tspan; dt, nsaves, ...)` (fixed_time_stepping.jl:79-81).
    Result arrays have NumPy shape (nsaves+save_start, batch..., n_d, ..., n_1) == Julia
    (n_1, ..., n_d, batch..., nsaves+save_start)."""
    it = StrangSplittingIterator(prob, tspan, dt=dt, nsaves=nsaves, save_start=save_start,
                                 noise_source=noise_source, record_noise=record_noise,
                                 fft_workers=fft_workers)
    return it.solve()


def pump_times(tspan, dt, nsaves):
    """The sequence of times at which the reference evaluates the pump:
    t0 at init, then for every step (t already incremented by dt) t+dt/2 and t+dt.
    Returns (t0, array of shape (nsteps, 2))."""
    _dt, ts, sps = resolve_fixed_timestepping(dt, tspan, nsaves)
    t = ts[0]
    out = np.empty((nsaves * sps, 2), dtype=ts.dtype)
    for i in range(nsaves * sps):
        t = t + _dt
        out[i, 0] = t + _dt / 2
        out[i, 1] = t + _dt
    return ts[0], out

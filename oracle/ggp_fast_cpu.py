"""Multithreaded CPU stand-in for timing the reference's CPU path -- TEST / BENCH INFRASTRUCTURE ONLY.

The reference runs its point-wise kernel on KernelAbstractions' multithreaded CPU backend and its transforms on
multithreaded FFTW (src/kernels.jl:37-54, src/misc.jl:53-64).  Neither Julia nor FFTW exist in this image, and the
line-by-line oracle (ggp_oracle.py) is a *checker*: single-threaded NumPy element-wise code around pocketfft, 5-7x
slower than what a tuned CPU library does on the same cores.  Timing the GPU path against that would flatter it, so
`bench.py`'s `cpu_baseline` / `--impl reference` legs use this module where it applies: the same algorithm, in the same
order, with torch's CPU kernels (MKL FFT, threaded element-wise ops) -- the strongest CPU implementation this image has.

Scope: exactly what the headline workload needs -- one component, scalar dispersion table, optional scalar potential
table, nonlinearity G = c + g|u|^2 (recognised by probing the closure), no pump, no noise.  Anything else: `supported()`
is False and the caller falls back to the oracle.  The tables come from the oracle itself (ggp_oracle.get_exponential,
i.e. src/misc.jl:12-20), and tests/test_fast_cpu.py checks this module against the oracle step for step.

Step (src/strang_splitting.jl:86-90, src/kernels.jl:44-49 with absent pump / noise):
    u <- cis(-dt/2 G(u)) exp_V u ;  u <- ifft(exp_D fft(u)) ;  u <- cis(-dt/2 G(u)) exp_V u
"""
import numpy as np

import ggp_oracle as O


def _fit_kerr(f, param):
    """G(u) = c + g |u|^2 from the closure, or None.  Least squares on random probes + held-out check."""
    rng = np.random.default_rng(0xFA57)

    def ev(u):
        v = f(O.SVector([u]), param)
        if isinstance(v, O.SMatrix):
            return None
        if isinstance(v, O.SVector):
            if len(v) != 1:
                return None
            v = v[0]
        return np.asarray(v, dtype=np.complex128) + np.zeros(u.shape, dtype=np.complex128)

    u = rng.uniform(0.2, 2.0, 16) * np.exp(2j * np.pi * rng.uniform(size=16))
    y = ev(u)
    if y is None:
        return None
    A = np.stack([np.ones(16), np.abs(u) ** 2], axis=1)
    coef, *_ = np.linalg.lstsq(A, y, rcond=None)
    v = rng.uniform(0.2, 2.0, 8) * np.exp(2j * np.pi * rng.uniform(size=8))
    pred = coef[0] + coef[1] * np.abs(v) ** 2
    truth = ev(v)
    if np.abs(pred - truth).max() > 1e-9 * max(1e-300, np.abs(truth).max(), np.abs(coef).max()):
        return None
    return complex(coef[0]), complex(coef[1])


def supported(prob):
    if len(prob.u0) != 1:
        return False
    if not O._is_add_id(prob.pump) or not O._is_add_id(prob.position_noise_func):
        return False
    if O._is_add_id(prob.dispersion):
        return False
    if not O._is_add_id(prob.nonlinearity) and _fit_kerr(prob.nonlinearity, prob.param) is None:
        return False
    return True


class FastStrang:
    """Same constructor contract as ggp_oracle.StrangSplittingIterator for the supported problems."""

    def __init__(self, prob, tspan, *, dt, nsaves, threads=None):
        import torch
        self.torch = torch
        if threads:
            torch.set_num_threads(int(threads))
        it = O.StrangSplittingIterator(prob, tspan, dt=dt, nsaves=nsaves)      # tables, dt resolution, grids: the oracle's
        self.dt, self.ts, self.steps_per_save = it.dt, it.ts, it.steps_per_save
        dtype = prob.u0[0].dtype
        self.real = np.float32 if dtype == np.complex64 else np.float64
        d = prob.ndim
        self.dims = tuple(range(-d, 0))

        def table(t):
            if t is None or O._is_mul_id(t):
                return None
            a = np.asarray(t[0] if isinstance(t, O.SVector) else t)
            return torch.from_numpy(np.ascontiguousarray(a.astype(dtype)))

        self.expD = table(it.exp_Ddt)
        self.expV = table(it.exp_Vdt)
        if self.expD is None:
            raise ValueError("FastStrang needs a scalar dispersion table")
        self.kerr = None
        if not O._is_add_id(prob.nonlinearity):
            c, g = _fit_kerr(prob.nonlinearity, prob.param)
            self.kerr = (c, g)
        self.u = torch.from_numpy(np.ascontiguousarray(prob.u0[0]).copy())

    def _half(self):
        torch, u = self.torch, self.u
        if self.kerr is not None:
            c, g = self.kerr
            h = self.real(self.dt / 2)
            a2 = u.real * u.real + u.imag * u.imag
            ang = (a2 * self.real(g.real) + self.real(c.real)) * (-h)                      # Re(-dt/2 G)
            if g.imag != 0 or c.imag != 0:                                                   # |cis(-dt/2 G)| = exp(dt/2 Im G)
                mod = torch.exp((a2 * self.real(g.imag) + self.real(c.imag)) * h)
            else:
                mod = torch.ones_like(ang)
            u = u * torch.polar(mod, ang)
        if self.expV is not None:
            u = u * self.expV
        self.u = u

    def step(self, t=None, dt=None):
        torch = self.torch
        self._half()
        f = torch.fft.fftn(self.u, dim=self.dims)
        f *= self.expD
        self.u = torch.fft.ifftn(f, dim=self.dims)
        self._half()

    def state(self):
        return self.u.numpy()

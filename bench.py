#!/usr/bin/env python
"""bench.py -- grid-point-steps/s of the Strang-splitting step on B200, next to the CPU reference path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5|c1] [--grid n]

A "step" is one Strang step (V/2 - D - V/2, src/strang_splitting.jl:86-90) of the whole synthetic field.  At N = 1
the workload is BASELINE.json configs[1] (C2: 2-D scalar Kerr GPE 2048^2 ComplexF32).  C2 does not shard ("replicas
only"), so at N > 1 `value` is N independent C2 grids (weak scaling); the two configs that DO shard -- the
trajectory-sharded Truncated-Wigner ensemble (configs[3]) and the slab-decomposed 3-D grid (configs[4], 512^3 and
1024^3) -- are reported in the top-level `sharded` block of every line, N = 1 included, with a checksum that must be
identical for every N.

How `value` is timed (all on the device, CUDA events on the plan's stream, nothing subtracted):
  value            K steps, ONE event pair around the kernels of each steady-state step (strided pass + the
                   contiguous-axis kernel that closes the step, programmatic-launch overlap between them included),
                   252 MiB overwritten BETWEEN the steps, outside every event pair => each step starts with a cold L2
                   (the timing rule for a working set -- 32 MiB -- smaller than the 126 MB L2).  value = points*K / sum.
  chained          the same K steps back to back, one event pair, no flush: what a 10 000-step production run sees
                   (state stays L2-resident).
  cold_kernel_sum  per-kernel event pairs with a flush after EVERY kernel (no intra-step L2 reuse, no launch overlap):
                   the per-kernel numbers the roofline block is computed from.

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "grid-point-steps/s"
L2_BYTES = 126 * 1024 * 1024


# --------------------------------------------------------------------------------------------------
# workloads (SURVEY §8d synthetic inputs)
# --------------------------------------------------------------------------------------------------
def make_workload(ns, name, nbatch=None, n=None, traj_range=None):
    import problems as P
    if name == "c2":
        N = n or 2048
        pb = P.kerr2d(ns, N=N, dtype=np.complex64, nsteps=1000)
        desc = f"C2: 2-D scalar Kerr GPE {N}^2 ComplexF32, D=|k|^2/2, g=+1, dt=1e-3 (BASELINE.json configs[1])"
        # algorithmic bytes per point: row kernel 2*c = 16; strided kernel 2*c (+ 8 for the exp_D table by SURVEY
        # §8d's figure -- never read here, the dispersion is a sum over axes and is held as two factors)
        return pb, dict(workload=desc, grid=[N, N], ncomp=1, nbatch=1, dtype="c64", points=N * N,
                        bytes_row=16, bytes_str=16, bytes_str_contract=24, b_alg_contract=56)
    if name == "c2_c128":
        N = n or 2048
        pb = P.kerr2d(ns, N=N, dtype=np.complex128, nsteps=1000)
        return pb, dict(workload=f"2-D scalar Kerr GPE {N}^2 ComplexF64", grid=[N, N], ncomp=1, nbatch=1,
                        dtype="c128", points=N * N, bytes_row=32, bytes_str=32, bytes_str_contract=48, b_alg_contract=112)
    if name == "c1":
        pb = P.quick_start(ns, kerr=True)
        return pb, dict(workload="C1: examples/quick_start.jl 128^2 ComplexF64 Kerr g=-6", grid=[128, 128],
                        ncomp=1, nbatch=1, dtype="c128", points=128 * 128, bytes_row=32, bytes_str=32,
                        bytes_str_contract=48, b_alg_contract=112)
    if name == "c3":
        N = n or 1024
        pb = P.exciton_polariton(ns, N=N, nsaves=1, tspan=(0, 100), dt=0.05, time_pump=True)
        return pb, dict(workload=f"C3: two-component exciton-polariton {N}^2 ComplexF64, 2x2 matrix-exponential "
                                 "dispersion, time-dependent pump (BASELINE.json configs[2])",
                        grid=[N, N], ncomp=2, nbatch=1, dtype="c128", points=N * N,
                        bytes_row=2 * 2 * 16 + 2 * 16, bytes_str=2 * 2 * 16 + 64, bytes_str_contract=2 * 2 * 16 + 64,
                        b_alg_contract=288)
    if name == "c3_c64":
        N = n or 1024
        pb = P.exciton_polariton(ns, N=N, nsaves=1, tspan=(0, 100), dt=0.05, time_pump=True, dtype=np.complex64)
        return pb, dict(workload=f"C3 in ComplexF32: two-component exciton-polariton {N}^2, 2x2 matrix-exponential "
                                 "dispersion, time-dependent pump",
                        grid=[N, N], ncomp=2, nbatch=1, dtype="c64", points=N * N,
                        bytes_row=2 * 2 * 8 + 2 * 8, bytes_str=2 * 2 * 8 + 32, bytes_str_contract=2 * 2 * 8 + 32,
                        b_alg_contract=144)
    if name == "c4":
        nb = nbatch or 4096
        pb = P.truncated_wigner(ns, ntraj=nb, N=256, ndim=2, dtype=np.complex128, tspan=(0, 20), dt=0.05,
                                traj_range=traj_range)
        return pb, dict(workload=f"C4: Truncated-Wigner ensemble, 256^2 polariton grid x {nb} trajectories "
                                 "ComplexF64, in-kernel Philox noise (BASELINE.json configs[3])",
                        grid=[256, 256], ncomp=1, nbatch=nb, dtype="c128", points=256 * 256 * nb,
                        bytes_row=32, bytes_str=32, bytes_str_contract=32, b_alg_contract=96)
    if name == "c5":
        N = n or 512
        rank, world = traj_range if traj_range is not None else (0, 1)
        pb = P.kerr3d_slab(ns, N=N, rank=rank, world=world, dtype=np.complex64)
        return pb, dict(workload=f"C5: 3-D BEC Kerr GPE {N}^3 ComplexF32, slab-decomposed over {world} GPU(s), the two "
                                 "all-to-all transposes fused into the FFT kernels as NVLink peer stores "
                                 "(BASELINE.json configs[4])",
                        grid=[N, N, N], ncomp=1, nbatch=1, dtype="c64", points=N * N * N // world,
                        bytes_row=16, bytes_str=16, bytes_str_contract=16, b_alg_contract=88)
    raise SystemExit(f"unknown workload {name}")


def config_of(meta, world=1):
    """The `config` object: identical in the GPU arm and in the reference arm (the driver compares them)."""
    w = meta["workload"]
    if w.startswith("C5"):
        w = w.replace(f"over {world} GPU(s)", "over the GPUs of the run")
    return dict(workload=w, grid=meta["grid"], ncomp=meta["ncomp"], nbatch=meta["nbatch"])


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, windows):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if not any(a - 0.05 <= ts <= b + 0.05 for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sample inside the timed region"],
                    "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(pw)), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU reference path (the oracle, timed on the host cores)
# --------------------------------------------------------------------------------------------------
def cpu_reference(workload, steps, warmup, budget_s):
    """Times the CPU restatement of the reference on the same workload, on all host threads.  kind = "port": the
    Julia/FFTW reference itself cannot run in this image (no Julia, no libfftw3).  Where it applies (one component,
    Kerr-type nonlinearity, no pump / noise: the headline C2) the step runs on torch's CPU kernels (MKL FFT + threaded
    element-wise ops, oracle/ggp_fast_cpu.py, validated against the oracle in tests/test_fast_cpu.py) -- 5-7x faster
    than the line-by-line NumPy oracle and the fairest stand-in for KernelAbstractions-CPU + FFTW this image offers;
    everything else is timed on the oracle itself (scipy.fft on all threads)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ggp_oracle as O
    cores = len(os.sched_getaffinity(0))
    for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):      # torchrun pins these to 1; the baseline uses every host core
        os.environ[var] = str(cores)
    name, nbatch = workload
    pb, meta = make_workload(O, name, nbatch=nbatch)
    prob = O.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    noise = None
    if name == "c4":
        rng = np.random.default_rng(1)

        def noise(shape, dtype):
            return ((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)).astype(dtype)
    it, how = None, f"oracle/ggp_oracle.py, scipy.fft workers={cores}"
    if not os.environ.get("GGP_CPU_ORACLE_ONLY"):
        try:
            import ggp_fast_cpu as F
            if F.supported(prob):
                it = F.FastStrang(prob, pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"], threads=cores)
                how = f"oracle/ggp_fast_cpu.py: torch CPU kernels, MKL FFT, {cores} threads"
        except Exception:
            it = None
    if it is None:
        it = O.StrangSplittingIterator(prob, pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"], noise_source=noise,
                                       fft_workers=cores)
    t = it.ts[0]
    t0 = time.perf_counter()
    t = t + it.dt
    it.step(t, it.dt)
    one = time.perf_counter() - t0
    w = max(0, min(warmup, int(0.2 * budget_s / max(one, 1e-9))))
    for _ in range(w):
        t = t + it.dt
        it.step(t, it.dt)
    k = max(1, min(steps, int(budget_s / max(one, 1e-9))))
    t0 = time.perf_counter()
    for _ in range(k):
        t = t + it.dt
        it.step(t, it.dt)
    el = time.perf_counter() - t0
    val = meta["points"] * k / el
    return dict(value=val, unit=METRIC, cores=cores, kind="port",
                sample=f"{k} Strang steps of the full {meta['workload'].split(':')[0]} field "
                       f"({how}), {el:.1f} s wall"), meta, k, el


# --------------------------------------------------------------------------------------------------
# GPU measurement
# --------------------------------------------------------------------------------------------------
def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "of fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel_class, workload):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(workload, {}).get(kernel_class)
    except Exception:
        return None


def measure(G, name, steps, warmup, device, nbatch=None, batch_offset=0, do_e2e=True, do_cold=True, comm=None,
            n=None, slab=None, keep=False):
    """Returns dict with chained / step-window / per-kernel / e2e numbers for one plan on this rank."""
    lib = G.lib.load()
    pb, meta = make_workload(G, name, nbatch=nbatch, n=n,
                             traj_range=(slab if name == "c5" else
                                         ((batch_offset, batch_offset + nbatch) if (name == "c4" and nbatch) else None)))
    prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    nsaves_steps = warmup + 6 * steps + 16
    dt = pb["dt"]
    tspan = (pb["tspan"][0], pb["tspan"][0] + type(dt)(4 * nsaves_steps) * dt)
    skw = dict(slab=slab, slab_local=True) if (slab is not None and slab[1] > 1) else {}
    it = G.init(prob, G.StrangSplitting(), tspan, dt=dt, nsaves=1, save_start=False, rng=1234, device=device,
                batch_offset=batch_offset, result_buffers=do_e2e, **skw)
    if comm is not None:
        comm(it)
    h = it.handle
    out = dict(meta=meta)
    windows = []
    it.advance(warmup)
    G.lib.check(lib.ggp_synchronize(h))
    l0 = lib.ggp_launch_count(h)
    # (1) chained: K back-to-back steps, state resident in HBM, one event pair on the launching stream
    barrier()
    tw0 = time.time()
    G.lib.check(lib.ggp_timer_begin(h))
    it.advance(steps)
    ms = C.c_float()
    G.lib.check(lib.ggp_timer_end(h, C.byref(ms)))
    windows.append((tw0, time.time()))
    out["chained_ms"] = float(ms.value)
    out["launches"] = int(lib.ggp_launch_count(h) - l0)
    # (2) step windows: one event pair per steady-state step; cold: 252 MiB overwritten between the steps
    for tag, flush in (("steps_warm", 0), ("steps_cold", 2 * L2_BYTES if do_cold else None)):
        if flush is None:
            continue
        G.lib.check(lib.ggp_profile_steps_enable(h, 1, flush))
        barrier()
        tw0 = time.time()
        it.advance(steps)
        tot, cnt = C.c_double(), C.c_int64()
        G.lib.check(lib.ggp_profile_steps_read(h, C.byref(tot), C.byref(cnt)))
        windows.append((tw0, time.time()))
        G.lib.check(lib.ggp_profile_steps_enable(h, 0, 0))
        out[tag] = dict(ms=float(tot.value), steps=int(cnt.value))
    # (3) per-kernel events, warm and with a flush after every kernel
    for tag, flush in (("prof", 0), ("flush", 2 * L2_BYTES if do_cold else None)):
        if flush is None:
            continue
        G.lib.check(lib.ggp_debug_l2_flush(h, flush))
        G.lib.check(lib.ggp_profile_enable(h, 1))
        tw0 = time.time()
        it.advance(steps)
        tot = (C.c_double * 4)()
        cnt = (C.c_int64 * 4)()
        G.lib.check(lib.ggp_profile_read(h, tot, cnt))
        windows.append((tw0, time.time()))
        G.lib.check(lib.ggp_profile_enable(h, 0))
        G.lib.check(lib.ggp_debug_l2_flush(h, 0))
        out[tag] = dict(ms=[float(x) for x in tot], n=[int(x) for x in cnt])
    # (4) end to end through the public API: pinned host u0 -> device, `solve!`-style stepping, device -> pinned host
    if do_e2e:
        u0 = [np.array(x) for x in pb["u0"]]
        sps = min(steps, 1000)
        it.upload(u0)                       # the page-locked input buffers now hold u0
        it.advance(min(warmup, 5)); it.save_async(0); it.save_wait(); it.fetch()   # warm: snapshot buffer, copy stream
        for dst, src in zip(it.u, u0):
            np.copyto(dst, src)
        barrier()
        tw0 = time.time()
        t0 = time.perf_counter()
        it.upload()                         # H2D of the inputs from page-locked host memory
        done, d2h = 0, 0
        while done < steps:
            k = min(sps, steps - done)
            it.advance(k)
            it.save_async(0)          # streaming save: device snapshot + PCIe transfer overlapped with the next interval
            done += k
            d2h += 1
        it.save_wait()                # the last saved state has landed in page-locked host memory
        el = time.perf_counter() - t0
        windows.append((tw0, time.time()))
        sbytes = sum(x.nbytes for x in u0)
        out["e2e"] = dict(seconds=el, h2d_bytes_per_step=sbytes / steps, d2h_bytes_per_step=sbytes * d2h / steps,
                          saves=d2h)
    out["windows"] = windows
    out["iter"] = it
    return out


_DIST = None


def barrier():
    if _DIST is not None:
        import torch
        torch.cuda.synchronize()
        _DIST.barrier()


def allmax(x):
    if _DIST is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    _DIST.all_reduce(t, op=_DIST.ReduceOp.MAX)
    return float(t.item())


def allsum(x):
    if _DIST is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    _DIST.all_reduce(t, op=_DIST.ReduceOp.SUM)
    return float(t.item())


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints to fd 1 (e.g. NCCL's
    version banner) has been redirected to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def committed_n1():
    """Builder-measured N = 1 figures of the sharded configs (profiles/r02_sharded_n1.json), the denominator of the
    efficiencies printed at N > 1 (the driver computes its own from the per-N lines)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_sharded_n1.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def slab_checksum(it, rank, world, N):
    """Position-sensitive checksum of a slab-decomposed field: sum |u|^2 w(x, y, z) with fixed global weights, summed
    over ranks.  A misplaced block of the fused transposes changes it; the norm alone would not (a permutation keeps
    it).  Must be identical (to fp64 summation order) for every number of GPUs."""
    u = it.fetch()[0]                                   # local z-slab (n3/world, n2, n1)
    nz = u.shape[0]
    z = (rank * nz + np.arange(nz, dtype=np.float64))[:, None, None]
    y = np.arange(u.shape[1], dtype=np.float64)[None, :, None]
    x = np.arange(u.shape[2], dtype=np.float64)[None, None, :]
    acc = 0.0
    for k in range(nz):                                 # plane by plane: no 8 GiB temporaries
        a2 = np.abs(u[k].astype(np.complex128)) ** 2
        w = 1.0 + 0.5 * np.sin(0.37 * x[0] + 0.11) * np.cos(0.23 * y[0] - 0.4) + 0.25 * np.sin(0.19 * z[k, 0, 0] + 0.7)
        acc += float((a2 * w).sum())
    return allsum(acc)


def sharded_block(G, a, world, rank, local, peak, windows):
    """The two configs that shard, at this run's N: value, time per step, fraction of the contract HBM
    roofline per GPU, per-kernel times, checksum."""
    out = {}
    n1 = committed_n1()
    lib = G.lib.load()

    def eff(key, value):
        ref = n1.get(key, {}).get("value")
        return dict(efficiency_vs_n1=(value / (world * ref)) if (ref and world > 1) else (1.0 if world == 1 else None),
                    n1_value=ref, n1_source="profiles/r02_sharded_n1.json (builder-measured, same bench code)")

    # ---- C4: trajectory-sharded Truncated-Wigner ensemble, strong scaling over a fixed 4096 trajectories
    try:
        nb = a.nbatch // world
        ksteps = 20
        r4 = measure(G, "c4", ksteps, 3, local, nbatch=nb, batch_offset=rank * nb, do_e2e=False, do_cold=False,
                     comm=(lambda it: attach_comm(G, it, world, rank)) if world > 1 else None)
        it4 = r4["iter"]
        barrier()
        t0 = time.perf_counter()
        nk = it4.observe(G.lib.OBS_MOMENTUM)      # all-reduced over ranks when world > 1
        obs_s = allmax(time.perf_counter() - t0)
        ms4 = allmax(r4["chained_ms"])
        pts4 = allsum(float(r4["meta"]["points"]))
        v4 = pts4 * ksteps / (ms4 * 1e-3)
        out["c4_ensemble"] = dict(
            workload=r4["meta"]["workload"].replace(f"x {nb} ", f"x {a.nbatch} "), scaling="strong",
            ntraj_total=nb * world, ntraj_per_gpu=nb, value=v4, unit=METRIC, ms_per_step=ms4 / ksteps, steps=ksteps,
            frac_of_hbm_roofline_contract_per_gpu=96 * pts4 / world / (ms4 / ksteps * 1e-3) / 1e9 / peak,
            collective="none per step; ncclAllReduce of n(k) (512 KiB) per observation",
            observable_allreduce_s=obs_s, checksum_nk_sum=float(nk.sum()),
            per_kernel_ms=dict(row=r4["prof"]["ms"][0] / max(1, r4["prof"]["n"][0]),
                               str=r4["prof"]["ms"][1] / max(1, r4["prof"]["n"][1])), **eff("c4_ensemble", v4))
        windows += r4["windows"]
        it4.close()
    except Exception as e:  # the sharded block must never take the headline down
        out["c4_ensemble"] = dict(error=repr(e))
    # ---- C5: slab-decomposed 3-D grid, strong scaling; 1024^3 is BASELINE.json's size
    for N in [int(x) for x in a.c5_grids.split(",") if x]:
        key = f"c5_slab_{N}"
        try:
            k5 = 20 if N <= 512 else 10
            r5 = measure(G, "c5", k5, 3, local, n=N, slab=(rank, world), do_e2e=False, do_cold=False,
                         comm=(lambda it: attach_comm(G, it, world, rank)) if world > 1 else None)
            ms5 = allmax(r5["chained_ms"])
            pts5 = allsum(float(r5["meta"]["points"]))
            v5 = pts5 * k5 / (ms5 * 1e-3)
            chk = slab_checksum(r5["iter"], rank, world, N)
            out[key] = dict(workload=r5["meta"]["workload"], scaling="strong", value=v5, unit=METRIC,
                            ms_per_step=ms5 / k5, steps=k5,
                            frac_of_hbm_roofline_contract_per_gpu=88 * pts5 / world / (ms5 / k5 * 1e-3) / 1e9 / peak,
                            collective="2 all-to-all transposes per step, fused into the FFT kernels as NVLink peer stores",
                            checksum_weighted_norm=chk,
                            per_kernel_ms=dict(row=r5["prof"]["ms"][0] / max(1, r5["prof"]["n"][0]),
                                               str_d=r5["prof"]["ms"][1] / max(1, r5["prof"]["n"][1]),
                                               str_fi=r5["prof"]["ms"][2] / max(1, r5["prof"]["n"][2])), **eff(key, v5))
            windows += r5["windows"]
            r5["iter"].close()
        except Exception as e:
            out[key] = dict(error=repr(e))
    return out


def main():
    global _DIST, _REAL_STDOUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--grid", dest="n", type=int, default=None, help="grid edge override (size sweep)")
    ap.add_argument("--nbatch", type=int, default=4096, help="total trajectories of the c4 ensemble")
    ap.add_argument("--c5-grids", default="512,1024", help="grid edges of the slab-decomposed 3-D runs in `sharded`")
    ap.add_argument("--sizes", default="1024,4096,8192", help="2-D size sweep reported in extra.size_sweep (N = 1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the sharded block, the size sweep and C3")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if a.gpus > 1 and world == 1 and a.impl == "ours":
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))

    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)       # (after the re-launch above: the torchrun children must inherit the real stdout)
    os.dup2(2, 1)

    if a.impl == "reference":
        if rank != 0:
            return
        cb, meta, k, el = cpu_reference((a.workload, None), a.steps, a.warmup, budget_s=90.0)
        line = dict(impl="reference", metric=METRIC, value=cb["value"], unit=METRIC, n_gpus=a.gpus, steps=k,
                    warmup=a.warmup, ms_per_step=1e3 * el / k, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype=meta["dtype"], data="synthetic", config=config_of(meta, max(a.gpus, 1)),
                    note="reference arm = CPU restatement of the reference on all host threads (" + cb["sample"] +
                         "); the Julia/FFTW reference cannot run in this image",
                    cpu_baseline=cb,
                    e2e=dict(value=cb["value"], unit=METRIC, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0)
        emit(line)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        _DIST = dist
    import ggp_b200 as G
    lib = G.lib.load()

    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("GGP_BENCH_NO_CLOCKS"):      # (A/B switch: does the polling disturb the timing?)
        sampler.start()

    # ---- headline workload ---------------------------------------------------------------------
    res = measure(G, a.workload, a.steps, a.warmup, local, n=a.n,
                  nbatch=(a.nbatch // world if a.workload == "c4" else None),
                  batch_offset=(rank * (a.nbatch // world) if a.workload == "c4" else 0),
                  slab=((rank, world) if a.workload == "c5" else None),
                  comm=((lambda it: attach_comm(G, it, world, rank)) if (a.workload == "c5" and world > 1) else None),
                  do_e2e=(a.workload != "c5"))
    if "e2e" not in res:
        res["e2e"] = dict(seconds=float("nan"), h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    meta = res["meta"]
    pts = meta["points"]
    total_pts = allsum(float(pts))
    chained_ms = allmax(res["chained_ms"])
    sc = res.get("steps_cold") or res["steps_warm"]
    cold_ms_per_step = allmax(sc["ms"] / max(1, sc["steps"]))
    warm_ms_per_step = allmax(res["steps_warm"]["ms"] / max(1, res["steps_warm"]["steps"]))
    fl = res.get("flush") or res["prof"]
    ker_ms_flush = allmax(sum(fl["ms"]))
    ker_ms_warm = allmax(sum(res["prof"]["ms"]))
    e2e_s = allmax(res["e2e"]["seconds"])
    peak, peak_note = peaks()
    # dominant kernel (cold-L2 per-kernel events)
    names = ["row_kernel (inverse FFT_x + V/2 + V/2 + forward FFT_x)", "str_kernel (FFT_y x exp_D x inverse FFT_y)",
             "str_kernel forward-only / inverse-only", "oned_kernel"]
    bytes_per_pt = [meta["bytes_row"], meta["bytes_str"], meta["bytes_row"], meta["bytes_row"] + meta["bytes_str"]]
    dom = int(np.argmax(fl["ms"]))
    avg_ms = fl["ms"][dom] / max(1, fl["n"][dom])
    alg_bytes = bytes_per_pt[dom] * pts
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    warm_avg = res["prof"]["ms"][dom] / max(1, res["prof"]["n"][dom])
    contract_bytes = (meta["bytes_str_contract"] if dom == 1 else bytes_per_pt[dom]) * pts
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                    traffic=ncu_traffic("row" if dom == 0 else "str", a.workload), peak_source=peak_note,
                    kernel=names[dom], alg_bytes_per_launch=alg_bytes, avg_launch_ms=avg_ms,
                    alg_bytes_note="bytes the kernel must move: state read + write (the exp_D table is never read: "
                                   "separable dispersion held as two factors)" if dom == 1 and meta["bytes_str"] != meta["bytes_str_contract"] else "state read + write + tables",
                    frac_with_contract_table_share=contract_bytes / (avg_ms * 1e-3) / 1e9 / peak,
                    l2="cold: 252 MiB overwritten after every kernel (per-kernel event pairs)",
                    warm_l2=dict(avg_launch_ms=warm_avg, achieved=alg_bytes / (warm_avg * 1e-3) / 1e9),
                    per_kernel_ms=dict(row=fl["ms"][0] / max(1, fl["n"][0]), str=fl["ms"][1] / max(1, fl["n"][1])),
                    step_contract=dict(b_alg_bytes_per_point=meta["b_alg_contract"],
                                       frac_cold=meta["b_alg_contract"] * pts / (cold_ms_per_step * 1e-3) / 1e9 / peak,
                                       frac_chained=meta["b_alg_contract"] * pts / (chained_ms / a.steps * 1e-3) / 1e9 / peak))
    value_cold = total_pts / (cold_ms_per_step * 1e-3)
    value_chained = total_pts * a.steps / (chained_ms * 1e-3)
    line = dict(metric=METRIC, value=value_cold, unit=METRIC, n_gpus=world, steps=a.steps, warmup=a.warmup,
                ms_per_step=cold_ms_per_step, higher_is_better=True,
                scaling="strong" if a.workload in ("c4", "c5") else "weak", vs_baseline=None, dtype=meta["dtype"],
                data="synthetic", config=config_of(meta, world),
                parallelism=("single GPU" if world == 1 else
                             (f"{world} GPUs, trajectories sharded {a.nbatch // world}/GPU" if a.workload == "c4"
                              else (f"{world} GPUs, z-slabs" if a.workload == "c5"
                                    else f"{world} independent replicas (C2 does not shard)"))),
                headline="value = points*K / sum over the K steps of one CUDA-event pair per steady-state step (strided "
                         "kernel + the contiguous-axis kernel closing the step); 252 MiB overwritten BETWEEN the steps, "
                         "outside the event pairs, so every step starts L2-cold; nothing is subtracted",
                chained=dict(value=value_chained, ms_per_step=chained_ms / a.steps,
                             what="K steps back to back, one event pair, no flush (state L2-resident as in production)"),
                step_windows_warm=dict(value=total_pts / (warm_ms_per_step * 1e-3), ms_per_step=warm_ms_per_step),
                cold_kernel_sum=dict(value=total_pts * a.steps / (ker_ms_flush * 1e-3), ms_per_step=ker_ms_flush / a.steps,
                                     what="per-kernel event pairs, L2 flushed after EVERY kernel"),
                warm_kernel_sum=dict(value=total_pts * a.steps / (ker_ms_warm * 1e-3), ms_per_step=ker_ms_warm / a.steps),
                roofline=roofline,
                e2e=dict(value=total_pts * a.steps / e2e_s, unit=METRIC,
                         h2d_bytes_per_step=res["e2e"]["h2d_bytes_per_step"],
                         d2h_bytes_per_step=res["e2e"]["d2h_bytes_per_step"],
                         what="u0 in page-locked host memory -> ggp_set_state (H2D), ggp_step in save intervals of <=1000 "
                              "steps, ggp_save_async of the state to page-locked host memory after each + final "
                              "ggp_save_wait (solve! of the host interface); plan creation excluded"),
                gpu_launches=res["launches"])
    windows = list(res["windows"])
    res["iter"].close()

    # ---- the configs that shard, the size sweep, C3 ----------------------------------------------
    extra = {}
    if not a.no_extra and a.workload == "c2":
        line["sharded"] = sharded_block(G, a, world, rank, local, peak, windows)
        if world == 1:
            sweep = {}
            for N in [int(x) for x in a.sizes.split(",") if x]:
                try:
                    k = 200 if N <= 2048 else (60 if N <= 4096 else 20)
                    r = measure(G, "c2", k, 5, local, n=N, do_e2e=False)
                    cold = r["steps_cold"]["ms"] / r["steps_cold"]["steps"]
                    p2 = r["meta"]["points"]
                    sweep[str(N)] = dict(cold_ms_per_step=cold, chained_ms_per_step=r["chained_ms"] / k,
                                         frac_cold=56 * p2 / (cold * 1e-3) / 1e9 / peak,
                                         frac_chained=56 * p2 / (r["chained_ms"] / k * 1e-3) / 1e9 / peak,
                                         per_kernel_cold_ms=dict(row=r["flush"]["ms"][0] / max(1, r["flush"]["n"][0]),
                                                                 str=r["flush"]["ms"][1] / max(1, r["flush"]["n"][1])))
                    windows += r["windows"]
                    r["iter"].close()
                except Exception as e:
                    sweep[str(N)] = dict(error=repr(e))
            extra["size_sweep_c64"] = dict(what="C2's problem at other grid sizes, same timing as the headline; "
                                                "frac = 56 B/point contract / time / measured HBM peak", **sweep)
            for nm, key, b in (("c3", "c3", 288), ("c3_c64", "c3_c64", 144)):
                try:
                    r3 = measure(G, nm, 200, 5, local, do_e2e=False, do_cold=False)
                    ms3 = r3["chained_ms"]
                    extra[key] = dict(workload=r3["meta"]["workload"], value=r3["meta"]["points"] * 200 / (ms3 * 1e-3),
                                      unit=METRIC, ms_per_step=ms3 / 200,
                                      frac_of_hbm_roofline_contract=b * r3["meta"]["points"] / (ms3 / 200 * 1e-3) / 1e9 / peak,
                                      per_kernel_ms=dict(row=r3["prof"]["ms"][0] / max(1, r3["prof"]["n"][0]),
                                                         str=r3["prof"]["ms"][1] / max(1, r3["prof"]["n"][1])))
                    windows += r3["windows"]
                    r3["iter"].close()
                except Exception as e:
                    extra[key] = dict(error=repr(e))
            # the generic plan (axes of any length: Bluestein; one kernel per stage): the price of leaving the fused kernels
            try:
                rg = measure(G, "c2", 50, 5, local, n=1000, do_e2e=False, do_cold=False)
                msg = rg["chained_ms"] / 50
                extra["generic_plan_1000"] = dict(
                    workload="C2's problem on a 1000^2 grid (not a power of two): generic plan, Bluestein transforms of "
                             "length 2048, 2d + 2 sweeps per step", value=rg["meta"]["points"] * 50 / (rg["chained_ms"] * 1e-3),
                    unit=METRIC, ms_per_step=msg, frac_of_hbm_roofline_contract=56 * rg["meta"]["points"] / (msg * 1e-3) / 1e9 / peak)
                windows += rg["windows"]
                rg["iter"].close()
            except Exception as e:
                extra["generic_plan_1000"] = dict(error=repr(e))
    if extra:
        line["extra"] = extra

    if rank == 0:
        line["clocks"] = sampler.stop(windows)
        if not a.no_cpu and world == 1:           # reported baseline: rank 0 at N = 1 only
            try:
                cb, _, _, _ = cpu_reference((a.workload, None if a.workload != "c4" else 8), 10 ** 9, 1, a.cpu_budget)
                line["cpu_baseline"] = cb
            except Exception as e:
                line["cpu_baseline"] = dict(value=None, unit=METRIC, cores=0, kind="port", sample=f"failed: {e!r}")
        emit(line)
    if world > 1:
        _DIST.barrier()
        _DIST.destroy_process_group()


def attach_comm(G, it, world, rank):
    """NCCL communicator for the all-reduce of the saved observables: the 128-byte unique id is
    made by rank 0 and broadcast with torch.distributed (plumbing only)."""
    import torch
    lib = G.lib.load()
    uid = (C.c_char * 128)()
    if rank == 0:
        G.lib.check(lib.ggp_comm_unique_id(uid))
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device="cuda")
    _DIST.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    buf = C.create_string_buffer(raw, 128)
    G.lib.check(lib.ggp_comm_init(it.handle, world, rank, buf))
    if it.slab is not None and not os.environ.get("GGP_SLAB_NCCL"):
        # slab plans: peer-memory exchange fused into the FFT kernels (CUDA IPC handles gathered here)
        G.parallel.attach_p2p(G, it, _DIST)
    if it.slab is None:
        it.observe(G.lib.OBS_NORM)      # warm-up collective: NCCL sets its channels up lazily on first use


if __name__ == "__main__":
    main()

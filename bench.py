#!/usr/bin/env python
"""bench.py -- grid-point-steps/s of the Strang-splitting step on B200, next to the CPU reference path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c1]

A "step" is one Strang step (V/2 - D - V/2, src/strang_splitting.jl:86-90) of the whole synthetic
field.  At N = 1 the workload is BASELINE.json configs[1] (C2: 2-D scalar Kerr GPE 2048^2 ComplexF32).
C2 does not shard ("replicas only"), so at N > 1 `value` is N independent C2 grids (weak
scaling), and the trajectory-sharded Truncated-Wigner ensemble (configs[3], strong scaling over a
fixed 4096 trajectories, NCCL all-reduce of the observable) is reported in `extra.c4_ensemble` of
every line, N = 1 included, so its scaling can be read off the same files.

One JSON line on stdout (rank 0).  Keys follow the driver contract plus `roofline`, `cpu_baseline`,
`e2e`, `gpu_launches`, `clocks`.
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
        return pb, dict(workload=desc, grid=[N, N], ncomp=1, nbatch=1, dtype="c64", points=N * N,
                        bytes_row=16, bytes_str=24, b_alg_contract=56)
    if name == "c2_c128":
        N = n or 2048
        pb = P.kerr2d(ns, N=N, dtype=np.complex128, nsteps=1000)
        return pb, dict(workload=f"2-D scalar Kerr GPE {N}^2 ComplexF64", grid=[N, N], ncomp=1, nbatch=1,
                        dtype="c128", points=N * N, bytes_row=32, bytes_str=48, b_alg_contract=112)
    if name == "c1":
        pb = P.quick_start(ns, kerr=True)
        return pb, dict(workload="C1: examples/quick_start.jl 128^2 ComplexF64 Kerr g=-6", grid=[128, 128],
                        ncomp=1, nbatch=1, dtype="c128", points=128 * 128, bytes_row=32, bytes_str=48,
                        b_alg_contract=112)
    if name == "c3":
        N = n or 1024
        pb = P.exciton_polariton(ns, N=N, nsaves=1, tspan=(0, 100), dt=0.05, time_pump=True)
        return pb, dict(workload=f"C3: two-component exciton-polariton {N}^2 ComplexF64, 2x2 matrix-exponential "
                                 "dispersion, time-dependent pump (BASELINE.json configs[2])",
                        grid=[N, N], ncomp=2, nbatch=1, dtype="c128", points=N * N,
                        bytes_row=2 * 2 * 16 + 2 * 16, bytes_str=2 * 2 * 16 + 64, b_alg_contract=288)
    if name == "c3_c64":
        N = n or 1024
        pb = P.exciton_polariton(ns, N=N, nsaves=1, tspan=(0, 100), dt=0.05, time_pump=True, dtype=np.complex64)
        return pb, dict(workload=f"C3 in ComplexF32: two-component exciton-polariton {N}^2, 2x2 matrix-exponential "
                                 "dispersion, time-dependent pump",
                        grid=[N, N], ncomp=2, nbatch=1, dtype="c64", points=N * N,
                        bytes_row=2 * 2 * 8 + 2 * 8, bytes_str=2 * 2 * 8 + 32, b_alg_contract=144)
    if name == "c4":
        nb = nbatch or 4096
        pb = P.truncated_wigner(ns, ntraj=nb, N=256, ndim=2, dtype=np.complex128, tspan=(0, 20), dt=0.05,
                                traj_range=traj_range)
        return pb, dict(workload=f"C4: Truncated-Wigner ensemble, 256^2 polariton grid x {nb} trajectories "
                                 "ComplexF64, in-kernel Philox noise (BASELINE.json configs[3])",
                        grid=[256, 256], ncomp=1, nbatch=nb, dtype="c128", points=256 * 256 * nb,
                        bytes_row=32, bytes_str=32, b_alg_contract=96)
    if name == "c5":
        N = n or 512
        rank, world = traj_range if traj_range is not None else (0, 1)
        pb = P.kerr3d_slab(ns, N=N, rank=rank, world=world, dtype=np.complex64)
        return pb, dict(workload=f"C5: 3-D BEC Kerr GPE {N}^3 ComplexF32, slab-decomposed over {world} GPU(s), all-to-all transposes fused into the FFT kernels as NVLink peer stores "
                                 "all-to-all transpose (BASELINE.json configs[4]; 1024^3 with --grid 1024)",
                        grid=[N, N, N], ncomp=1, nbatch=1, dtype="c64", points=N * N * N // world,
                        bytes_row=16, bytes_str=16, b_alg_contract=88)
    raise SystemExit(f"unknown workload {name}")


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


def measure(G, name, steps, warmup, device, nbatch=None, batch_offset=0, do_e2e=True, do_flush=True, comm=None,
            n=None, slab=None):
    """Returns dict with chained / flushed / per-kernel / e2e numbers for one plan on this rank."""
    lib = G.lib.load()
    pb, meta = make_workload(G, name, nbatch=nbatch, n=n,
                             traj_range=(slab if name == "c5" else
                                         ((batch_offset, batch_offset + nbatch) if (name == "c4" and nbatch) else None)))
    prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    nsaves_steps = steps + warmup + steps + steps + steps + 8
    dt = pb["dt"]
    tspan = (pb["tspan"][0], pb["tspan"][0] + type(dt)(4 * nsaves_steps) * dt)
    skw = dict(slab=slab, slab_local=True) if (slab is not None and slab[1] > 1) else {}
    it = G.init(prob, G.StrangSplitting(), tspan, dt=dt, nsaves=1, save_start=False, rng=1234, device=device,
                batch_offset=batch_offset, **skw)
    if comm is not None:
        comm(it)
    h = it.handle
    out = dict(meta=meta)
    windows = []
    # warm-up
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
    # (2) per-kernel events (+ optional L2 flush after every kernel)
    for tag, flush in (("prof", 0), ("flush", 2 * L2_BYTES if do_flush else None)):
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
    # (3) end to end through the public API: pinned host u0 -> device, `solve!`-style stepping, device -> host
    if do_e2e:
        u0 = [np.array(x) for x in pb["u0"]]
        sps = min(steps, 1000)
        it.upload(u0)
        it.advance(min(warmup, 5)); it.fetch()
        barrier()
        tw0 = time.time()
        t0 = time.perf_counter()
        it.upload(u0)
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
    # (2b) cold-L2 bracket: K steps with a flush after every kernel inside ONE event pair, minus the same
    #      number of flushes timed alone (per-kernel event pairs add ~2-5 us of gap per launch)
    if do_flush:
        G.lib.check(lib.ggp_debug_l2_flush(h, 2 * L2_BYTES))
        l1 = lib.ggp_launch_count(h)
        tw0 = time.time()
        G.lib.check(lib.ggp_timer_begin(h))
        it.advance(steps)
        ms = C.c_float()
        G.lib.check(lib.ggp_timer_end(h, C.byref(ms)))
        nflush = int(lib.ggp_launch_count(h) - l1)
        msf = C.c_float()
        G.lib.check(lib.ggp_debug_flush_only(h, nflush, C.byref(msf)))
        windows.append((tw0, time.time()))
        G.lib.check(lib.ggp_debug_l2_flush(h, 0))
        out["cold_bracket_ms"] = float(ms.value) - float(msf.value)
        out["cold_bracket_detail"] = dict(total_ms=float(ms.value), flush_only_ms=float(msf.value), flushes=nflush)
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


def main():
    global _DIST, _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--grid", dest="n", type=int, default=None, help="grid edge override (size sweep)")
    ap.add_argument("--nbatch", type=int, default=4096, help="total trajectories of the c4 ensemble")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
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

    if a.impl == "reference":
        if rank != 0:
            return
        cb, meta, k, el = cpu_reference((a.workload, None), a.steps, a.warmup, budget_s=90.0)
        line = dict(impl="reference", metric=METRIC, value=cb["value"], unit=METRIC, n_gpus=a.gpus, steps=k,
                    warmup=a.warmup, ms_per_step=1e3 * el / k, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype=meta["dtype"], data="synthetic",
                    config=dict(workload=meta["workload"], grid=meta["grid"], nbatch=meta["nbatch"],
                                note="reference arm = CPU restatement of the reference on all host threads (" + cb["sample"] +
                                     "); the Julia/FFTW reference cannot run in this image"),
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
    if rank == 0:
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
    chained_ms = allmax(res["chained_ms"])
    total_pts = allsum(float(pts))
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
    roofline = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                    traffic=ncu_traffic("row" if dom == 0 else "str", a.workload), peak_source=peak_note,
                    kernel=names[dom], alg_bytes_per_launch=alg_bytes, avg_launch_ms=avg_ms,
                    l2="cold: 252 MiB overwritten after every kernel",
                    warm_l2=dict(avg_launch_ms=warm_avg, achieved=alg_bytes / (warm_avg * 1e-3) / 1e9),
                    per_kernel_ms=dict(row=fl["ms"][0] / max(1, fl["n"][0]), str=fl["ms"][1] / max(1, fl["n"][1])),
                    step_contract=dict(b_alg_bytes_per_point=meta["b_alg_contract"],
                                       frac_cold=meta["b_alg_contract"] * pts / ((res.get("cold_bracket_ms", ker_ms_flush)) / a.steps * 1e-3) / 1e9 / peak,
                                       frac_chained=meta["b_alg_contract"] * pts / (chained_ms / a.steps * 1e-3) / 1e9 / peak))
    cold_ms = allmax(res["cold_bracket_ms"]) if "cold_bracket_ms" in res else ker_ms_flush
    value_cold = total_pts * a.steps / (cold_ms * 1e-3)
    value_chained = total_pts * a.steps / (chained_ms * 1e-3)
    line = dict(metric=METRIC, value=value_cold, unit=METRIC, n_gpus=world, steps=a.steps, warmup=a.warmup,
                ms_per_step=cold_ms / a.steps, higher_is_better=True,
                scaling="strong" if a.workload in ("c4", "c5") else "weak", vs_baseline=None, dtype=meta["dtype"],
                data="synthetic",
                config=dict(workload=meta["workload"], grid=meta["grid"], ncomp=meta["ncomp"], nbatch=meta["nbatch"],
                            parallelism=("single GPU" if world == 1 else
                                         (f"{world} GPUs, trajectories sharded {4096 // world}/GPU" if a.workload == "c4"
                                          else f"{world} independent replicas (C2 does not shard)")),
                            l2="L2 flushed (252 MiB overwritten) after EVERY kernel of the K timed steps; `value` = "
                               "points*K / (one CUDA-event bracket around the K flushed steps minus the same number of "
                               "flushes timed alone); `cold_kernel_sum` = the same from per-kernel event pairs; `chained` = "
                               "the K steps back to back without flush (state + table stay L2-resident, as in a production run)"),
                cold_kernel_sum=dict(value=total_pts * a.steps / (ker_ms_flush * 1e-3), ms_per_step=ker_ms_flush / a.steps),
                cold_bracket=res.get("cold_bracket_detail"),
                chained=dict(value=value_chained, ms_per_step=chained_ms / a.steps),
                warm_kernel_sum=dict(value=total_pts * a.steps / (ker_ms_warm * 1e-3), ms_per_step=ker_ms_warm / a.steps),
                roofline=roofline,
                e2e=dict(value=total_pts * a.steps / e2e_s, unit=METRIC,
                         h2d_bytes_per_step=res["e2e"]["h2d_bytes_per_step"],
                         d2h_bytes_per_step=res["e2e"]["d2h_bytes_per_step"],
                         what="pinned host u0 -> ggp_set_state, ggp_step in save intervals of <=1000 steps, "
                              "ggp_save_async of the state to pinned host after each + final ggp_save_wait (solve! of the "
                              "host interface); plan creation excluded"),
                gpu_launches=res["launches"])
    windows = list(res["windows"])
    res["iter"].close()

    # ---- extras: the sharded ensemble (C4) and C3 ------------------------------------------------
    extra = {}
    if not a.no_extra and a.workload == "c2":
        try:
            nb = 4096 // world
            ksteps = 20
            r4 = measure(G, "c4", ksteps, 3, local, nbatch=nb, batch_offset=rank * nb, do_e2e=False, do_flush=False,
                         comm=(lambda it: attach_comm(G, it, world, rank)) if world > 1 else None)
            it4 = r4["iter"]
            barrier()
            t0 = time.perf_counter()
            nk = it4.observe(G.lib.OBS_MOMENTUM)      # all-reduced over ranks when world > 1
            obs_s = allmax(time.perf_counter() - t0)
            ms4 = allmax(r4["chained_ms"])
            pts4 = allsum(float(r4["meta"]["points"]))
            k4 = int(np.argmax(r4["prof"]["ms"]))
            extra["c4_ensemble"] = dict(
                workload=r4["meta"]["workload"].replace(f"x {nb} ", "x 4096 ") if world > 1 else r4["meta"]["workload"],
                scaling="strong", ntraj_total=nb * world, ntraj_per_gpu=nb, value=pts4 * ksteps / (ms4 * 1e-3),
                unit=METRIC, ms_per_step=ms4 / ksteps, steps=ksteps,
                frac_of_hbm_roofline_contract=96 * pts4 / world / (ms4 / ksteps * 1e-3) / 1e9 / peak,
                observable_allreduce_s=obs_s, nk_sum=float(nk.sum()),
                per_kernel_ms=dict(row=r4["prof"]["ms"][0] / max(1, r4["prof"]["n"][0]),
                                   str=r4["prof"]["ms"][1] / max(1, r4["prof"]["n"][1])), dominant=["row", "str"][k4 if k4 < 2 else 0])
            windows += r4["windows"]
            it4.close()
        except Exception as e:  # extras must never take the headline down
            extra["c4_ensemble"] = dict(error=repr(e))
        if True:    # N = 1: the unsharded 512^3 grid, the reference point of the slab scaling
            try:
                k5 = 20
                r5 = measure(G, "c5", k5, 3, local, n=512, slab=(rank, world), do_e2e=False, do_flush=False,
                             comm=(lambda it: attach_comm(G, it, world, rank)) if world > 1 else None)
                ms5 = allmax(r5["chained_ms"])
                pts5 = allsum(float(r5["meta"]["points"]))
                extra["c5_slab"] = dict(workload=r5["meta"]["workload"], scaling="strong", value=pts5 * k5 / (ms5 * 1e-3),
                                        unit=METRIC, ms_per_step=ms5 / k5, steps=k5,
                                        frac_of_hbm_roofline_contract=88 * pts5 / world / (ms5 / k5 * 1e-3) / 1e9 / peak,
                                        per_kernel_ms=dict(row=r5["prof"]["ms"][0] / max(1, r5["prof"]["n"][0]),
                                                           str_d=r5["prof"]["ms"][1] / max(1, r5["prof"]["n"][1]),
                                                           str_fi=r5["prof"]["ms"][2] / max(1, r5["prof"]["n"][2])))
                windows += r5["windows"]
                r5["iter"].close()
            except Exception as e:
                extra["c5_slab"] = dict(error=repr(e))
        if world == 1:
            try:
                r3 = measure(G, "c3", 200, 5, local, do_e2e=False, do_flush=False)
                ms3 = r3["chained_ms"]
                extra["c3"] = dict(workload=r3["meta"]["workload"], value=r3["meta"]["points"] * 200 / (ms3 * 1e-3),
                                   unit=METRIC, ms_per_step=ms3 / 200,
                                   frac_of_hbm_roofline_contract=288 * r3["meta"]["points"] / (ms3 / 200 * 1e-3) / 1e9 / peak)
                r3["iter"].close()
            except Exception as e:
                extra["c3"] = dict(error=repr(e))
            try:
                r3 = measure(G, "c3_c64", 200, 5, local, do_e2e=False, do_flush=False)
                ms3 = r3["chained_ms"]
                extra["c3_c64"] = dict(workload=r3["meta"]["workload"], value=r3["meta"]["points"] * 200 / (ms3 * 1e-3),
                                       unit=METRIC, ms_per_step=ms3 / 200,
                                       frac_of_hbm_roofline_contract=144 * r3["meta"]["points"] / (ms3 / 200 * 1e-3) / 1e9 / peak)
                r3["iter"].close()
            except Exception as e:
                extra["c3_c64"] = dict(error=repr(e))
    if extra:
        line["extra"] = extra

    if rank == 0:
        line["clocks"] = sampler.stop(windows)
        if not a.no_cpu and world >= 1:
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

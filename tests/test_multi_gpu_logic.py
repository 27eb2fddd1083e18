"""world_size-2 `gloo` tests (CPU) of the N > 1 path: trajectory sharding + the observable
all-reduce.  The per-rank compute is stood in for by the CPU oracle -- the point here is the host
logic: shards tile the ensemble exactly once, a sharded ensemble reproduces the unsharded one
(trajectories are independent units: no data-path collective), and the summed observable is the
full-ensemble observable."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ggp_b200 as G
    import ggp_oracle as O
    import problems as P
    ntraj = 10
    pb = P.truncated_wigner(O, ntraj=ntraj, N=16, ndim=2, tspan=(0, 0.5))
    lo, hi = G.parallel.shard_range(ntraj, world, rank)
    # identical global noise stream on every rank, sliced by trajectory (== counter-based stream indexed globally)
    rng = np.random.default_rng(99)

    def noise_full(shape, dtype):
        full = (ntraj,) + tuple(shape[1:])
        z = (rng.standard_normal(full) + 1j * rng.standard_normal(full)) / np.sqrt(2)
        return z[lo:hi].astype(dtype)

    u0 = (pb["u0"][0][lo:hi],)
    kw = dict(pb["kwargs"])
    kw["noise_prototype"] = tuple(np.empty_like(x) for x in u0)
    prob = O.GrossPitaevskiiProblem(u0, pb["lengths"], **kw)
    _, sol = O.solve(prob, O.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1, save_start=False,
                     noise_source=noise_full)
    local = sol[0][0]
    nk_local = (np.abs(np.fft.fft2(local)) ** 2).sum(0) / local[0].size ** 2
    nk = G.parallel.allreduce_observable(nk_local, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, local))
    if rank == 0:
        q.put((nk, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_ensemble_equals_full_ensemble_gloo():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ggp_b200 as G
    import ggp_oracle as O
    import problems as P
    world, ntraj = 2, 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    nk, gathered = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # shards tile [0, ntraj) exactly once
    spans = sorted((lo, hi) for lo, hi, _ in gathered)
    assert spans[0][0] == 0 and spans[-1][1] == ntraj and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    # unsharded reference run with the same global noise stream
    pb = P.truncated_wigner(O, ntraj=ntraj, N=16, ndim=2, tspan=(0, 0.5))
    rng = np.random.default_rng(99)
    prob = O.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    _, sol = O.solve(prob, O.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1, save_start=False,
                     noise_source=lambda s, d: ((rng.standard_normal(s) + 1j * rng.standard_normal(s)) / np.sqrt(2)).astype(d))
    full = sol[0][0]
    cat = np.concatenate([x for _, _, x in sorted(gathered, key=lambda g: g[0])], axis=0)
    assert np.array_equal(cat, full)
    ref = (np.abs(np.fft.fft2(full)) ** 2).sum(0) / full[0].size ** 2
    assert np.allclose(nk, ref, rtol=1e-12)


def test_shard_range_properties():
    import ggp_b200 as G
    for n in (1, 7, 4096, 10 ** 4):
        for w in (1, 2, 3, 4, 8):
            spans = [G.parallel.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1

"""Multi-GPU plumbing for the one shape of the path that shards with no data-path collective:
stochastic trajectory ensembles.  One process per GPU; trajectories (the trailing
batch dims of u0, src/problem.jl:112) are split into contiguous blocks; tables are replicated; the
Philox counter uses the GLOBAL trajectory index (ggp_desc.batch_offset), so results do not depend on
the number of ranks.  The only collective is the sum of the saved observables.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_range(ntraj: int, world: int, rank: int):
    """Contiguous block [lo, hi) of trajectories owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(ntraj, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_problem(G, prob, world: int, rank: int):
    """The rank-local GrossPitaevskiiProblem: a slice of the leading (slowest) NumPy axis of every
    u0 / noise_prototype array == the last Julia batch dim.  Returns (local_prob, batch_offset) where
    batch_offset counts trajectories (all batch dims flattened) before this rank's block."""
    nb = prob.u0[0].shape[0]
    lo, hi = shard_range(nb, world, rank)
    per = int(np.prod(prob.u0[0].shape[1:len(prob.u0[0].shape) - prob.ndim], dtype=np.int64)) if prob.u0[0].ndim - prob.ndim > 1 else 1
    u0 = tuple(x[lo:hi] for x in prob.u0)
    proto = prob.noise_prototype
    if isinstance(proto, tuple):
        proto = tuple(np.empty_like(x) for x in u0) if proto[0].shape == prob.u0[0].shape else proto
    local = G.GrossPitaevskiiProblem(u0, prob.lengths, dispersion=prob.dispersion, potential=prob.potential,
                                     nonlinearity=prob.nonlinearity, pump=prob.pump,
                                     position_noise_func=prob.position_noise_func,
                                     momentum_noise_func=prob.momentum_noise_func, noise_prototype=proto,
                                     param=prob.param)
    return local, lo * per


def attach_nccl(G, it, dist, device="cuda"):
    """Give the plan an NCCL communicator (for ggp_observe's all-reduce).  The 128-byte unique id is
    created by rank 0 inside libggp and broadcast with torch.distributed -- plumbing only."""
    import torch
    lib = G.lib.load()
    uid = (C.c_char * 128)()
    if dist.get_rank() == 0:
        G.lib.check(lib.ggp_comm_unique_id(uid))
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=device)
    dist.broadcast(t, 0)
    buf = C.create_string_buffer(bytes(t.cpu().tolist()), 128)
    G.lib.check(lib.ggp_comm_init(it.handle, dist.get_world_size(), dist.get_rank(), buf))


IPC_BLOB_BYTES = 512


def attach_p2p(G, it, dist):
    """Slab-decomposed plans: map every rank's slabs into every other rank (CUDA IPC handles gathered with
    torch.distributed -- plumbing only) so that the strided FFT kernels write across NVLink directly
    (ggp_slab_ipc_export / ggp_slab_ipc_attach, include/ggp.h)."""
    lib = G.lib.load()
    blob = C.create_string_buffer(IPC_BLOB_BYTES)
    G.lib.check(lib.ggp_slab_ipc_export(it.handle, blob))
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, bytes(blob.raw))
    allb = C.create_string_buffer(b"".join(gathered), IPC_BLOB_BYTES * dist.get_world_size())
    G.lib.check(lib.ggp_slab_ipc_attach(it.handle, allb))
    dist.barrier()


def allreduce_observable(local: np.ndarray, dist) -> np.ndarray:
    """Host-side sum over ranks of an observable that was computed per rank (used when the plan has
    no NCCL communicator attached, e.g. the gloo tests)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()

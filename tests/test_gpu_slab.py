"""3-D slab decomposition (BASELINE config C5's shape): z-slabs on 2, 4 and 8 GPUs, compared with the unsharded
CPU oracle -- with the transposes fused into the FFT kernels as peer stores over NVLink (CUDA IPC), with the scatter
passes writing a local staging buffer that the copy engines push to the peers (GGP_SLAB_CE=1), and with the
ncclSend/ncclRecv transposes.  Every world size exercises its own scatter index math
(`dst_shift`, `dst_base` of StrParams).  Needs >= `world` CUDA devices (gpurun --gpus N); smaller boxes skip."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _cases(ns):
    import problems as P
    out = []
    for dtype in (np.complex128, np.complex64):
        pb = P.kerr3d(ns, N=32, dtype=dtype, nsteps=6)
        out.append(("kerr3d_%s" % np.dtype(dtype).name, pb))
    # rectangular grid + scalar potential (z-slab table slicing) + non-separable dispersion (full table path)
    rng = np.random.default_rng(3)
    shape = (64, 32, 16)   # numpy (n3, n2, n1)
    u0 = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex128)

    def disp(ks, p):
        return (ks[0] ** 2 + ks[1] ** 2 + ks[2] ** 2) / 2 + 0.05 * ks[0] * ks[2]

    def pot(r, p):
        return 0.3 * (r[2] - 2.0) ** 2 + 0.1 * r[0] - 0.02j

    def nl(u, p):
        return 0.5 * ns.abs2(u[0])
    out.append(("rect_potential", dict(u0=(u0,), lengths=(5.0, 7.0, 9.0),
                                       kwargs=dict(dispersion=disp, potential=pot, nonlinearity=nl),
                                       tspan=(0.0, 0.05), dt=0.01, nsaves=1)))
    return out


def _worker(rank, world, port, q, mode):
    p2p = mode != "nccl"
    if mode == "copy_engines":          # scatter passes write a local staging buffer, the copy engines push it
        os.environ["GGP_SLAB_CE"] = "1"
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import ggp_b200 as G
    results = {}
    for name, pb in _cases(G):
        prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
        it = G.init(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"], device=rank,
                    slab=(rank, world))
        G.parallel.attach_nccl(G, it, dist)
        if p2p:
            G.parallel.attach_p2p(G, it, dist)
        ts, sol = G.solve_(it)
        it.close()
        gathered = [None] * world
        dist.all_gather_object(gathered, sol[0][-1])
        results[name] = np.concatenate(gathered, axis=0)      # z-slabs stack along numpy axis 0
    # position noise on a slab plan: the Philox counters are global element indices, so the decomposed run must
    # reproduce the single-GPU run of the same seed (rank 0 runs that one as well)
    import problems as P
    pb = P.kerr3d(G, N=32, dtype=np.complex128, nsteps=4)
    kw = dict(pb["kwargs"], noise_prototype=(np.empty_like(pb["u0"][0]),), position_noise_func=lambda u, r, p: 0.3)
    prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **kw)
    it = G.init(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1, device=rank, slab=(rank, world), rng=77)
    G.parallel.attach_nccl(G, it, dist)
    if p2p:
        G.parallel.attach_p2p(G, it, dist)
    ts, sol = G.solve_(it)
    it.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, sol[0][-1])
    if rank == 0:
        whole = np.concatenate(gathered, axis=0)
        _, ref = G.solve(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=1, device=0, rng=77, show_progress=False)
        noiseless = G.solve(G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"]), G.StrangSplitting(),
                            pb["tspan"], dt=pb["dt"], nsaves=1, device=0, show_progress=False)[1]
        results["noise_philox_err"] = float(np.linalg.norm((whole - ref[0][-1]).ravel()) / np.linalg.norm(ref[0][-1].ravel()))
        results["noise_effect"] = float(np.linalg.norm((ref[0][-1] - noiseless[0][-1]).ravel()) / np.linalg.norm(ref[0][-1].ravel()))
    if rank == 0:
        q.put(results)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("mode", ["peer_stores", "nccl", "copy_engines"])
def test_slab_decomposition_matches_oracle(mode, world):
    p2p = mode != "nccl"
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ggp_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300) + {"peer_stores": 301, "nccl": 0, "copy_engines": 602}[mode] + 7 * world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    results = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for name, pb in _cases(O):
        prob = O.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
        _, sol = O.solve(prob, O.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"])
        ref = sol[0][-1].astype(np.complex128)
        got = results[name].astype(np.complex128)
        err = np.linalg.norm((got - ref).ravel()) / np.linalg.norm(ref.ravel())
        tol = 1e-4 if sol[0].dtype == np.complex64 else 1e-10
        assert err <= tol, (name, err)
    # slab plan with position noise == single-GPU plan with the same seed (and the noise did something)
    assert results["noise_effect"] > 1e-3 and results["noise_philox_err"] <= 1e-10, (results["noise_philox_err"], results["noise_effect"])

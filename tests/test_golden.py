"""Committed golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py from
the oracle): the oracle must keep reproducing them (CPU), and the CUDA path must match them (GPU)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as MG  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))


def _rel(a, b):
    a = np.asarray(a, dtype=np.complex128)
    b = np.asarray(b, dtype=np.complex128)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


@pytest.mark.parametrize("name", sorted(MG.CASES))
def test_oracle_reproduces_golden(name):
    ts, sol = MG.run_oracle(name)
    assert np.array_equal(ts, GOLD[name + "/ts"])
    for c, s in enumerate(sol):
        g = GOLD[f"{name}/u{c}"]
        assert s[-1].dtype == g.dtype and s[-1].shape == g.shape
        assert _rel(s[-1], g) <= (1e-6 if g.dtype == np.complex64 else 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MG.CASES))
def test_cuda_path_matches_golden(name):
    import ggp_b200 as G
    pb, seed = MG.build(G, name)
    noise = None
    if seed is not None:          # the oracle's noise buffer, regenerated from the committed seed
        rec = []
        MG.run_oracle(name, record=rec)
        noise = rec
    prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    ts, sol = G.solve(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"],
                      save_start=pb.get("save_start", True), noise_buffers=noise)
    assert np.array_equal(ts, GOLD[name + "/ts"])
    num = den = 0.0
    for c, s in enumerate(sol):
        g = GOLD[f"{name}/u{c}"].astype(np.complex128)
        num += np.linalg.norm((s[-1].astype(np.complex128) - g).ravel()) ** 2
        den += np.linalg.norm(g.ravel()) ** 2
    tol = 1e-4 if sol[0].dtype == np.complex64 else 1e-10
    assert np.sqrt(num / den) <= tol

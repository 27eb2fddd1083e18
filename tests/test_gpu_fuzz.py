"""Randomised parity: 160 random legal problems (tests/problems.py: fuzz) through the CUDA path and the CPU oracle.
Dimensions, axis lengths (powers of two and others), batch dims, 1-3 components, every table kind, Number / SVector /
SMatrix nonlinearities, every pump route and host-fed noise are drawn independently, so the fused kernels, the
component-parallel kernels and the generic plan all meet combinations no dedicated test names.
Tolerance: BASELINE's (rel. L2 <= 1e-10 ComplexF64, <= 1e-4 ComplexF32)."""
import numpy as np
import pytest

import problems as P
from test_gpu_parity import run_both, rel_l2, TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import ggp_b200
    ggp_b200.load()
    assert ggp_b200.lib.load().ggp_device_count() >= 1
    return ggp_b200


@pytest.mark.parametrize("seed", range(160))
def test_random_problem_matches_oracle(G, seed):
    import ggp_oracle as O
    pb = P.fuzz(O, seed)
    g, o = run_both(G, P.fuzz, noise_seed=(1000 + seed) if pb["noisy"] else None, seed=seed)
    err = rel_l2(g, o)
    print(f"\nseed {seed}: {pb['desc']}  rel L2 {err:.2e}")
    assert np.isfinite(err) and err <= TOL[np.dtype(pb['u0'][0].dtype)], pb["desc"]

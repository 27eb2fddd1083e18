"""BASELINE config C2 over its FULL tspan (north_star: "after the full tspan"): 2048^2 ComplexF32,
dt = 1e-3, 10 000 steps on the GPU against the committed CPU-oracle fixture tests/golden/c2_full_tspan_v1.npz (every
16th point of the field at 20 / 100 / 1000 / 2000 / 5000 / 10 000 steps in ComplexF32 AND ComplexF64, made by
tests/golden/make_c2_full_tspan.py).  The three mutual distances are printed.

Gate.  <= 1e-4 against both oracles up to 1000 steps.  At 10 000 steps the ComplexF32 gate is below the distance
between any two correct fp32 implementations (SURVEY §8d': pocketfft-fp32 vs MKL-fp32 3.1e-4; the fixture's own
fp32-oracle-to-fp64-oracle distance is recorded in it), so the full-tspan assertion is the meaningful one: the GPU run
must be at least as close to the fp64 solution as the fp32 CPU oracle is (x 1.25), and its norm must not drift."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "c2_full_tspan_v1.npz")


@pytest.mark.skipif(not os.path.exists(FIX), reason="fixture not generated yet (tests/golden/make_c2_full_tspan.py)")
def test_c2_full_tspan_10k_steps():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import c2_full_tspan_gpu as T
    fix = np.load(FIX)
    out = T.run(fix)
    for r in out["checkpoints"]:
        print(f"\n{r['steps']:6d} steps: ours-o32 {r['ours_vs_o32']:.3e}  ours-o64 {r['ours_vs_o64']:.3e}  "
              f"o32-o64 {r['o32_vs_o64_subsample']:.3e}  norm ours/o64 {r['norm_ours'] / r['norm_o64'] - 1:+.2e}", end="")
    print()
    for r in out["checkpoints"]:
        if r["steps"] <= 1000:
            assert r["ours_vs_o32"] <= 1e-4 and r["ours_vs_o64"] <= 1e-4, r
        assert r["ours_vs_o64"] <= max(1e-4, 1.25 * r["o32_vs_o64_subsample"]), r
        # every fp32 transform loses norm systematically (the MKL-fp32 oracle: -1.1e-7 per step, this backend -6e-8);
        # the GPU run must not drift more than the fp32 CPU oracle does
        assert abs(r["norm_ours"] / r["norm_o64"] - 1) <= max(1e-4, 1.25 * abs(r["norm_o32"] / r["norm_o64"] - 1)), r

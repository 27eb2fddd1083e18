"""GPU tests of the two rows next to the hot path: N2 streaming saves (ggp_save_async /
ggp_save_wait behind solve!'s `map(copy!, slice, iter.u)`, src/fixed_time_stepping.jl:48) and N3 checkpoint /
resume (absent in the reference).  Both are bit-exactness properties of the CUDA path itself; parity with the
oracle of the same runs is test_gpu_parity.py's job."""
import numpy as np
import pytest

import problems as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import ggp_b200
    ggp_b200.load()
    assert ggp_b200.lib.load().ggp_device_count() >= 1
    return ggp_b200


def _iterator(G, pb, **kw):
    prob = G.GrossPitaevskiiProblem(pb["u0"], pb["lengths"], **pb["kwargs"])
    return G.init(prob, G.StrangSplitting(), pb["tspan"], dt=pb["dt"], nsaves=pb["nsaves"],
                  save_start=pb.get("save_start", True), show_progress=False, **kw)


@pytest.mark.parametrize("factory,kw", [
    (P.kerr2d, dict(N=256, dtype=np.complex64, nsteps=64)),
    (P.exciton_polariton, dict(N=64, nsaves=8, tspan=(0, 20), time_pump=True)),
])
def test_streaming_saves_equal_blocking_saves(G, factory, kw):
    """solve! with overlapped saves == stepping with a blocking ggp_get_state after every interval, bit for bit,
    for every saved slice (the snapshot is taken in stream order, later steps must not leak into it)."""
    pb = factory(G, **kw)
    if factory is P.kerr2d:
        pb["nsaves"] = 8
    it = _iterator(G, pb)
    ts, sol = G.solve_(it)
    sol = [np.array(x) for x in sol]
    it.close()
    it2 = _iterator(G, pb)
    off = 1 if it2.save_start else 0
    for n in range(it2.nsaves):
        it2.advance(it2.steps_per_save)
        u = it2.fetch()
        for c in range(it2.M):
            assert np.array_equal(sol[c][n + off], u[c]), (n, c)
    for c in range(it2.M):
        assert np.array_equal(sol[c][0], pb["u0"][c].astype(sol[c].dtype))
    it2.close()


def test_save_async_back_to_back_and_wait_idempotent(G):
    """Two saves without stepping in between land the same data; ggp_save_wait may be called repeatedly and
    before any save."""
    pb = P.kerr2d(G, N=128, dtype=np.complex128, nsteps=16)
    pb["nsaves"] = 4
    it = _iterator(G, pb)
    it.save_wait()
    it.advance(3)
    it.save_async(1)
    it.save_async(2)
    it.advance(2)          # must not disturb the snapshots in flight
    it.save_wait()
    it.save_wait()
    assert np.array_equal(it.result[0][1], it.result[0][2])
    assert not np.array_equal(it.result[0][1], it.result[0][0])
    ref = _iterator(G, pb)
    ref.advance(3)
    assert np.array_equal(ref.fetch()[0], it.result[0][1])
    ref.close()
    it.close()


@pytest.mark.parametrize("case", ["kerr_c64", "polariton_pump", "wigner_philox_2d", "wigner_philox_1d"])
def test_checkpoint_resume_is_bit_identical(G, case):
    """n1 steps, checkpoint, NEW plan, restore, n2 steps == n1 then n2 steps on one plan -- including the Philox
    stream (counter word in the blob) and the one-dt-late pump amplitude F_now (quirk Q1)."""
    if case == "kerr_c64":
        pb, kw = P.kerr2d(G, N=128, dtype=np.complex64, nsteps=40), {}
    elif case == "polariton_pump":
        pb, kw = P.exciton_polariton(G, N=64, nsaves=4, tspan=(0, 20), time_pump=True), {}
    elif case == "wigner_philox_2d":
        pb, kw = P.truncated_wigner(G, ntraj=8, N=64, ndim=2, tspan=(0, 2), dt=0.05), dict(rng=99)
    else:
        pb, kw = P.windowed_ft(G, ntraj=32), dict(rng=7)
    a = _iterator(G, pb, **kw)
    total = a.nsaves * a.steps_per_save
    n1 = max(1, total // 3)
    n2 = min(total - n1, 2 * n1 + 1)
    a.advance(n1)
    blob = a.checkpoint()
    a.advance(n2)
    want = [np.array(x) for x in a.fetch()]
    a.close()
    b = _iterator(G, pb, **kw)
    b.restore(blob)
    assert b._step_index == n1
    b.advance(n2)
    got = b.fetch()
    for w, g in zip(want, got):
        assert np.array_equal(w, g)
    # and the restored run is not trivially equal to a fresh run of n2 steps
    c = _iterator(G, pb, **kw)
    c.advance(n2)
    assert not np.array_equal(c.fetch()[0], want[0])
    c.close()
    b.close()


def test_checkpoint_rejects_foreign_blob(G):
    pb = P.kerr2d(G, N=128, dtype=np.complex64, nsteps=8)
    a = _iterator(G, pb)
    blob = a.checkpoint()
    other = _iterator(G, P.kerr2d(G, N=64, dtype=np.complex64, nsteps=8))
    with pytest.raises(ValueError):
        other.restore(blob)
    same_size = _iterator(G, P.kerr2d(G, N=128, dtype=np.complex64, nsteps=8))
    bad = bytearray(blob)
    bad[0] ^= 0xFF
    with pytest.raises(G.lib.GgpError):
        same_size.restore(bytes(bad))
    # a shape mismatch at equal byte count (256x64 vs 128x128) is caught by the header, not by the size
    lib = G.lib.load()
    n = int(lib.ggp_checkpoint_bytes(a.handle))
    buf = np.frombuffer(blob, dtype=np.uint8)[:n].copy()
    rect = _iterator(G, _rect_problem(G))
    assert int(lib.ggp_checkpoint_bytes(rect.handle)) == n
    assert lib.ggp_checkpoint_load(rect.handle, buf.ctypes.data, n) == -1
    assert b"different shape" in lib.ggp_last_error()
    assert lib.ggp_checkpoint_save(a.handle, buf.ctypes.data, n - 1) == -1
    for it in (a, other, same_size, rect):
        it.close()


def _rect_problem(G):
    pb = P.kerr2d(G, N=128, dtype=np.complex64, nsteps=8)
    u0 = np.ascontiguousarray(np.resize(pb["u0"][0], (64, 256)))      # Julia (256, 64)
    pb["u0"] = (u0,)
    pb["lengths"] = (pb["lengths"][0] * 2, pb["lengths"][1] / 2)
    return pb

"""GPU parity of BorisPush.push against the oracle and the reference-generated
golden vectors: bit-for-bit (fp64, no tolerance needed; the north-star bound is
1e-12 per step / 1e-9 after 1000 steps)."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import (M_E, M_P, Q_E, Q_P, assert_bits_equal, beam_ensemble, make_owner, thermal_ensemble)
from oracle import boris as oboris

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tp():
    import turbopy_b200
    return turbopy_b200


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(f"{golden_dir}/boris.npz")


def pusher(tp, dt):
    return tp.BorisPush(make_owner(dt=dt), {"type": "BorisPush"})


def to_layout(tp, arr, layout):
    if layout == "soa":
        return tp.to_soa(arr)
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def test_division_selftest(tp):
    from turbopy_b200 import _lib
    bad, taken = C.c_int64(-1), C.c_int64(-1)
    n = 200_000_000
    _lib.check(_lib.lib().tp_selftest_division(n, 12345, C.byref(bad), C.byref(taken), None))
    assert bad.value == 0
    assert taken.value > 0.45 * n        # the fast path is the common case, not a dead branch


def test_roots_selftest(tp):
    """The branch-free in-range reciprocal / sqrt of the fast path equal IEEE
    1.0/b and sqrt(a) wherever the guard accepts the operand."""
    from turbopy_b200 import _lib
    bad, taken = C.c_int64(-1), C.c_int64(-1)
    n = 200_000_000
    _lib.check(_lib.lib().tp_selftest_roots(n, 777, C.byref(bad), C.byref(taken), None))
    assert bad.value == 0
    assert taken.value > n


def test_known_answer_crossed_fields(tp, golden):
    """The reference's own known-answer test, tests/test_computetools.py:40-62."""
    tool = pusher(tp, 1e-9)
    x = torch.zeros((1, 3), dtype=torch.float64, device="cuda")
    p = torch.zeros((1, 3), dtype=torch.float64, device="cuda")
    E = np.array([[10.0, 0, 0]]); B = np.array([[0, 0, 10.0]])
    for _ in range(10):
        assert tool.push(x, p, Q_P, M_P, E, B) is None
    assert np.allclose(x.cpu().numpy(), [[2.20028479e-09, -1.04482737e-08, 0]])
    assert np.allclose(p.cpu().numpy() / M_P, [[0.47183691, -1.88168585, 0]])
    assert_bits_equal(x.cpu().numpy(), golden["crossed_x"], "x vs reference")
    assert_bits_equal(p.cpu().numpy(), golden["crossed_p"], "p vs reference")


def test_free_streaming(tp):
    """tests/test_computetools.py:15-37: E=B=0, x = v*N*dt."""
    tool = pusher(tp, 1e-9)
    x = torch.zeros((1, 3), dtype=torch.float64, device="cuda")
    v = np.array([[0, 0, 3.0e2]])
    p = torch.from_numpy(M_P * v).cuda()
    for _ in range(10):
        tool.push(x, p, Q_P, M_P, np.zeros(3), np.zeros(3))
    assert np.allclose(x.cpu().numpy(), v * 10 * 1e-9)


@pytest.mark.parametrize("kind", ["thermal", "g10", "beam"])
@pytest.mark.parametrize("layout", ["soa", "aos"])
@pytest.mark.parametrize("fields", ["uni", "pp"])
def test_one_step_golden(tp, golden, kind, layout, fields):
    tool = pusher(tp, float(golden["dt"]))
    pos = to_layout(tp, golden[f"{kind}_pos0"], layout)
    mom = to_layout(tp, golden[f"{kind}_mom0"], layout)
    if fields == "uni":
        E, B = golden["Eu"], golden["Bu"]
    else:
        E = torch.from_numpy(golden[f"{kind}_Ep"]).cuda()
        B = to_layout(tp, golden[f"{kind}_Bp"], layout)
    ptr = (pos.data_ptr(), mom.data_ptr())
    tool.push(pos, mom, Q_E, M_E, E, B)
    assert (pos.data_ptr(), mom.data_ptr()) == ptr            # in place, same storage
    assert_bits_equal(pos.cpu().numpy(), golden[f"{kind}_{fields}_pos1"], "position")
    assert_bits_equal(mom.cpu().numpy(), golden[f"{kind}_{fields}_mom1"], "momentum")


def test_uniform_field_as_device_tensor(tp, golden):
    tool = pusher(tp, float(golden["dt"]))
    pos = tp.to_soa(golden["thermal_pos0"]); mom = tp.to_soa(golden["thermal_mom0"])
    E = torch.from_numpy(golden["Eu"]).cuda(); B = torch.from_numpy(golden["Bu"]).cuda().reshape(1, 3)
    tool.push(pos, mom, Q_E, M_E, E, B)
    assert_bits_equal(pos.cpu().numpy(), golden["thermal_uni_pos1"])
    assert_bits_equal(mom.cpu().numpy(), golden["thermal_uni_mom1"])


def test_trajectory_1000_steps(tp, golden):
    """Config 1 subset: 1000 steps in uniform E x B, bit-identical to the reference
    (north-star bound: 1e-9 relative after 1000 steps)."""
    tool = pusher(tp, 1e-12)
    pos = tp.to_soa(golden["traj_pos0"]); mom = tp.to_soa(golden["traj_mom0"])
    for _ in range(1000):
        tool.push(pos, mom, Q_E, M_E, golden["Eu"], golden["Bu"])
    assert_bits_equal(pos.cpu().numpy(), golden["traj_pos1000"])
    assert_bits_equal(mom.cpu().numpy(), golden["traj_mom1000"])


def test_beam_trajectory_2000_steps(tp, golden):
    """Config 4 subset: gamma up to 1e3 in B = 10 T."""
    tool = pusher(tp, float(golden["beam_dt"]))
    pos = tp.to_soa(golden["beam_traj_pos0"]); mom = tp.to_soa(golden["beam_traj_mom0"])
    for _ in range(2000):
        tool.push(pos, mom, Q_E, M_E, np.zeros(3), np.array([0, 0, 10.0]))
    assert_bits_equal(pos.cpu().numpy(), golden["beam_traj_pos2000"])
    assert_bits_equal(mom.cpu().numpy(), golden["beam_traj_mom2000"])


def _sha(a):
    import hashlib
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


@pytest.mark.parametrize("how", ["push", "push_steps_100"])
def test_beam_10k_steps_config4(tp, golden_dir, how):
    """BASELINE configs[3] at its stated length: 65 536 beam particles (gamma up to 1e3, B = 10 T) over all 10 000
    steps, bit-for-bit against the UNMODIFIED reference (tests/golden/beam_10k.npz, oracle/make_golden_10k.py;
    north-star bound: 1e-9 relative after 1000 steps -- this is exact after 10 000).  Through ``push`` step by step
    and through ``push_steps(K=100)`` x 100; checkpoints at steps 100 / 1000 / 5000 / 10 000."""
    g = np.load(f"{golden_dir}/beam_10k.npz")
    n, steps = int(g["n"]), int(g["steps"])
    pos0, mom0 = beam_ensemble(n, seed=4567)
    tool = pusher(tp, float(g["dt"]))
    pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
    E, B = np.zeros(3), np.array([0.0, 0.0, float(g["B0"])])
    done = 0
    for cp in (int(c) for c in g["checkpoints"]):
        if how == "push":
            for _ in range(cp - done):
                tool.push(pos, mom, Q_E, M_E, E, B)
        else:
            for _ in range((cp - done) // 100):
                tool.push_steps(pos, mom, Q_E, M_E, E, B, 100)
        done = cp
        hp, hm = pos.cpu().numpy(), mom.cpu().numpy()
        assert np.array_equal(_sha(hp), g[f"sha_pos_{cp}"]), f"position differs from the reference at step {cp}"
        assert np.array_equal(_sha(hm), g[f"sha_mom_{cp}"]), f"momentum differs from the reference at step {cp}"
    assert done == steps
    assert_bits_equal(hp[::16], g["pos_final_16"])
    assert_bits_equal(hm[::16], g["mom_final_16"])


@pytest.mark.parametrize("n", [1, 2, 3, 255, 256, 257, 100_001, 1 << 20])
@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_ragged_sizes_vs_oracle(tp, n, layout):
    dt = 1e-12
    tool = pusher(tp, dt)
    pos0, mom0 = thermal_ensemble(n, seed=n)
    rng = np.random.default_rng(n + 1)
    E = rng.normal(0, 1e6, (n, 3)); B = rng.normal(0, 5.0, (n, 3))
    pos = to_layout(tp, pos0, layout); mom = to_layout(tp, mom0, layout)
    Ed = torch.from_numpy(E).cuda(); Bd = torch.from_numpy(B).cuda()
    steps = 3
    for _ in range(steps):
        tool.push(pos, mom, Q_E, M_E, Ed, Bd)
    o = [np.ascontiguousarray(pos0[:, k]) for k in range(3)] + [np.ascontiguousarray(mom0[:, k]) for k in range(3)]
    e = [np.ascontiguousarray(E[:, k]) for k in range(3)]; b = [np.ascontiguousarray(B[:, k]) for k in range(3)]
    for _ in range(steps):
        oboris.push_soa(*o, Q_E, M_E, *e, *b, dt)
    assert_bits_equal(pos.cpu().numpy(), np.stack(o[:3], axis=1), "position")
    assert_bits_equal(mom.cpu().numpy(), np.stack(o[3:], axis=1), "momentum")


def test_empty_ensemble(tp):
    tool = pusher(tp, 1e-12)
    pos = torch.zeros((0, 3), dtype=torch.float64, device="cuda")
    mom = torch.zeros((0, 3), dtype=torch.float64, device="cuda")
    tool.push(pos, mom, Q_E, M_E, np.zeros(3), np.zeros(3))


def test_zero_fields_signed_zero(tp):
    """E = B = 0 with a negative charge makes -0.0 numerators: the sign of zero
    must survive the shared-reciprocal division like it does in numpy."""
    dt = 1e-12
    tool = pusher(tp, dt)
    pos0 = np.zeros((64, 3)); mom0 = np.zeros((64, 3))
    mom0[::2, 0] = 1e-22
    pos = tp.to_soa(pos0); mom = tp.to_soa(mom0)
    tool.push(pos, mom, Q_E, M_E, np.zeros(3), np.zeros(3))
    rp, rm = pos0.copy(), mom0.copy()
    oboris.push_aos(rp, rm, Q_E, M_E, np.zeros(3), np.zeros(3), dt)
    assert_bits_equal(pos.cpu().numpy(), rp)
    assert_bits_equal(mom.cpu().numpy(), rm)


def test_host_path_numpy_arrays(tp, golden):
    """End-to-end path: numpy (n,3) arrays in, same arrays updated in place."""
    tool = pusher(tp, float(golden["dt"]))
    for fields in ("uni", "pp"):
        pos = golden["g10_pos0"].copy(); mom = golden["g10_mom0"].copy()
        E, B = (golden["Eu"], golden["Bu"]) if fields == "uni" else (golden["g10_Ep"], golden["g10_Bp"])
        tool.push(pos, mom, Q_E, M_E, E, B)
        assert_bits_equal(pos, golden[f"g10_{fields}_pos1"])
        assert_bits_equal(mom, golden[f"g10_{fields}_mom1"])
    # chunked streaming (several chunks, ragged tail) and SoA host arrays (Fortran order)
    n = 600_001
    pos0, mom0 = thermal_ensemble(n, seed=5)
    E, B = np.array([0.0, 1e5, 0.0]), np.array([0.0, 0.0, 1.0])
    rp, rm = pos0.copy(), mom0.copy()
    oboris.push_aos(rp, rm, Q_E, M_E, E, B, float(golden["dt"]))
    pos, mom = pos0.copy(), mom0.copy()
    tool.push(pos, mom, Q_E, M_E, E, B)
    assert_bits_equal(pos, rp); assert_bits_equal(mom, rm)
    posf, momf = np.asfortranarray(pos0), np.asfortranarray(mom0)
    tool.push(posf, momf, Q_E, M_E, E, B)
    assert_bits_equal(posf, rp); assert_bits_equal(momf, rm)


def test_error_behaviour(tp):
    tool = pusher(tp, 1e-12)
    pos = torch.zeros((4, 3), dtype=torch.float64, device="cuda"); mom = torch.zeros_like(pos)
    with pytest.raises(ValueError):            # np.cross with a scalar B (reference: ValueError)
        tool.push(pos, mom, Q_E, M_E, np.zeros(3), 0)
    with pytest.raises(IndexError):            # (3,) particles (reference: IndexError)
        tool.push(pos[0], mom[0], Q_E, M_E, np.zeros(3), np.zeros(3))
    with pytest.raises(TypeError):
        tool.push(pos.float(), mom.float(), Q_E, M_E, np.zeros(3), np.zeros(3))
    with pytest.raises(TypeError):             # no CPU implementation to fall back to
        tool.push(pos.cpu(), mom.cpu(), Q_E, M_E, np.zeros(3), np.zeros(3))


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_push_tma_ring_vs_oracle(tp, layout):
    """Uniform-field push through the TMA ring (>= one 960-particle tile per SM),
    with a ragged tail and particles that must take the recovery path
    (zero momentum is fine; 1e-200 momenta fall below the fast path's domain)."""
    n = 148 * 1024 * 3 + 501
    dt = 1e-12
    tool = pusher(tp, dt)
    pos0, mom0 = thermal_ensemble(n, seed=99)
    mom0[::1000] = 0.0
    mom0[7::5000] = 1e-200
    E, B = np.array([3e4, 1e5, -2e4]), np.array([0.5, -0.25, 1.0])
    pos = to_layout(tp, pos0, layout); mom = to_layout(tp, mom0, layout)
    for _ in range(2):
        tool.push(pos, mom, Q_E, M_E, E, B)
    rp, rm = pos0.copy(), mom0.copy()
    for _ in range(2):
        oboris.push_aos(rp, rm, Q_E, M_E, E, B, dt)
    assert_bits_equal(pos.cpu().numpy(), rp, "position")
    assert_bits_equal(mom.cpu().numpy(), rm, "momentum")


@pytest.mark.parametrize("layout", ["soa", "aos"])
@pytest.mark.parametrize("flayout", ["soa", "aos", "mixed", "strided"])
def test_push_per_particle_fields_tma_ring_vs_oracle(tp, layout, flayout):
    """push() with (n,3) E and B through the field ring (480-particle tiles, >= one per SM), every
    particle / field layout pair, a ragged tail, recovery-path particles; "mixed" and "strided"
    field layouts are not bulk-copyable and must take the register path with the same bits."""
    n = 148 * 480 * 3 + 377
    dt = 1e-12
    tool = pusher(tp, dt)
    pos0, mom0 = thermal_ensemble(n, seed=123)
    mom0[::777] = 0.0
    mom0[5::4000] = 1e-200
    rng = np.random.default_rng(321)
    E = rng.normal(0, 1e6, (n, 3)); B = rng.normal(0, 5.0, (n, 3))
    B[3::2500] = 0.0
    pos = to_layout(tp, pos0, layout); mom = to_layout(tp, mom0, layout)
    if flayout == "strided":
        wide = torch.zeros((n, 4), dtype=torch.float64, device="cuda")
        wide[:, :3] = torch.from_numpy(E).cuda()
        Ed = wide[:, :3]
        Bd = to_layout(tp, B, "aos")
    else:
        Ed = to_layout(tp, E, "aos" if flayout == "aos" else "soa")
        Bd = to_layout(tp, B, "soa" if flayout == "soa" else "aos")
    before = tp._lib.lib().tp_launch_count()
    for _ in range(2):
        tool.push(pos, mom, Q_E, M_E, Ed, Bd)
    assert tp._lib.lib().tp_launch_count() - before == 2
    o = [np.ascontiguousarray(pos0[:, k]) for k in range(3)] + [np.ascontiguousarray(mom0[:, k]) for k in range(3)]
    e = [np.ascontiguousarray(E[:, k]) for k in range(3)]; b = [np.ascontiguousarray(B[:, k]) for k in range(3)]
    for _ in range(2):
        oboris.push_soa(*o, Q_E, M_E, *e, *b, dt)
    assert_bits_equal(pos.cpu().numpy(), np.stack(o[:3], axis=1), "position")
    assert_bits_equal(mom.cpu().numpy(), np.stack(o[3:], axis=1), "momentum")
    assert_bits_equal(Ed.cpu().numpy(), E, "E untouched")
    assert_bits_equal(Bd.cpu().numpy(), B, "B untouched")


def test_invariants_full_size(tp):
    """Config-4-shaped property test at 16 Mi particles: with E = 0 the push is a
    pure rotation, so |p| is conserved to rounding over many steps, and a random
    subset matches the oracle bit for bit (particles are independent)."""
    n = 16 << 20
    g = torch.Generator(device="cuda").manual_seed(7)
    pos = tp.soa_particles(n); mom = tp.soa_particles(n)
    pos.copy_(torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=g))
    mom.copy_(torch.randn((n, 3), dtype=torch.float64, device="cuda", generator=g) * (M_E * 2.9979e8 * 5.0))
    idx = torch.randint(0, n, (4096,), device="cuda", generator=g)
    sub_pos, sub_mom = pos[idx].cpu().numpy(), mom[idx].cpu().numpy()
    p0 = torch.linalg.vector_norm(mom, dim=1)
    dt = 1e-13
    tool = pusher(tp, dt)
    B = np.array([0.3, -0.2, 10.0])
    for _ in range(20):
        tool.push(pos, mom, Q_E, M_E, np.zeros(3), B)
    p1 = torch.linalg.vector_norm(mom, dim=1)
    assert float(((p1 - p0).abs() / p0).max()) < 1e-13
    for _ in range(20):
        oboris.push_aos(sub_pos, sub_mom, Q_E, M_E, np.zeros(3), B, dt)
    assert_bits_equal(pos[idx].cpu().numpy(), sub_pos)
    assert_bits_equal(mom[idx].cpu().numpy(), sub_mom)


@pytest.mark.parametrize("layout", ["soa", "aos"])
@pytest.mark.parametrize("fields", ["uni", "pp"])
@pytest.mark.parametrize("n", [1000, 148 * 480 * 2 + 333])
def test_push_steps_equals_repeated_push(tp, layout, fields, n):
    """push_steps(K) == K x push, bit for bit, on the ring and on the register path, incl. particles that
    take the recovery path (zero and 1e-200 momenta) -- and both equal the oracle's K steps."""
    dt, K = 1e-12, 7
    tool = pusher(tp, dt)
    pos0, mom0 = thermal_ensemble(n, seed=n + 1)
    mom0[::97] = 0.0
    mom0[5::501] = 1e-200
    rng = np.random.default_rng(4)
    if fields == "uni":
        E, B = np.array([3e4, 1e5, -2e4]), np.array([0.5, -0.25, 1.0])
        Ed, Bd = E, B
    else:
        E = rng.normal(0, 1e6, (n, 3)); B = rng.normal(0, 5.0, (n, 3))
        Ed, Bd = to_layout(tp, E, "aos"), to_layout(tp, B, "aos")
    pa, ma = to_layout(tp, pos0, layout), to_layout(tp, mom0, layout)
    pb, mb = to_layout(tp, pos0, layout), to_layout(tp, mom0, layout)
    before = tp._lib.lib().tp_launch_count()
    tool.push_steps(pa, ma, Q_E, M_E, Ed, Bd, K)
    assert tp._lib.lib().tp_launch_count() - before == 1
    for _ in range(K):
        tool.push(pb, mb, Q_E, M_E, Ed, Bd)
    assert_bits_equal(pa.cpu().numpy(), pb.cpu().numpy(), "position")
    assert_bits_equal(ma.cpu().numpy(), mb.cpu().numpy(), "momentum")
    m = min(n, 5000)
    rp, rm = pos0[:m].copy(), mom0[:m].copy()
    for _ in range(K):
        oboris.push_aos(rp, rm, Q_E, M_E, E if fields == "uni" else E[:m], B if fields == "uni" else B[:m], dt)
    assert_bits_equal(pa[:m].cpu().numpy(), rp, "position vs oracle")
    assert_bits_equal(ma[:m].cpu().numpy(), rm, "momentum vs oracle")
    tool.push_steps(pa, ma, Q_E, M_E, Ed, Bd, 0)                 # no-op
    assert_bits_equal(pa.cpu().numpy(), pb.cpu().numpy())
    with pytest.raises(ValueError):
        tool.push_steps(pa, ma, Q_E, M_E, Ed, Bd, -1)


def test_host_path_pins_numpy_arrays_for_their_lifetime(tp):
    """The host path page-locks the ndarray that owns the particle memory the second time it sees it and unpins
    it when that array dies (weakref finalizer), so freshly allocated arrays at recycled addresses are pinned
    afresh and the results stay exact; views share their owner's registration; small and foreign buffers are
    left alone."""
    import gc

    from turbopy_b200 import tools
    dt = 1e-12
    tool = pusher(tp, dt)
    E, B = np.array([3e4, 1e5, -2e4]), np.array([0.5, -0.25, 1.0])
    n = 400_000                                               # 9.6 MB per array: above the pinning threshold
    gc.collect()
    base = len(tools._pinned)
    states = lambda: sorted(str(st) for _, st in tools._pinned.values())   # noqa: E731
    for round_ in range(3):
        pos0, mom0 = thermal_ensemble(n, seed=40 + round_)
        pos, mom = pos0.copy(), mom0.copy()
        tool.push(pos, mom, Q_E, M_E, E, B)
        assert len(tools._pinned) == base + 2 and states().count("None") == 2       # seen once: not pinned yet
        tool.push(pos[:1000], mom[:1000], Q_E, M_E, E, B)     # a view: same owners, second sight -> pinned
        assert len(tools._pinned) == base + 2 and states().count("True") == 2
        tool.push(pos, mom, Q_E, M_E, E, B)
        rp, rm = pos0.copy(), mom0.copy()
        oboris.push_aos(rp, rm, Q_E, M_E, E, B, dt)
        oboris.push_aos(rp[:1000], rm[:1000], Q_E, M_E, E, B, dt)
        oboris.push_aos(rp, rm, Q_E, M_E, E, B, dt)
        assert_bits_equal(pos, rp); assert_bits_equal(mom, rm)
        del pos, mom
        gc.collect()
        assert len(tools._pinned) == base                     # unpinned / forgotten with their owners
    small = np.zeros((10, 3)); tool.push(small, small.copy(), Q_E, M_E, E, B)
    assert len(tools._pinned) == base
    buf = bytearray(8 * 3 * n)
    foreign = np.frombuffer(buf, dtype=np.float64).reshape(n, 3)
    own = np.zeros((n, 3))
    tool.push(foreign, own, Q_E, M_E, E, B)                    # memory owned by a bytearray: never pinned
    tool.push(foreign, own, Q_E, M_E, E, B)
    assert len(tools._pinned) == base + 1 and states().count("True") == 1

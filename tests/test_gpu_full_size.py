"""Parity at BASELINE.json's full sizes through size-independent properties.

Particles are independent given the fields, so a random subset run through the
oracle is an exact (bit-for-bit) check of those particles inside the full-size
run; stencils are local, so sampled windows are exact; sums are checked through
additivity over halves ("checksum of checksums") and against the oracle on the
subset."""
import numpy as np
import pytest
import torch

from helpers import C_LIGHT, M_E, Q_E, assert_bits_equal, make_grid, make_owner
from oracle import boris as oboris, fd as ofd, gather as ogather, moments as omom

pytestmark = pytest.mark.gpu


def subset(gen, n, k=4096):
    return torch.randint(0, n, (k,), device="cuda", generator=gen)


def test_config1_million_particles_1000_steps_graph():
    """configs[0]: 1 M particles in uniform E x B, 1000 steps (here as 10 replays of
    a 100-step CUDA graph); a 4096-particle subset against the oracle, bit for bit
    (north-star bound: 1e-9 relative after 1000 steps)."""
    import turbopy_b200 as tp
    n, dt = 1_000_000, 1e-12
    rng = np.random.default_rng(1234)
    pos0 = rng.uniform(0, 1, (n, 3)); mom0 = M_E * C_LIGHT * rng.normal(0, 0.1, (n, 3))
    E, B = np.array([0.0, 1e5, 0.0]), np.array([0.0, 0.0, 1.0])
    tool = tp.BorisPush(make_owner(dt=dt), {"type": "BorisPush"})
    pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
    tool.push(pos, mom, Q_E, M_E, E, B)                      # step 1 eagerly
    with tp.StepGraph() as g:
        for _ in range(111):
            tool.push(pos, mom, Q_E, M_E, E, B)
    for _ in range(9):
        g.replay()                                           # 1 + 9 * 111 = 1000 steps
    idx = np.random.default_rng(5).integers(0, n, 4096)
    sp, sm = pos0[idx].copy(), mom0[idx].copy()
    for _ in range(1000):
        oboris.push_aos(sp, sm, Q_E, M_E, E, B, dt)
    assert_bits_equal(pos.cpu().numpy()[idx], sp); assert_bits_equal(mom.cpu().numpy()[idx], sm)
    g.close()
    # the same 1000 steps as ONE temporally blocked launch (static fields): every particle identical
    pos2, mom2 = tp.to_soa(pos0), tp.to_soa(mom0)
    tool.push_steps(pos2, mom2, Q_E, M_E, E, B, 1000)
    assert torch.equal(pos2, pos) and torch.equal(mom2, mom)


def test_config2_pif_16mi_particles():
    """configs[1] (the bench workload): 16 Mi electrons, E_y plane wave on a 4097-node
    grid, 20 steps of the example's clock; subset against the oracle."""
    import turbopy_b200 as tp
    n, ng, steps = 16 << 20, 4097, 20
    dt, omega, k = 1e-8 / 20, 2e8, 2e8 / 2.998e8
    grid = make_grid(ng)
    gen = torch.Generator(device="cuda").manual_seed(2345)
    pos, mom = tp.soa_particles(n, fill=0.0), tp.soa_particles(n, fill=0.0)
    pos[:, 0].copy_(0.05 + 0.9 * torch.rand(n, dtype=torch.float64, device="cuda", generator=gen))
    idx = subset(gen, n)
    sp, sm = pos[idx].cpu().numpy(), mom[idx].cpu().numpy()
    tool = tp.BorisPush(make_owner(dt=dt, grid=grid), {"type": "BorisPush"})
    for s in range(steps):
        Ey = np.cos(2 * np.pi * (-omega * (dt * s) + k * (grid.r - 0.5)))
        tool.gather_push(pos, mom, Q_E, M_E, E_grid=(None, torch.from_numpy(Ey).cuda(), None))
        ey, st = ogather.interp_b(sp[:, 0], grid.r, Ey)
        assert (st == 0).all()
        E = np.zeros_like(sp); E[:, 1] = ey
        oboris.push_aos(sp, sm, Q_E, M_E, E, np.zeros(3), dt)
    tool.check_status()
    assert_bits_equal(pos[idx].cpu().numpy(), sp); assert_bits_equal(mom[idx].cpu().numpy(), sm)


def test_config3_shard_125m_particles_1mi_grid():
    """configs[2], one rank's share: 125 M particles, 2^20+1-node replicated grid, six
    field components, kinetic-energy diagnostic; subset vs oracle, and the diagnostic's
    sums are additive over the two halves of the shard (what the all-reduce relies on)."""
    import turbopy_b200 as tp
    n, ng, dt = 125_000_000, (1 << 20) + 1, 1e-12
    grid = make_grid(ng)
    gen = torch.Generator(device="cuda").manual_seed(3456)
    pos, mom = tp.soa_particles(n), tp.soa_particles(n)
    for c in range(3):
        pos[:, c].copy_(0.05 + 0.9 * torch.rand(n, dtype=torch.float64, device="cuda", generator=gen))
        mom[:, c].copy_(torch.randn(n, dtype=torch.float64, device="cuda", generator=gen) * (M_E * C_LIGHT * 0.1))
    comps = [1e5 * np.cos(2 * np.pi * (k + 1) * grid.r) for k in range(3)] + \
            [np.cos(2 * np.pi * (k + 4) * grid.r) for k in range(3)]
    E = torch.from_numpy(np.stack(comps[:3], axis=1)).cuda()
    B = torch.from_numpy(np.stack(comps[3:], axis=1)).cuda()
    idx = subset(gen, n, 16384)
    sp, sm = pos[idx].cpu().numpy(), mom[idx].cpu().numpy()
    tool = tp.BorisPush(make_owner(dt=dt, grid=grid), {"type": "BorisPush"})
    for _ in range(3):
        tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)
        F = [ogather.interp_b(sp[:, 0], grid.r, c)[0] for c in comps]
        oboris.push_aos(sp, sm, Q_E, M_E, np.stack(F[:3], axis=1), np.stack(F[3:], axis=1), dt)
    tool.check_status()
    assert_bits_equal(pos[idx].cpu().numpy(), sp); assert_bits_equal(mom[idx].cpu().numpy(), sm)
    c2 = C_LIGHT ** 2
    whole = tp.reduce_moments(mom, M_E, c2).cpu().numpy()
    h = n // 2
    lo = tp.reduce_moments(mom[:h], M_E, c2).cpu().numpy()
    hi = tp.reduce_moments(mom[h:], M_E, c2).cpu().numpy()
    assert whole[5] == n and lo[5] + hi[5] == n
    assert abs(whole[0] - (lo[0] + hi[0])) <= 1e-12 * whole[0]              # KE additive over shards
    assert abs(whole[4] - (lo[4] + hi[4])) <= 1e-12 * whole[4]
    sub = tp.reduce_moments(tp.to_soa(sm), M_E, c2).cpu().numpy()
    ref = omom.moments(sm[:, 0], sm[:, 1], sm[:, 2], M_E, c2)
    assert abs(sub[0] - ref[0]) <= 1e-12 * ref[0]


def test_config3_sort_by_cell_125m_particles():
    """configs[2] size, the optional reordering: size-independent properties of the result -- keys ascend,
    ties keep their order (stable), it is a permutation (integer checksums of every array are unchanged,
    sum of perm = n(n-1)/2), and the arrays are exactly the gathers of the originals on a sample."""
    import turbopy_b200 as tp
    n, ng = 125_000_000, (1 << 20) + 1
    grid = make_grid(ng)
    gen = torch.Generator(device="cuda").manual_seed(3457)
    pos, mom = tp.soa_particles(n), tp.soa_particles(n)
    for c in range(3):
        pos[:, c].copy_(0.05 + 0.9 * torch.rand(n, dtype=torch.float64, device="cuda", generator=gen))
        mom[:, c].copy_(torch.randn(n, dtype=torch.float64, device="cuda", generator=gen))
    sums = [int(a[:, c].contiguous().view(torch.int64).sum()) for a in (pos, mom) for c in range(3)]
    idx = subset(gen, n, 4096)
    x_before = pos[:, 0].clone()
    perm = tp.sort_by_cell(pos, mom, grid, return_perm=True)
    assert [int(a[:, c].contiguous().view(torch.int64).sum()) for a in (pos, mom) for c in range(3)] == sums
    assert int(perm.sum()) == n * (n - 1) // 2
    inv_dr = 1.0 / grid.dr
    keys = torch.clamp(((pos[:, 0] - float(grid.r[0])) * inv_dr).floor().to(torch.int64), 0, ng - 2)
    dk = keys[1:] - keys[:-1]
    assert int(dk.min()) >= 0                                               # ascending cells
    dp = perm[1:] - perm[:-1]
    assert bool(((dk > 0) | (dp > 0)).all())                                # stable inside a cell
    del keys, dk, dp
    assert torch.equal(pos[idx, 0], x_before[perm[idx]])                    # new[i] = old[perm[i]]


def test_config4_beam_256mi_particles():
    """configs[3]: 256 Mi particles, gamma log-uniform in [1, 1e3], B = 10 T, E = 0:
    |p| is conserved to rounding (pure rotation) and a subset matches the oracle."""
    import turbopy_b200 as tp
    n, B0 = 256 << 20, 10.0
    dt = 0.02 * (2 * np.pi * M_E / (abs(Q_E) * B0))
    gen = torch.Generator(device="cuda").manual_seed(4567)
    pos, mom = tp.soa_particles(n, fill=0.0), tp.soa_particles(n, fill=0.0)
    gam = torch.exp(torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * np.log(1e3))
    mom[:, 0].copy_(torch.sqrt(gam * gam - 1.0) * (M_E * C_LIGHT))
    del gam
    for c in (1, 2):
        mom[:, c].copy_(0.01 * mom[:, 0] * torch.randn(n, dtype=torch.float64, device="cuda", generator=gen))
    idx = subset(gen, n, 8192)
    sp, sm = pos[idx].cpu().numpy(), mom[idx].cpu().numpy()
    p0 = torch.sqrt(mom[:, 0] ** 2 + mom[:, 1] ** 2 + mom[:, 2] ** 2)
    tool = tp.BorisPush(make_owner(dt=dt), {"type": "BorisPush"})
    Bv = np.array([0.0, 0.0, B0])
    steps = 50
    for _ in range(steps):
        tool.push(pos, mom, Q_E, M_E, np.zeros(3), Bv)
    p1 = torch.sqrt(mom[:, 0] ** 2 + mom[:, 1] ** 2 + mom[:, 2] ** 2)
    assert float(((p1 - p0).abs() / p0).max()) < 5e-14
    del p0, p1
    for _ in range(steps):
        oboris.push_aos(sp, sm, Q_E, M_E, np.zeros(3), Bv, dt)
    assert_bits_equal(pos[idx].cpu().numpy(), sp); assert_bits_equal(mom[idx].cpu().numpy(), sm)


def test_config5_fd_256mi_point_grid():
    """configs[4]: centered / upwind ddx and radial del2 on a 256 Mi-point grid, chained
    as a field update f += dt * D f; sampled windows (+-1 halo) against the oracle."""
    import turbopy_b200 as tp
    n = 256 << 20
    r_max = 1.0
    step = 1.0 / (n - 1)
    # Grid.r without materialising a host grid object with ~10 arrays: the oracle's formula on the device
    lin = torch.arange(0, n, dtype=torch.float64, device="cuda") * step
    lin[-1] = 1.0
    r = 0.0 + (r_max - 0.0) * lin
    del lin
    r_host = r.cpu().numpy()
    from types import SimpleNamespace
    grid = SimpleNamespace(r=r_host, dr=(r_max - 0.0) / (n - 1), num_points=n, r_min=0.0, r_max=r_max,
                           cell_widths=r_host[1:] - r_host[:-1])
    tool = tp.FiniteDifference(make_owner(grid=grid), {"type": "FiniteDifference", "method": "centered"})
    f = torch.sin(2 * np.pi * r) + 0.1 * r * r
    f_host = f.cpu().numpy()
    D = tool.del2_radial()
    windows = [slice(0, 4096), slice(n // 2 - 2048, n // 2 + 2048), slice(n - 4096, n)]

    def check(got, fn):
        got = got.cpu().numpy()
        for w in windows:
            lo, hi = max(w.start - 1, 0), min(w.stop + 1, n)
            ref = fn(lo, hi)
            assert_bits_equal(got[w], ref[w.start - lo:w.stop - lo])

    check(tool.centered_difference(f), lambda lo, hi: ofd.centered_difference(f_host[lo:hi], grid.dr))
    check(tool.upwind_left(f), lambda lo, hi: ofd.upwind_left(f_host[lo:hi], grid.cell_widths[lo:hi - 1]))
    check(D @ f, lambda lo, hi: _del2_window(f_host, r_host, grid.dr, lo, hi, n))
    # chained update: two explicit steps f += dt * D f stay finite and local (window check of step 2)
    dtau = 1e-20
    f1 = f + dtau * (D @ f)
    f2 = f1 + dtau * (D @ f1)
    assert torch.isfinite(f2[1:-1]).all()
    f1h = f1.cpu().numpy()
    got = f2.cpu().numpy()
    for w in windows[1:2]:
        lo, hi = w.start - 1, w.stop + 1
        ref = f1h[lo:hi] + dtau * _del2_window(f1h, r_host, grid.dr, lo, hi, n)
        assert_bits_equal(got[w], ref[1:-1])


def _del2_window(f, r, dr, lo, hi, n):
    """Oracle on the slice [lo, hi): its first / last rows use the boundary formulas,
    which is only right when they ARE the global boundary rows -- the callers compare
    [start-lo : stop-lo] with lo = start-1, hi = stop+1 (clamped), which drops exactly
    the rows that are not."""
    return ofd.del2_radial_apply(f[lo:hi], r[lo:hi], dr)

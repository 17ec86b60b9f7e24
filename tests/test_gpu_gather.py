"""GPU parity of the linear field gather (Interpolators.interpolate1D, formulas
A/B/C) and of the fused gather+push kernel, bit-for-bit against the oracle and
the reference-generated golden vectors."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from helpers import M_E, Q_E, assert_bits_equal, make_grid, make_owner, thermal_ensemble
from oracle import boris as oboris
from oracle import gather as ogather

pytestmark = pytest.mark.gpu

FORMULA_FN = {"A": None, "B": ogather.interp_b, "C": ogather.interp_c}


@pytest.fixture(scope="module")
def tp():
    import turbopy_b200
    return turbopy_b200


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(f"{golden_dir}/gather.npz")


def oracle_gather(formula, xq, grid, y):
    if formula == "A":
        return ogather.interp_a(xq, grid.r, grid.dr, y, grid.r_min, grid.r_max)
    return FORMULA_FN[formula](xq, grid.r, y)


# ---------------------------------------------------------------- Interpolators.interpolate1D
@pytest.mark.parametrize("tag", ["n30", "n4097"])
def test_interpolate1d_golden(tp, golden, tag):
    tool = tp.Interpolators(make_owner(), {"type": "Interpolators"})
    r = torch.from_numpy(golden[f"{tag}_r"]).cuda()
    fields = golden[f"{tag}_fields"]
    xq = torch.from_numpy(golden[f"{tag}_xq"]).cuda()
    for c in range(6):                                   # 1-D y -> numpy.interp rounding (B)
        f = tool.interpolate1D(r, torch.from_numpy(fields[c]).cuda())
        assert_bits_equal(f(xq).cpu().numpy(), golden[f"{tag}_B"][c], f"B comp {c}")
    f = tool.interpolate1D(r, torch.from_numpy(fields).cuda())   # 2-D y -> _call_linear rounding (C)
    out = f(xq)
    assert out.shape == (6, xq.shape[0])
    assert_bits_equal(out.cpu().numpy(), golden[f"{tag}_C"], "C")
    # numpy in -> numpy out, as the reference's callers use it
    fn = tool.interpolate1D(golden[f"{tag}_r"], fields[0])
    got = fn(golden[f"{tag}_xq"])
    assert isinstance(got, np.ndarray)
    assert_bits_equal(got, golden[f"{tag}_B"][0])


def test_interpolate1d_reference_test(tp):
    """tests/test_computetools.py:71-79: f(x) == y at the nodes."""
    tool = tp.Interpolators(make_owner(), {"type": "Interpolator"})
    x = np.arange(0, 10, 1).astype(np.float64)
    y = np.exp(x)
    xnew = np.arange(0, 1, 0.1)
    f1 = tool.interpolate1D(x, y)
    assert np.allclose(f1(x), y)
    ref, _ = ogather.interp_b(xnew, x, y)
    assert_bits_equal(f1(xnew), ref)
    if tp.stock_tools.get("Interpolators") is None:      # no turboPy to hand spline kinds to
        with pytest.raises(NotImplementedError):
            tool.interpolate1D(x, y, 'quadratic')


def test_interpolate1d_bounds_error(tp):
    tool = tp.Interpolators(make_owner(), {"type": "Interpolators"})
    g = make_grid(30)
    f = tool.interpolate1D(g.r, g.r * 2.0)
    with pytest.raises(ValueError, match="above the interpolation range"):
        f(np.array([0.5, 1.0 + 1e-9]))
    with pytest.raises(ValueError, match="below the interpolation range"):
        f(np.array([-1e-9, 0.5]))
    assert_bits_equal(f(np.array([0.0, 1.0])), np.array([0.0, 2.0]))    # both ends are in range


def test_interpolate1d_nonuniform_grid(tp):
    """interpolate1D takes any sorted x: the uniform-grid index guess must fall
    back to the binary search."""
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(0, 1, 500)) ** 3
    x[0], x[-1] = 0.0, 1.0
    y = np.sin(7 * x)
    xq = rng.uniform(0, 1, 20000)
    tool = tp.Interpolators(make_owner(), {"type": "Interpolators"})
    ref, st = ogather.interp_b(xq, x, y)
    assert (st == 0).all()
    assert_bits_equal(tool.interpolate1D(x, y)(xq), ref)
    assert_bits_equal(tool.interpolate1D(x, y)(xq), np.interp(xq, x, y))


def test_pif_golden_formula_a(tp, golden_dir):
    """What the reference's particle_in_field goldens pin (e_0.5.csv): the
    create_interpolator gather of the EMWave field at x = 0.5."""
    pif = np.load(f"{golden_dir}/pif.npz")
    grid = make_grid(30)
    assert_bits_equal(grid.r, pif["r"])
    owner = make_owner(grid=grid)
    pusher = tp.BorisPush(owner, {"type": "BorisPush"})
    # gather only: push with q = 0 leaves momentum alone, so read E through a probe
    # particle at x = 0.5 whose momentum picks up q*E*dt exactly once.
    for step in (0, 1, 7, 20):
        Ey = torch.from_numpy(pif["fields"][step]).cuda()
        val, st = ogather.interp_a(np.array([0.5]), grid.r, grid.dr, pif["fields"][step])
        assert st[0] == 0 and val[0] == pif["e_at_half"][step]
        pos = torch.tensor([[0.5, 0.0, 0.0]], dtype=torch.float64, device="cuda")
        mom = torch.zeros((1, 3), dtype=torch.float64, device="cuda")
        pusher.gather_push(pos, mom, Q_E, M_E, E_grid=(None, Ey, None), formula="create_interpolator")
        rp = np.array([[0.5, 0.0, 0.0]]); rm = np.zeros((1, 3))
        oboris.push_aos(rp, rm, Q_E, M_E, np.array([0.0, pif["e_at_half"][step], 0.0]), np.zeros(3),
                        owner.clock.dt)
        assert_bits_equal(mom.cpu().numpy(), rm)
        assert_bits_equal(pos.cpu().numpy(), rp)
    pusher.check_status()


# ---------------------------------------------------------------- fused gather + push
def fused_reference(formula, pos0, mom0, grid, comps, axis, dt, steps):
    """Oracle: gather each present component at pos[:,axis], then push."""
    pos, mom = pos0.copy(), mom0.copy()
    for _ in range(steps):
        F = []
        for y in comps:
            if y is None:
                F.append(np.zeros(len(pos)))
            else:
                val, st = oracle_gather(formula, pos[:, axis], grid, y)
                assert (st == 0).all()
                F.append(val)
        E = np.stack(F[:3], axis=1); B = np.stack(F[3:], axis=1)
        oboris.push_aos(pos, mom, Q_E, M_E, E, B, dt)
    return pos, mom


def smooth_fields(grid, amp_e=1e5, amp_b=1.0):
    return [amp_e * np.cos(2 * np.pi * (k + 1) * grid.r) for k in range(3)] + \
           [amp_b * np.sin(2 * np.pi * (k + 1) * grid.r + 0.3) for k in range(3)]


@pytest.mark.parametrize("formula", ["A", "B", "C"])
@pytest.mark.parametrize("ng", [30, 1025, 4097, (1 << 20) + 1])
@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_fused_vs_oracle(tp, formula, ng, layout):
    """ng = 30 / 1025 stage everything in shared memory via TMA; 4097 x 6 components
    and the 1 Mi grid (config 3) take the L2 path."""
    if ng > 5000 and layout == "aos":
        pytest.skip("large grid covered by the SoA case")
    grid = make_grid(ng)
    dt = 1e-12
    owner = make_owner(dt=dt, grid=grid)
    tool = tp.BorisPush(owner, {"type": "BorisPush"})
    n = 20001
    pos0, mom0 = thermal_ensemble(n, seed=ng)
    pos0[:, 0] = np.random.default_rng(ng + 1).uniform(0.02, 0.98, n)
    pos0[:5, 0] = grid.r[[0, 1, ng // 2, ng - 2, ng - 1]]         # exactly on nodes, both ends
    mom0[0, 0], mom0[4, 0] = abs(mom0[0, 0]), -abs(mom0[4, 0])      # the two end particles drift inward
    comps = smooth_fields(grid)
    E = torch.from_numpy(np.stack(comps[:3], axis=1)).cuda()        # (ng,3) generate_field(3) layout
    Bt = tuple(torch.from_numpy(c).cuda() for c in comps[3:])       # three separate (ng,) arrays
    pos = tp.to_soa(pos0) if layout == "soa" else torch.from_numpy(pos0).cuda()
    mom = tp.to_soa(mom0) if layout == "soa" else torch.from_numpy(mom0).cuda()
    steps = 2
    for _ in range(steps):
        tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=Bt, axis=0, formula=formula)
    tool.check_status()
    rp, rm = fused_reference(formula, pos0, mom0, grid, comps, 0, dt, steps)
    assert_bits_equal(pos.cpu().numpy(), rp, "position")
    assert_bits_equal(mom.cpu().numpy(), rm, "momentum")


@pytest.mark.parametrize("formula", ["A", "B", "C"])
@pytest.mark.parametrize("ng", [513, (1 << 18) + 1])
@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_fused_tma_pipeline_vs_oracle(tp, formula, ng, layout):
    """Enough particles (> one 1024-particle tile per SM) for the TMA-pipelined
    kernel: several ring wrap-arounds per CTA, a ragged scalar tail, on-node and
    out-of-grid particles inside full tiles (exact path through shared memory)."""
    grid = make_grid(ng)
    dt = 1e-12
    tool = tp.BorisPush(make_owner(dt=dt, grid=grid), {"type": "BorisPush"})
    n = 148 * 1024 * 5 + 777
    pos0, mom0 = thermal_ensemble(n, seed=ng + 5)
    pos0[:, 0] = np.random.default_rng(ng).uniform(0.02, 0.98, n)
    on_node = np.arange(0, n, 9973)
    pos0[on_node, 0] = grid.r[(on_node * 7) % (ng - 2) + 1]
    outside = np.array([5, 100_003, n - 3])
    pos0[outside, 0] = [-0.25, 1.5, 1.0 + 1e-15]
    comps = smooth_fields(grid)
    E = torch.from_numpy(np.stack(comps[:3], axis=1)).cuda()
    B = torch.from_numpy(np.stack(comps[3:], axis=1)).cuda()
    if layout == "soa":
        pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
    else:                                            # numpy's (n,3) C order: the AoS tile mode of the ring
        pos, mom = torch.from_numpy(pos0).cuda(), torch.from_numpy(mom0).cuda()
    tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B, axis=0, formula=formula)
    with pytest.raises((ValueError, AssertionError), match="3 particle"):
        tool.check_status()
    keep = np.ones(n, dtype=bool); keep[outside] = False
    rp, rm = fused_reference(formula, pos0[keep], mom0[keep], grid, comps, 0, dt, 1)
    gp, gm = pos.cpu().numpy(), mom.cpu().numpy()
    assert_bits_equal(gp[keep], rp, "position"); assert_bits_equal(gm[keep], rm, "momentum")
    assert_bits_equal(gp[outside], pos0[outside]); assert_bits_equal(gm[outside], mom0[outside])


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_fused_partial_components_and_axis(tp, axis):
    """The pif shape: only E_y on the grid, B absent (None), any gather axis."""
    grid = make_grid(4097)
    dt = 5e-10
    tool = tp.BorisPush(make_owner(dt=dt, grid=grid), {"type": "BorisPush"})
    n = 5000
    pos0, mom0 = thermal_ensemble(n, seed=axis)
    mom0[:] = 0.0
    k = 2e8 / 2.998e8
    Ey = np.cos(2 * np.pi * (k * (grid.r - 0.5)))
    pos = tp.to_soa(pos0); mom = tp.to_soa(mom0)
    tool.gather_push(pos, mom, Q_E, M_E, E_grid=(None, torch.from_numpy(Ey).cuda(), None), B_grid=None, axis=axis)
    tool.check_status()
    rp, rm = fused_reference("B", pos0, mom0, grid, [None, Ey, None, None, None, None], axis, dt, 1)
    assert_bits_equal(pos.cpu().numpy(), rp); assert_bits_equal(mom.cpu().numpy(), rm)


def test_fused_out_of_range_particles_flagged(tp):
    grid = make_grid(30)
    tool = tp.BorisPush(make_owner(grid=grid), {"type": "BorisPush"})
    pos0, mom0 = thermal_ensemble(1000, seed=9)
    pos0[17, 0] = 1.0 + 1e-12
    pos0[400, 0] = -1e-300
    pos = tp.to_soa(pos0); mom = tp.to_soa(mom0)
    Ey = torch.ones(30, dtype=torch.float64, device="cuda")
    tool.gather_push(pos, mom, Q_E, M_E, E_grid=(None, Ey, None))
    with pytest.raises(ValueError, match="first index 17"):
        tool.check_status()
    tool.check_status()                                           # flags were reset
    out_p, out_m = pos.cpu().numpy(), mom.cpu().numpy()
    for i in (17, 400):                                           # flagged particles untouched
        assert_bits_equal(out_p[i], pos0[i]); assert_bits_equal(out_m[i], mom0[i])
    assert not np.array_equal(out_m[18], mom0[18])


def test_fused_host_path(tp):
    """End-to-end: numpy particles and numpy grid fields through the C ABI's host call."""
    grid = make_grid(1025)
    dt = 1e-12
    tool = tp.BorisPush(make_owner(dt=dt, grid=grid), {"type": "BorisPush"})
    n = 300_001
    pos0, mom0 = thermal_ensemble(n, seed=21)
    comps = smooth_fields(grid)
    E = np.stack(comps[:3], axis=1)
    pos, mom = pos0.copy(), mom0.copy()
    tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=tuple(comps[3:]))
    rp, rm = fused_reference("B", pos0, mom0, grid, comps, 0, dt, 1)
    assert_bits_equal(pos, rp); assert_bits_equal(mom, rm)
    pos[5, 0] = 2.0
    with pytest.raises(ValueError):
        tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=None)


@pytest.mark.parametrize("formula", ["A", "B", "C"])
@pytest.mark.parametrize("ng,n", [(30, 3001), (1025, 148 * 960 * 2 + 55), ((1 << 17) + 1, 148 * 960 + 9)])
def test_gather_push_steps_equals_repeated_calls(tp, formula, ng, n):
    """gather_push_steps(K) == K x gather_push on static grid fields, bit for bit (register path, staged ring,
    L2 records), including particles that walk off the grid during the pass: they are flagged and keep the
    state they left with -- exactly what K single-step calls leave behind."""
    grid = make_grid(ng)
    dt, K = 2e-11, 5
    tool_a = tp.BorisPush(make_owner(dt=dt, grid=grid), {"type": "BorisPush"})
    tool_b = tp.BorisPush(make_owner(dt=dt, grid=grid), {"type": "BorisPush"})
    pos0, mom0 = thermal_ensemble(n, seed=ng + 11)
    pos0[:, 0] = np.random.default_rng(ng).uniform(0.05, 0.95, n)
    fast = np.arange(3, n, max(7, n // 40))                       # relativistic particles heading for the right edge
    pos0[fast, 0] = 0.995
    mom0[fast] = 0.0
    mom0[fast, 0] = 30.0 * M_E * 2.9979e8
    comps = smooth_fields(grid)
    E = torch.from_numpy(np.stack(comps[:3], axis=1)).cuda()
    B = torch.from_numpy(np.stack(comps[3:], axis=1)).cuda()
    name = {"A": "create_interpolator", "B": "interp", "C": "interp1d_nd"}[formula]
    pa, ma = tp.to_soa(pos0), tp.to_soa(mom0)
    pb, mb = tp.to_soa(pos0), tp.to_soa(mom0)
    tool_a.gather_push_steps(pa, ma, Q_E, M_E, K, E_grid=E, B_grid=B, formula=name)
    left = 0
    for _ in range(K):
        tool_b.gather_push(pb, mb, Q_E, M_E, E_grid=E, B_grid=B, formula=name)
        try:
            tool_b.check_status()
        except (ValueError, AssertionError):
            left += 1
    assert left > 0                                               # the fast particles did leave during the pass
    assert_bits_equal(pa.cpu().numpy(), pb.cpu().numpy(), "position")
    assert_bits_equal(ma.cpu().numpy(), mb.cpu().numpy(), "momentum")
    gone = (pb.cpu().numpy()[:, 0] > 1.0)
    with pytest.raises((ValueError, AssertionError), match=f"{int(gone.sum())} particle"):
        tool_a.check_status()


@pytest.mark.parametrize("ng", [2, 30, 4097])
def test_interpolate1d_linspace_abscissas_same_bits(tp, ng, monkeypatch):
    """interpolate1D over turboPy's own grid coordinates rebuilds x[j], x[j+1] in registers; over any other
    abscissas (or with TURBOPY_B200_LINSPACE=0) it loads them.  Same bits, 1-D and 2-D y, on-node / end / NaN points."""
    grid = make_grid(ng, -0.75, 2.25)
    tool = tp.Interpolators(make_owner(grid=grid), {"type": "Interpolators"})
    rng = np.random.default_rng(ng)
    y = np.cos(3 * grid.r) * 1e3
    y2 = np.stack([y, -y, y * y])
    xq = np.concatenate([rng.uniform(grid.r[0], grid.r[-1], 200_001), grid.r[rng.integers(0, ng, 50)],
                         [grid.r[0], grid.r[-1], np.nan]])
    f1, f2 = tool.interpolate1D(grid.r, y), tool.interpolate1D(grid.r, y2)
    assert f1._lin is not None and f2._lin is not None
    monkeypatch.setenv("TURBOPY_B200_LINSPACE", "0")
    g1, g2 = tool.interpolate1D(grid.r, y), tool.interpolate1D(grid.r, y2)
    assert g1._lin is None and g2._lin is None
    with np.errstate(all="ignore"):
        want1, _ = ogather.interp_b(xq, grid.r, y)
        want2, _ = ogather.interp_c(xq, grid.r, y2)
    for got, want in ((f1(xq), want1), (g1(xq), want1), (f2(xq), want2), (g2(xq), want2)):
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert_bits_equal(np.nan_to_num(got, nan=-7.0), np.nan_to_num(want, nan=-7.0))
    monkeypatch.delenv("TURBOPY_B200_LINSPACE")
    assert tool.interpolate1D(grid.r ** 3, y)._lin is None or ng == 2        # stretched abscissas: not a linspace


@pytest.mark.parametrize("formula", ["A", "B", "C"])
@pytest.mark.parametrize("shape", ["ey_only", "e_only", "ey_bz"])
@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_mask_specialised_kernels_vs_oracle(tp, formula, shape, layout, monkeypatch):
    """The staged TMA kernel has compile-time instances for the common component sets (0x02 = E_y only, the pif shape;
    0x07 = E only); any other set (here E_y + B_z) runs the generic kernel.  Same bits as the oracle, with the
    specialised kernels switched on and off, incl. on-node / out-of-grid particles inside full tiles."""
    ng = 513
    grid = make_grid(ng)
    dt = 2e-12
    n = 148 * 960 * 2 + 77
    pos0, mom0 = thermal_ensemble(n, seed=ng + len(shape))
    pos0[:, 0] = np.random.default_rng(ng).uniform(0.02, 0.98, n)
    pos0[np.arange(0, n, 9973), 0] = grid.r[(np.arange(0, n, 9973) * 7) % (ng - 2) + 1]
    comps = smooth_fields(grid)
    keep = {"ey_only": [1], "e_only": [0, 1, 2], "ey_bz": [1, 5]}[shape]
    comps = [c if k in keep else None for k, c in enumerate(comps)]
    dev = [None if c is None else torch.from_numpy(c).cuda() for c in comps]
    E = tuple(dev[:3]); B = None if all(c is None for c in dev[3:]) else tuple(dev[3:])
    rp, rm = fused_reference(formula, pos0, mom0, grid, comps, 0, dt, 2)
    for switch in ("1", "0"):
        monkeypatch.setenv("TURBOPY_B200_MASK_KERNELS", switch)   # (read once per process: the second pass re-checks the default)
        tool = tp.BorisPush(make_owner(dt=dt, grid=grid), {"type": "BorisPush"})
        pos = tp.to_soa(pos0) if layout == "soa" else torch.from_numpy(pos0).cuda()
        mom = tp.to_soa(mom0) if layout == "soa" else torch.from_numpy(mom0).cuda()
        for _ in range(2):
            tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B, formula={"A": "create_interpolator", "B": "interp", "C": "interp1d_nd"}[formula])
        tool.check_status()
        assert_bits_equal(pos.cpu().numpy(), rp, f"position ({switch})")
        assert_bits_equal(mom.cpu().numpy(), rm, f"momentum ({switch})")


@pytest.mark.parametrize("formula", ["A", "B", "C"])
@pytest.mark.parametrize("gridkind", ["linspace01", "linspace_offset", "jitter_small", "jitter_large"])
@pytest.mark.parametrize("n", [7001, 148 * 960 * 2 + 301])
def test_formula_a_window_margin_edges(tp, gridkind, n, formula):
    """Formula A's fast path decides 'the window holds exactly nodes j, j+1' from a margin around the two nodes
    instead of loading r[j-1] and r[j+2] (csrc/gather_math.cuh: window_margin).  Particles a few ulps / a few 1e-16 dr
    away from a node sit inside that margin and must take the complete semantics; on a grid whose cells are
    narrower than Grid.dr the window can hold three nodes (core.py:729 assert -> flagged, untouched); with large
    jitter the margin is switched off.  Register-path (small n) and TMA kernel (large n), staged grids.
    Formulas B and C ride along on the same points and grids (no window, nothing flagged)."""
    if formula != "A" and gridkind == "linspace_offset":
        pytest.skip("formula A only")
    ng = 1025
    rng = np.random.default_rng(n + len(gridkind))
    if gridkind == "linspace01":
        grid = make_grid(ng)
    elif gridkind == "linspace_offset":
        grid = make_grid(ng, -3.7, 11.3)
    else:
        grid = make_grid(ng)
        amp = 1e-3 if gridkind == "jitter_small" else 0.2
        r = grid.r.copy()
        r[1:-1] += amp * grid.dr * rng.uniform(-1, 1, ng - 2)
        grid.r = r
        grid.cell_widths = r[1:] - r[:-1]
    span = grid.r[-1] - grid.r[0]
    dt = 1e-13
    tool = tp.BorisPush(make_owner(dt=dt, grid=grid), {"type": "BorisPush"})
    pos0, mom0 = thermal_ensemble(n, seed=n)
    pos0[:, 0] = grid.r[0] + span * rng.uniform(0.01, 0.99, n)
    # every third particle next to a node: k ulps or a tiny fraction of dr off, both sides
    near = np.arange(0, n, 3)
    j = rng.integers(1, ng - 1, len(near))
    kind = rng.integers(0, 3, len(near))
    node = grid.r[j]
    ulps = rng.choice([1, 2, 3, 5, 8, 16, 100, 10_000], len(near)) * rng.choice([-1, 1], len(near))
    frac = rng.choice([1e-15, 1e-14, 1e-13, 1e-12, 1e-10, 1e-6, 2e-3, 1.6e-2], len(near)) * rng.choice([-1, 1], len(near))
    x_ulp = node + ulps * np.spacing(np.abs(node))
    x_frac = node + frac * grid.dr
    pos0[near, 0] = np.where(kind == 0, node, np.where(kind == 1, x_ulp, x_frac))
    comps = smooth_fields(grid)
    _, st = oracle_gather(formula, pos0[:, 0], SimpleNamespace(r=grid.r, dr=grid.dr, r_min=grid.r[0], r_max=grid.r[-1]), comps[0])
    bad = st != 0
    if gridkind.startswith("linspace") or formula != "A":
        assert not bad.any()
    E = torch.from_numpy(np.stack(comps[:3], axis=1)).cuda()
    B = torch.from_numpy(np.stack(comps[3:], axis=1)).cuda()
    pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
    tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B, axis=0, formula=formula)
    if bad.any():
        with pytest.raises((ValueError, AssertionError), match=rf"^{int(bad.sum())} particle"):
            tool.check_status()
    else:
        tool.check_status()
    rp, rm = fused_reference(formula, pos0[~bad], mom0[~bad], grid, comps, 0, dt, 1)
    gp, gm = pos.cpu().numpy(), mom.cpu().numpy()
    assert_bits_equal(gp[~bad], rp, "position"); assert_bits_equal(gm[~bad], rm, "momentum")
    assert_bits_equal(gp[bad], pos0[bad]); assert_bits_equal(gm[bad], mom0[bad])


@pytest.mark.parametrize("gridkind", ["linspace", "stretched"])
@pytest.mark.parametrize("ng", [30, 1025, 4097])
def test_interpolate1d_ring_kernel_vs_numpy(tp, gridkind, ng):
    """1-D y with enough points for the TMA-ring kernel (interp_cells_ring_kernel: >= 4 tiles of 1920 points per
    SM): linspace abscissas (x rebuilt in registers) and stretched ones ({x0, x1} table), points on nodes / at both
    ends / one ulp off nodes inside full tiles (exact path patches the stage slot), NaN abscissas, a ragged tail;
    bit-identical to numpy.interp and to the register-path kernel; out-of-range points raise like interp1d."""
    rng = np.random.default_rng(ng)
    grid = make_grid(ng, -0.5, 1.5)
    x = grid.r.copy()
    if gridkind == "stretched":
        x = -0.5 + 2.0 * np.linspace(0.0, 1.0, ng) ** 1.7
    y = np.sin(5.0 * x) + 0.1 * x
    n = 148 * 1920 * 5 + 1237
    xq = rng.uniform(x[0], x[-1], n)
    idx = rng.integers(0, n, 20_000)
    node = x[rng.integers(0, ng, len(idx))]
    xq[idx] = np.where(rng.random(len(idx)) < 0.5, node, np.clip(np.nextafter(node, rng.choice([-np.inf, np.inf], len(idx))),
                                                                  x[0], x[-1]))
    xq[[0, 1, n - 1, n - 2]] = [x[0], x[-1], x[0], x[-1]]
    nan_at = rng.integers(0, n, 50)
    xq[nan_at] = np.nan
    tool = tp.Interpolators(make_owner(), {"type": "Interpolators"})
    f = tool.interpolate1D(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda())
    got = f(torch.from_numpy(xq).cuda()).cpu().numpy()
    ref = np.interp(xq, x, y)
    ref[nan_at] = np.nan                                          # numpy.interp hands a NaN abscissa back
    assert_bits_equal(np.isnan(got), np.isnan(ref), "NaN pattern")
    ok = ~np.isnan(ref)
    assert_bits_equal(got[ok], ref[ok], "ring kernel vs numpy.interp")
    # unaligned query view -> not bulk-copyable -> the register-path kernel; same bits
    xq_dev = torch.empty(n + 1, dtype=torch.float64, device="cuda")[1:]
    xq_dev.copy_(torch.from_numpy(xq))
    got2 = f(xq_dev).cpu().numpy()
    assert_bits_equal(got2[ok], ref[ok], "register-path kernel vs numpy.interp")
    xq[12345] = x[-1] + 1e-9
    with pytest.raises(ValueError, match="above the interpolation range"):
        f(torch.from_numpy(xq).cuda())


@pytest.mark.parametrize("gridkind", ["linspace", "stretched"])
@pytest.mark.parametrize("ng,k", [(30, 2), (1025, 6), (1025, 7), (2049, 3), (4097, 6), (5001, 2), (9001, 3)])
def test_interpolate1d_nd_cell_planes_vs_scipy(tp, gridkind, ng, k):
    """(k, ng) y with many more points than cells: interp_cells_nc_kernel ({y[j], y[j+1]} planes in shared memory, as
    many components per pass as fit, at most six: k = 7 takes two passes, (4097, 6) two on a linspace grid and three on
    a stretched one, (5001, 2) stretched one component per pass; (9001, 3) stays on interp_kernel).  Points on
    nodes (searchsorted's LEFT cell), at both ends, one ulp off nodes, NaN abscissas, a ragged tail and an unaligned
    query view; bit-identical to scipy's interp1d and to the oracle's restatement of _call_linear."""
    from scipy.interpolate import interp1d
    rng = np.random.default_rng(ng * 10 + k)
    grid = make_grid(ng, -0.5, 1.5)
    x = grid.r.copy()
    if gridkind == "stretched":
        x = -0.5 + 2.0 * np.linspace(0.0, 1.0, ng) ** 1.7
    y = np.stack([np.sin((c + 2.0) * x) * 10.0 ** (c - 2) + 0.1 * c * x for c in range(k)])
    n = 300_003
    xq = rng.uniform(x[0], x[-1], n)
    idx = rng.integers(0, n, 5_000)
    node = x[rng.integers(0, ng, len(idx))]
    xq[idx] = np.where(rng.random(len(idx)) < 0.5, node, np.clip(np.nextafter(node, rng.choice([-np.inf, np.inf], len(idx))),
                                                                  x[0], x[-1]))
    xq[[0, 1, n - 1, n - 2]] = [x[0], x[-1], x[0], x[-1]]
    nan_at = rng.integers(0, n, 20)
    xq[nan_at] = np.nan
    tool = tp.Interpolators(make_owner(grid=grid), {"type": "Interpolators"})
    f = tool.interpolate1D(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda())
    assert (f._lin is not None) == (gridkind == "linspace")
    got = f(torch.from_numpy(xq).cuda()).cpu().numpy()
    assert got.shape == (k, n)
    with np.errstate(all="ignore"):
        ref = interp1d(x, y)(xq)
        want, _ = ogather.interp_c(xq, x, y)
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.isnan(got[:, nan_at]).all()
    assert_bits_equal(np.nan_to_num(got, nan=-7.0), np.nan_to_num(ref, nan=-7.0), "vs scipy interp1d")
    assert_bits_equal(np.nan_to_num(got, nan=-7.0), np.nan_to_num(want, nan=-7.0), "vs the oracle")
    xq_dev = torch.empty(n + 1, dtype=torch.float64, device="cuda")[1:]          # 8-byte aligned only: scalar accesses
    xq_dev.copy_(torch.from_numpy(xq))
    assert_bits_equal(np.nan_to_num(f(xq_dev).cpu().numpy(), nan=-7.0), np.nan_to_num(ref, nan=-7.0), "unaligned queries")
    xq[777] = x[0] - 1e-9
    with pytest.raises(ValueError, match="below the interpolation range"):
        f(torch.from_numpy(xq).cuda())

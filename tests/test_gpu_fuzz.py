"""Seeded differential fuzzing of the two particle entry points against the oracle: random sizes (register
path and TMA rings), layouts (SoA, C-order, strided views), field modes, grids (uniform / non-uniform,
staged / L2), formulas, component subsets, axes, on-node and out-of-grid particles.  Bit for bit."""
import re

import numpy as np
import pytest
import torch

from helpers import M_E, Q_E, assert_bits_equal, make_grid, make_owner
from oracle import boris as oboris
from oracle import gather as ogather

pytestmark = pytest.mark.gpu
C_LIGHT = 2.9979e8


def device_array(tp, a, kind):
    """(n,3) numpy array -> CUDA tensor in one of the layouts the C ABI accepts"""
    if kind == "soa":
        return tp.to_soa(a)
    if kind == "aos":
        return torch.from_numpy(np.ascontiguousarray(a)).cuda()
    wide = torch.zeros((a.shape[0], 5), dtype=torch.float64, device="cuda")      # padded rows: generic strides
    wide[:, 1:4] = torch.from_numpy(a).cuda()
    return wide[:, 1:4]


def random_particles(rng, n):
    pos = rng.uniform(0.0, 1.0, (n, 3))
    scale = 10.0 ** rng.uniform(-3, 3)
    mom = rng.normal(0, scale * M_E * C_LIGHT, (n, 3))
    if n > 3:
        mom[rng.integers(0, n, max(1, n // 50))] = 0.0
        mom[rng.integers(0, n)] = 1e-200
        mom[rng.integers(0, n)] = rng.normal(0, 1e-310, 3)                        # denormals
    return pos, mom


def random_field(rng, tp, n, name):
    mode = rng.choice(["uni_np", "uni_dev", "uni_13", "pp_soa", "pp_aos", "pp_strided", "zero"])
    amp = (10.0 ** rng.uniform(2, 7)) if name == "E" else (10.0 ** rng.uniform(-3, 1.5))
    if mode == "zero":
        v = np.zeros(3)
        return v, v
    if mode.startswith("uni"):
        v = rng.normal(0, amp, 3)
        dev = v if mode == "uni_np" else torch.from_numpy(v).cuda() if mode == "uni_dev" else torch.from_numpy(v.reshape(1, 3)).cuda()
        return v, dev
    v = rng.normal(0, amp, (n, 3))
    return v, device_array(tp, v, mode[3:])


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_push(seed):
    import turbopy_b200 as tp
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 3000)) if seed % 4 else int(rng.integers(148 * 480, 148 * 480 * 3))
    dt = 10.0 ** rng.uniform(-14, -10)
    tool = tp.BorisPush(make_owner(dt=dt), {"type": "BorisPush"})
    pos0, mom0 = random_particles(rng, n)
    layout = rng.choice(["soa", "aos", "strided"])
    E, Ed = random_field(rng, tp, n, "E")
    B, Bd = random_field(rng, tp, n, "B")
    pos, mom = device_array(tp, pos0, layout), device_array(tp, mom0, layout)
    steps = int(rng.integers(1, 4))
    if rng.random() < 0.5:
        tool.push_steps(pos, mom, Q_E, M_E, Ed, Bd, steps)
    else:
        for _ in range(steps):
            tool.push(pos, mom, Q_E, M_E, Ed, Bd)
    rp, rm = pos0.copy(), mom0.copy()
    for _ in range(steps):
        oboris.push_aos(rp, rm, Q_E, M_E, E, B, dt)
    assert_bits_equal(pos.cpu().numpy(), rp, f"position (n={n}, {layout})")
    assert_bits_equal(mom.cpu().numpy(), rm, f"momentum (n={n}, {layout})")


@pytest.mark.parametrize("seed", range(60))
def test_fuzz_gather_push(seed):
    import turbopy_b200 as tp
    rng = np.random.default_rng(5000 + seed)
    ng = int(rng.choice([2, 3, 17, 30, 257, 1025, 4097, 9001, 70_001]))
    formula = str(rng.choice(["A", "B", "C"]))
    grid = make_grid(ng, float(rng.uniform(-2, 0)), float(rng.uniform(0.5, 3)))
    if formula != "A" and ng > 3 and rng.random() < 0.3:                          # non-uniform coordinates
        r = np.sort(rng.uniform(grid.r[0], grid.r[-1], ng))
        r[0], r[-1] = grid.r[0], grid.r[-1]
        if np.all(np.diff(r) > 0):
            grid.r = r
    n = int(rng.integers(1, 4000)) if seed % 3 else int(rng.integers(148 * 960, 148 * 960 * 2))
    axis = int(rng.integers(0, 3))
    dt = 10.0 ** rng.uniform(-14, -11)
    tool = tp.BorisPush(make_owner(dt=dt, grid=grid), {"type": "BorisPush"})
    pos0, mom0 = random_particles(rng, n)
    span = grid.r[-1] - grid.r[0]
    pos0[:, axis] = grid.r[0] + span * rng.uniform(0.0, 1.0, n)
    k = max(1, n // 40)
    pos0[rng.integers(0, n, k), axis] = grid.r[rng.integers(0, ng, k)]            # exactly on nodes (ends included)
    outside = np.unique(rng.integers(0, n, max(1, n // 200))) if rng.random() < 0.5 else np.array([], dtype=np.int64)
    pos0[outside, axis] = np.where(rng.random(len(outside)) < 0.5, grid.r[0] - span * 0.01, grid.r[-1] + span * 0.01)
    comps = [None] * 6
    for c in range(6):
        if rng.random() < 0.6:
            comps[c] = (1e5 if c < 3 else 1.0) * np.cos(rng.uniform(1, 9) * grid.r + rng.uniform(0, 6))
    as_block = rng.random() < 0.5

    def lower(three):
        if all(c is None for c in three):
            return None
        if as_block:                                                              # (ng,3) generate_field(3) layout
            return torch.from_numpy(np.stack([c if c is not None else np.zeros(ng) for c in three], axis=1)).cuda()
        return tuple(None if c is None else torch.from_numpy(c).cuda() for c in three)

    layout = rng.choice(["soa", "aos", "strided"])
    pos, mom = device_array(tp, pos0, layout), device_array(tp, mom0, layout)
    tool.gather_push(pos, mom, Q_E, M_E, E_grid=lower(comps[:3]), B_grid=lower(comps[3:]), axis=axis,
                     formula={"A": "create_interpolator", "B": "interp", "C": "interp1d_nd"}[formula])
    # oracle: status per particle from the coordinate alone; flagged particles stay untouched
    probe = np.zeros(ng)
    if formula == "A":
        _, st = ogather.interp_a(pos0[:, axis], grid.r, grid.dr, probe, grid.r[0], grid.r[-1])
    else:
        _, st = (ogather.interp_b if formula == "B" else ogather.interp_c)(pos0[:, axis], grid.r, probe)
    bad = st != 0
    if bad.any():
        with pytest.raises((ValueError, AssertionError)) as info:
            tool.check_status()
        assert int(re.match(r"(\d+) particle", str(info.value)).group(1)) == int(bad.sum())
        assert f"first index {int(np.flatnonzero(bad)[0])}" in str(info.value)
    else:
        tool.check_status()
    good = ~bad
    F = []
    for y in comps:
        if y is None or (as_block is False and y is None):
            F.append(np.zeros(int(good.sum())))
            continue
        if formula == "A":
            v, s2 = ogather.interp_a(pos0[good, axis], grid.r, grid.dr, y, grid.r[0], grid.r[-1])
        else:
            v, s2 = (ogather.interp_b if formula == "B" else ogather.interp_c)(pos0[good, axis], grid.r, y)
        assert (s2 == 0).all()
        F.append(v)
    rp, rm = pos0[good].copy(), mom0[good].copy()
    oboris.push_aos(rp, rm, Q_E, M_E, np.stack(F[:3], axis=1), np.stack(F[3:], axis=1), dt)
    gp, gm = pos.cpu().numpy(), mom.cpu().numpy()
    assert_bits_equal(gp[good], rp, f"position (ng={ng}, n={n}, {formula}, {layout})")
    assert_bits_equal(gm[good], rm, f"momentum (ng={ng}, n={n}, {formula}, {layout})")
    assert_bits_equal(gp[bad], pos0[bad]); assert_bits_equal(gm[bad], mom0[bad])


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_finite_difference(seed):
    """Every FiniteDifference stencil / operator on random grid sizes and operand alignments (a slice that
    starts 8 bytes into an allocation takes the scalar-access path)."""
    import turbopy_b200 as tp
    from oracle import fd as ofd
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.integers(3, 40)) if seed % 3 == 0 else int(rng.integers(40, 200_000))
    rmax = float(10.0 ** rng.uniform(-1, 2))
    grid = make_grid(n, 0.0, rmax)
    tool = tp.FiniteDifference(make_owner(grid=grid), {"type": "FiniteDifference", "method": "centered"})
    y = rng.normal(size=n) * 10.0 ** rng.uniform(-3, 3)
    if seed >= 12:                                            # non-finite and extreme operands
        for v in (np.inf, -np.inf, np.nan, 1e308, 5e-324, -0.0):
            y[rng.integers(0, n)] = v
    buf = torch.zeros(n + 3, dtype=torch.float64, device="cuda")
    off = int(rng.integers(0, 3))
    yd = buf[off:off + n]
    yd.copy_(torch.from_numpy(y))
    nn = lambda a: np.nan_to_num(np.asarray(a), nan=-7.0, posinf=7e300, neginf=-7e300)   # noqa: E731
    with np.errstate(all="ignore"):
        assert_bits_equal(nn(tool.centered_difference(yd).cpu().numpy()), nn(ofd.centered_difference(y, grid.dr)), "centered")
        assert_bits_equal(nn(tool.upwind_left(yd).cpu().numpy()), nn(ofd.upwind_left(y, grid.cell_widths)), "upwind")
    with np.errstate(all="ignore"):
        assert_bits_equal(nn((tool.del2_radial() @ yd).cpu().numpy()), nn(ofd.del2_radial_apply(y, grid.r, grid.dr)), "del2_radial")
        for name in ofd.BANDED_NAMES:
            if name in ("BC_left_extrap", "BC_left_avg", "BC_left_quad", "BC_right_extrap") and n < 4:
                continue
            data, offsets = ofd.banded_operator(name, grid.r, grid.dr)
            got = (getattr(tool, name)() @ yd).cpu().numpy()
            assert_bits_equal(nn(got), nn(ofd.dia_apply(data, offsets, y)), f"{name} n={n} off={off}")
    assert float(buf[:off].abs().sum()) == 0.0 and float(buf[off + n:].abs().sum()) == 0.0


@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_non_finite_inputs_follow_ieee_like_numpy(layout):
    """Infinite / NaN momenta, positions and fields take the recovery path (plain IEEE operations in the
    reference's order): the same entries come out NaN / +-inf as in numpy, every finite entry bit-identical.
    (NaN sign and payload are not compared: x86 and the GPU produce different default NaNs.)"""
    import turbopy_b200 as tp
    rng = np.random.default_rng(77)
    n = 148 * 480 * 2 + 91
    dt = 1e-12
    tool = tp.BorisPush(make_owner(dt=dt), {"type": "BorisPush"})
    pos0, mom0 = random_particles(rng, n)
    E = rng.normal(0, 1e5, (n, 3)); B = rng.normal(0, 1.0, (n, 3))
    idx = rng.choice(n, 600, replace=False)
    specials = [np.inf, -np.inf, np.nan, 1e308, -1e308, 5e-324, 0.0, -0.0]
    for k, i in enumerate(idx):
        target = (mom0, pos0, E, B)[k % 4]
        target[i, (k // 4) % 3] = specials[(k // 12) % len(specials)]
    pos, mom = device_array(tp, pos0, layout), device_array(tp, mom0, layout)
    tool.push(pos, mom, Q_E, M_E, device_array(tp, E, "aos"), device_array(tp, B, "aos"))
    rp, rm = pos0.copy(), mom0.copy()
    with np.errstate(all="ignore"):
        oboris.push_aos(rp, rm, Q_E, M_E, E, B, dt)
    for got, want, what in ((pos.cpu().numpy(), rp, "position"), (mom.cpu().numpy(), rm, "momentum")):
        assert np.array_equal(np.isnan(got), np.isnan(want)), what
        ok = ~np.isnan(want)
        assert_bits_equal(got[ok], want[ok], what)
    assert np.isnan(rm).any() and np.isinf(rp[~np.isnan(rp)]).any()       # the test did exercise both


@pytest.mark.parametrize("formula", ["B", "C"])
@pytest.mark.parametrize("ng", [257, 70_001])
def test_nan_positions_in_the_fused_gather(formula, ng):
    """A NaN coordinate is not "outside the grid" for interp1d (no flag); numpy.interp / scipy hand back NaN fields,
    so that particle's state turns NaN exactly like in the reference while its neighbours are untouched by it."""
    import turbopy_b200 as tp
    rng = np.random.default_rng(ng)
    n = 148 * 960 + 321
    dt = 1e-12
    grid = make_grid(ng)
    tool = tp.BorisPush(make_owner(dt=dt, grid=grid), {"type": "BorisPush"})
    pos0, mom0 = random_particles(rng, n)
    pos0[:, 0] = rng.uniform(0.02, 0.98, n)
    nan_idx = rng.choice(n, 50, replace=False)
    pos0[nan_idx, 0] = np.nan
    comps = [1e5 * np.cos((k + 1) * grid.r) for k in range(3)] + [np.sin((k + 2) * grid.r) for k in range(3)]
    E = torch.from_numpy(np.stack(comps[:3], axis=1)).cuda(); B = torch.from_numpy(np.stack(comps[3:], axis=1)).cuda()
    pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
    tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B, formula={"B": "interp", "C": "interp1d_nd"}[formula])
    tool.check_status()                                           # nothing flagged
    fn = ogather.interp_b if formula == "B" else ogather.interp_c
    F = [fn(pos0[:, 0], grid.r, c)[0] for c in comps]
    rp, rm = pos0.copy(), mom0.copy()
    with np.errstate(all="ignore"):
        oboris.push_aos(rp, rm, Q_E, M_E, np.stack(F[:3], axis=1), np.stack(F[3:], axis=1), dt)
    for got, want in ((pos.cpu().numpy(), rp), (mom.cpu().numpy(), rm)):
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        assert_bits_equal(got[ok], want[ok])
    assert np.isnan(rm[nan_idx]).all()


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_interpolate1d_awkward_points(seed):
    """interpolate1D on the GPU against the oracle (itself pinned to numpy / scipy): infinite node values
    (numpy.interp's NaN fallbacks), NaN abscissas, points exactly on and one ulp off the nodes, both ends,
    uniform and non-uniform grids, 1-D and 2-D y."""
    import turbopy_b200 as tp
    rng = np.random.default_rng(300 + seed)
    ng = int(rng.choice([2, 3, 11, 30, 257, 1025, 20_001]))
    lo, hi = float(rng.uniform(-3, 0)), float(rng.uniform(0.1, 4))
    xp = make_grid(ng, lo, hi).r
    if seed % 3 == 0 and ng > 3:
        xp = np.sort(np.concatenate([[lo, hi], rng.uniform(lo, hi, ng - 2)]))
        if len(np.unique(xp)) != ng:
            xp = make_grid(ng, lo, hi).r
    lo, hi = float(xp[0]), float(xp[-1])                      # the grid's own end points (r_max is rounded into r[-1])
    y = np.cos(3 * xp) * 10.0 ** rng.uniform(-2, 5)
    if seed % 2 and ng > 3:
        y[rng.integers(0, ng)] = np.inf
        y[rng.integers(0, ng)] = -np.inf
    xq = np.concatenate([rng.uniform(lo, hi, 5000), xp[rng.integers(0, ng, 200)], [lo, hi, np.nan],
                         np.nextafter(xp[rng.integers(0, ng, 100)], hi).clip(lo, hi),
                         np.nextafter(xp[rng.integers(0, ng, 100)], lo).clip(lo, hi)])
    tool = tp.Interpolators(make_owner(), {"type": "Interpolators"})

    def same(got, want, what):
        assert np.array_equal(np.isnan(got), np.isnan(want)), what
        ok = ~np.isnan(want)
        assert_bits_equal(got[ok], want[ok], what)

    with np.errstate(all="ignore"):
        want, st = ogather.interp_b(xq, xp, y)
        assert (st == 0).all()
        same(tool.interpolate1D(xp, y)(xq), want, f"B ng={ng}")
        y2 = np.stack([y, -2 * y, y * y])
        want, st = ogather.interp_c(xq, xp, y2)
        same(tool.interpolate1D(xp, y2)(xq), want, f"C ng={ng}")

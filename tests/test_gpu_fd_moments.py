"""GPU parity of the FiniteDifference stencils (bit-for-bit vs the oracle and
the reference-generated goldens) and of the moment reduction (1e-12 relative:
summation order is the only difference; parity unpinned by the reference)."""
import numpy as np
import pytest
import torch

from helpers import C_LIGHT, M_E, assert_bits_equal, beam_ensemble, make_grid, make_owner, thermal_ensemble
from oracle import fd as ofd
from oracle import moments as omom

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tp():
    import turbopy_b200
    return turbopy_b200


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(f"{golden_dir}/fd.npz")


def fd_tool(tp, n, rmax=1.0, method="centered"):
    grid = make_grid(n, 0.0, rmax)
    return tp.FiniteDifference(make_owner(grid=grid), {"type": "FiniteDifference", "method": method}), grid


@pytest.mark.parametrize("tag,n", [("n10", 10), ("n1025", 1025)])
def test_fd_golden(tp, golden, tag, n):
    tool, grid = fd_tool(tp, n, float(golden[f"{tag}_rmax"]))
    y = golden[f"{tag}_y"]
    yd = torch.from_numpy(y).cuda()
    assert_bits_equal(tool.centered_difference(yd).cpu().numpy(), golden[f"{tag}_centered"], "centered")
    assert_bits_equal(tool.upwind_left(yd).cpu().numpy(), golden[f"{tag}_upwind"], "upwind")
    D = tool.del2_radial()
    assert D.shape == (n, n)
    assert_bits_equal((D @ yd).cpu().numpy(), golden[f"{tag}_del2"], "del2_radial @ y")
    assert_bits_equal(D.dot(y), golden[f"{tag}_del2"], "numpy operand")
    # numpy in -> numpy out (how tests/test_computetools.py:122-138 call it)
    out = tool.centered_difference(y)
    assert isinstance(out, np.ndarray)
    assert_bits_equal(out, golden[f"{tag}_centered"])


def test_fd_one_mi_grid_windows(tp, golden):
    n = (1 << 20) + 1
    tool, grid = fd_tool(tp, n)
    y = np.sin(2 * np.pi * grid.r) + 0.1 * grid.r ** 2
    yd = torch.from_numpy(y).cuda()
    sel = golden["n1m_sel"]
    c = tool.centered_difference(yd).cpu().numpy()
    u = tool.upwind_left(yd).cpu().numpy()
    d = (tool.del2_radial() @ yd).cpu().numpy()
    assert_bits_equal(c[sel], golden["n1m_centered"]); assert_bits_equal(u[sel], golden["n1m_upwind"])
    assert_bits_equal(d[sel], golden["n1m_del2"])
    assert_bits_equal(c, ofd.centered_difference(y, grid.dr))
    assert_bits_equal(u, ofd.upwind_left(y, grid.cell_widths))
    assert_bits_equal(d, ofd.del2_radial_apply(y, grid.r, grid.dr))


@pytest.mark.parametrize("n", [2, 3, 4, 5, 7, 8, 9, 511, 512, 513])
def test_fd_small_and_odd_sizes(tp, n):
    tool, grid = fd_tool(tp, n, 3.0)
    y = np.random.default_rng(n).normal(size=n)
    yd = torch.from_numpy(y).cuda()
    assert_bits_equal(tool.centered_difference(yd).cpu().numpy(), ofd.centered_difference(y, grid.dr))
    assert_bits_equal(tool.upwind_left(yd).cpu().numpy(), ofd.upwind_left(y, grid.cell_widths))
    assert_bits_equal((tool.del2_radial() @ yd).cpu().numpy(), ofd.del2_radial_apply(y, grid.r, grid.dr))


def test_fd_setup_ddx_and_fresh_arrays(tp):
    """tests/test_computetools.py:104-119."""
    tool, grid = fd_tool(tp, 10, 10.0)
    y = torch.arange(0, 10, dtype=torch.float64, device="cuda")
    center = tool.setup_ddx()
    assert center == tool.centered_difference
    out = center(y)
    assert out.shape == (10,) and out.data_ptr() != y.data_ptr()
    tool.input_data["method"] = "upwind_left"
    assert tool.setup_ddx() == tool.upwind_left
    tool.input_data["method"] = "nope"
    with pytest.raises(AssertionError):
        tool.setup_ddx()


def test_del2_radial_matrix_entries(tp):
    """tests/test_computetools.py:167-187 compares D.toarray() element-wise;
    here against the oracle's rows (pinned to the reference in make_golden)."""
    tool, grid = fd_tool(tp, 10, 10.0)
    D = tool.del2_radial()
    A = D.toarray()
    below, diag, above = ofd.del2_radial_rows(grid.r, grid.dr)
    idx = np.arange(10)
    expect = np.zeros((10, 10))
    expect[idx, idx] = diag
    expect[idx[1:], idx[:-1]] = below[1:]
    expect[idx[:-1], idx[1:]] = above[:-1]
    assert np.array_equal(A, expect)
    assert list(D.offsets) == [-1, 0, 1]
    assert D.data.shape == (3, 10)


# ---------------------------------------------------------------- moments
@pytest.mark.parametrize("n", [1, 2, 1000, 100_001, 1 << 22])
@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_moments_vs_oracle(tp, n, layout):
    _, mom0 = (thermal_ensemble if n % 2 else beam_ensemble)(n, seed=n)
    mom = tp.to_soa(mom0) if layout == "soa" else torch.from_numpy(mom0).cuda()
    c2 = C_LIGHT ** 2
    got = tp.reduce_moments(mom, M_E, c2).cpu().numpy()
    ref = omom.moments(mom0[:, 0], mom0[:, 1], mom0[:, 2], M_E, c2)
    scale = np.array([ref[0], np.abs(mom0[:, 0]).sum(), np.abs(mom0[:, 1]).sum(), np.abs(mom0[:, 2]).sum(),
                      ref[4], 1.0])
    # tolerance 1e-12 relative to the sum of magnitudes (summation order differs; parity unpinned)
    assert np.all(np.abs(got[:6] - ref[:6]) <= 1e-12 * np.maximum(scale, 1e-300)), (got, ref)
    assert got[5] == n
    again = tp.reduce_moments(mom, M_E, c2).cpu().numpy()
    assert_bits_equal(got, again, "deterministic reduction")


def test_kinetic_energy_diagnostic(tp, tmp_path):
    n = 50_000
    _, mom0 = thermal_ensemble(n, seed=77)
    mom = tp.to_soa(mom0)
    owner = make_owner(num_steps=4)
    d = tp.KineticEnergyDiagnostic(owner, {"momentum": "Particles:momentum", "mass": M_E,
                                            "filename": "ke.csv", "directory": str(tmp_path)})
    d.inspect_resource({"Particles:momentum": mom})
    d.initialize()
    for _ in range(4):
        d.diagnose()
    d.finalize()
    rows = np.genfromtxt(tmp_path / "ke.csv", delimiter=",")
    assert rows.shape == (5, 6)
    ref = omom.moments(mom0[:, 0], mom0[:, 1], mom0[:, 2], M_E, C_LIGHT ** 2)
    assert abs(rows[0, 0] - ref[0]) <= 1e-12 * ref[0]
    assert rows[0, 5] == n


# ---------------------------------------------------------------- remaining banded operators (8f-1)
@pytest.mark.parametrize("name", ofd.BANDED_NAMES)
def test_banded_operator_goldens(tp, golden_dir, name):
    """Every other FiniteDifference matrix builder against the reference-generated
    goldens (stored diagonals, dense form, D @ y) and the oracle."""
    g = np.load(f"{golden_dir}/fd_operators.npz")
    for tag, n, rmax in (("n10", 10, 10.0), ("n1025", 1025, 1.0)):
        tool, grid = fd_tool(tp, n, rmax)
        D = getattr(tool, name)()
        assert D.shape == (n, n)
        assert list(D.offsets) == list(g[f"{tag}_{name}_offsets"])
        want = g[f"{tag}_{name}_data"]
        got = D.data
        assert np.array_equal(np.nan_to_num(got, nan=-7.0), np.nan_to_num(want, nan=-7.0)), "stored diagonals"
        y = g[f"{tag}_y"]
        out = (D @ torch.from_numpy(y).cuda()).cpu().numpy()
        assert_bits_equal(np.nan_to_num(out, nan=-7.0), np.nan_to_num(g[f"{tag}_{name}_apply"], nan=-7.0), "D @ y")
        data, offsets = ofd.banded_operator(name, grid.r, grid.dr)
        assert_bits_equal(np.nan_to_num(out, nan=-7.0), np.nan_to_num(ofd.dia_apply(data, offsets, y), nan=-7.0))
        if n == 10:
            assert np.array_equal(np.nan_to_num(D.toarray(), nan=-7.0),
                                  np.nan_to_num(g[f"{tag}_{name}_dense"], nan=-7.0)), "dense form"


def test_banded_operator_composition_and_large_grid(tp):
    """BC @ D composes lazily; a 16 Mi-point apply equals the oracle on sampled windows."""
    n = (1 << 24) + 1
    tool, grid = fd_tool(tp, n)
    y = np.sin(40.0 * grid.r) + grid.r ** 2
    yd = torch.from_numpy(y).cuda()
    for name in ("ddx", "del2", "BC_right_extrap"):
        out = (getattr(tool, name)() @ yd).cpu().numpy()
        data, offsets = ofd.banded_operator(name, grid.r, grid.dr)
        assert_bits_equal(out, ofd.dia_apply(data, offsets, y), name)
    op = tool.BC_left_extrap() @ tool.ddx()
    got = (op @ yd).cpu().numpy()
    d1, o1 = ofd.banded_operator("ddx", grid.r, grid.dr)
    d2, o2 = ofd.banded_operator("BC_left_extrap", grid.r, grid.dr)
    assert_bits_equal(got, ofd.dia_apply(d2, o2, ofd.dia_apply(d1, o1, y)))


@pytest.mark.parametrize("n", [3, 10, 1000, 300_001])
def test_linspace_grids_rebuild_r_in_registers_with_the_same_bits(tp, n, monkeypatch):
    """On a turboPy grid (r = r_min + (r_max - r_min) * linspace(0, 1, N)) upwind_left, del2_radial and the Poisson
    solve rebuild the node coordinates from three doubles instead of reading r / cell_widths; a grid whose array
    does not match the formula (or TURBOPY_B200_LINSPACE=0) reads the arrays.  Same bits either way."""
    from oracle import poisson as opoisson
    rng = np.random.default_rng(n)
    y = rng.normal(size=n)
    yd = torch.from_numpy(y).cuda()
    results = {}
    for mode in ("lin", "arrays"):
        if mode == "arrays":
            monkeypatch.setenv("TURBOPY_B200_LINSPACE", "0")
        tool, grid = fd_tool(tp, n, 2.5)                       # a fresh Grid object: probed once per grid
        gd = tp.GridOnDevice.of(grid, yd.device)
        assert (gd.linspace is not None) == (mode == "lin")
        ps = tp.PoissonSolver1DRadial(tool.owner, {"type": "PoissonSolver1DRadial"})
        with np.errstate(all="ignore"):
            results[mode] = (tool.upwind_left(yd).cpu().numpy(), (tool.del2_radial() @ yd).cpu().numpy(),
                             ps.solve(yd).cpu().numpy())
            want = (ofd.upwind_left(y, grid.cell_widths), ofd.del2_radial_apply(y, grid.r, grid.dr))
        for got, ref, what in zip(results[mode][:2], want, ("upwind", "del2_radial")):
            assert_bits_equal(np.nan_to_num(got, nan=-7.0), np.nan_to_num(ref, nan=-7.0), f"{what} ({mode})")
        ref = opoisson.solve(y, grid.r, grid.cell_widths)
        assert np.abs(results[mode][2] - ref).max() <= 1e-11 * max(np.abs(ref).max(), 1e-300)
    for a, b in zip(results["lin"], results["arrays"]):
        assert_bits_equal(np.nan_to_num(a, nan=-7.0), np.nan_to_num(b, nan=-7.0), "lin vs arrays")
    # a stretched grid is not a linspace: detected, arrays are read
    tool, grid = fd_tool(tp, max(n, 4), 2.5)
    grid.r = grid.r ** 1.5
    assert tp.GridOnDevice.of(grid, yd.device).linspace is None


@pytest.mark.parametrize("n,rmax", [(10, 1.0), (1025, 2.5), (100_003, 1.0)])
def test_del2_radial_axpy_is_numpys_update(tp, n, rmax, monkeypatch):
    """D.axpy(f, dtau) == f + dtau * (D @ f) with numpy's three roundings per point (the field update
    of configs[4] without an eager axpy), on a linspace grid (r rebuilt in registers) and reading r."""
    grid = make_grid(n, 0.0, rmax)
    rng = np.random.default_rng(n)
    f = rng.normal(size=n)
    dtau = 0.37 * grid.dr ** 2
    want = f + dtau * ofd.del2_radial_apply(f, grid.r, grid.dr)
    fd_t = torch.from_numpy(f).cuda()
    for lin in ("1", "0"):
        monkeypatch.setenv("TURBOPY_B200_LINSPACE", lin)
        tp.GridOnDevice._cache.clear()
        tool = tp.FiniteDifference(make_owner(grid=grid), {"type": "FiniteDifference", "method": "centered"})
        D = tool.del2_radial()
        got = D.axpy(fd_t, dtau)
        assert got.data_ptr() != fd_t.data_ptr()
        assert_bits_equal(got.cpu().numpy(), want, f"linspace={lin}")
        assert_bits_equal(D.axpy(f, dtau), want, "numpy operand")
    assert_bits_equal(tool.ddx().axpy(fd_t, 0.5).cpu().numpy(), f + 0.5 * ofd_ddx(f, grid.dr), "generic operator")


def ofd_ddx(f, dr):
    g = 1 / (2.0 * dr)
    out = np.zeros_like(f)
    out[:-1] = out[:-1] + (0.0 + g) * f[1:]
    out[1:] = (0.0 - g) * f[:-1] + out[1:]
    return out

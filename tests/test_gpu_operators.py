"""the tools registered over the stock names must not break callers of the stock API.
* ``BandedOperator`` supports the sparse-matrix arithmetic the reference itself uses on its dia matrices
  (``D1 + D2``, computetools.py:221) and that user modules use (scalar *, -, .T, tocsr ...): compared with scipy
  on the UNMODIFIED reference's own builders, bit for bit;
* ``interpolate1D`` sorts unsorted ``x`` like ``interp1d(assume_sorted=False)`` and hands spline kinds
  (tests/test_computetools.py:82-86 of the reference) to the stock tool."""
import numpy as np
import pytest
import torch

from helpers import assert_bits_equal
from oracle import refload

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refload.available(), reason="reference neither at /root/reference nor staged")]

BUILDERS = ["ddx", "ddr", "del2", "radial_curl", "del2_radial", "BC_left_extrap", "BC_left_avg", "BC_left_quad",
            "BC_left_flat", "BC_right_extrap"]


@pytest.fixture(scope="module")
def pair():
    turbopy = refload.load()
    import turbopy_b200 as tp
    from turbopy import computetools as stock          # the stock classes, reached past the overridden registry
    sim = refload.make_owner(turbopy, grid={"N": 67, "r_min": 0.0, "r_max": 2.0, "coordinate_system": "cylindrical"})
    ref = stock.FiniteDifference(sim, {"type": "FiniteDifference", "method": "centered"})
    mine = tp.FiniteDifference(sim, {"type": "FiniteDifference", "method": "centered"})
    return turbopy, tp, sim, ref, mine


@pytest.mark.parametrize("a", BUILDERS)
def test_diagonals_and_dense_form_without_forming_the_matrix(pair, a):
    _, _, _, ref, mine = pair
    A, S = getattr(mine, a)(), getattr(ref, a)()
    assert_bits_equal(A.toarray(), S.toarray(), a)
    offsets, data = A.diagonals()
    D = S.todia() if hasattr(S, "todia") else S
    dense = D.toarray()
    n = dense.shape[0]
    for k, off in enumerate(offsets):
        j = np.arange(max(0, off), min(n, n + off))
        # value equality: a stored -0.0 (radial_curl's first sub-diagonal entry at r = 0) reads +0.0 in scipy's dense form
        assert np.array_equal(data[k].cpu().numpy()[j], dense[j - off, j]), f"{a} diagonal {off}"


@pytest.mark.parametrize("a,b", [("ddx", "del2"), ("del2_radial", "ddr"), ("radial_curl", "del2"),
                                 ("BC_left_extrap", "BC_right_extrap"), ("ddr", "ddx")])
def test_operator_arithmetic_matches_scipy(pair, a, b):
    _, _, sim, ref, mine = pair
    A, B = getattr(mine, a)(), getattr(mine, b)()
    SA, SB = getattr(ref, a)(), getattr(ref, b)()
    rng = np.random.default_rng(5)
    f = rng.normal(size=sim.grid.num_points)
    fd = torch.from_numpy(f).cuda()
    cases = {
        "A+B": (A + B, SA + SB), "A-B": (A - B, SA - SB), "2.5*A": (2.5 * A, 2.5 * SA), "A*0.1": (A * 0.1, SA * 0.1),
        "-A": (-A, -SA), "A/3": (A / 3.0, SA / 3.0), "A.T": (A.T, SA.T), "A+scipy": (A + SB, SA + SB),
        "1.5*A-0.5*B": (1.5 * A - 0.5 * B, 1.5 * SA - 0.5 * SB),
    }
    for name, (got, want) in cases.items():
        assert got.shape == want.shape
        assert_bits_equal(got.toarray(), want.toarray(), f"{a},{b}: {name} entries")
        assert_bits_equal((got @ fd).cpu().numpy(), want @ f, f"{a},{b}: ({name}) @ f")
        assert_bits_equal(got @ f, want @ f, f"{a},{b}: ({name}) @ numpy f")
    for conv in ("tocsr", "tocsc", "tocoo", "todia"):
        assert_bits_equal(getattr(A, conv)().toarray(), SA.toarray(), conv)
    F = rng.normal(size=(sim.grid.num_points, 3))
    assert_bits_equal((A @ torch.from_numpy(F).cuda()).cpu().numpy(), SA @ F, "block of columns")
    # the reference's own implicit-solve idiom: spsolve on (I - dt D)
    from scipy import sparse
    from scipy.sparse.linalg import spsolve
    M = (sparse.identity(f.shape[0], format="dia") - 1e-3 * SB)
    mine_M = (-1e-3 * B + sparse.identity(f.shape[0], format="dia"))
    assert_bits_equal(mine_M.toarray(), M.toarray(), "I - dt D")
    assert np.allclose(spsolve(mine_M.tocsc(), f), spsolve(M.tocsc(), f), rtol=1e-12, atol=0)


def test_unsorted_abscissas_are_sorted_like_scipy(pair):
    from scipy.interpolate import interp1d
    _, tp, sim, _, _ = pair
    tool = tp.Interpolators(sim, {"type": "Interpolators"})
    rng = np.random.default_rng(9)
    x = np.sort(rng.uniform(0, 1, 200))
    y = np.sin(7 * x)
    y2 = np.stack([np.cos(3 * x), x ** 2])
    perm = rng.permutation(200)
    xq = rng.uniform(x[0], x[-1], 5000)
    assert_bits_equal(tool.interpolate1D(x[perm], y[perm])(xq), interp1d(x[perm], y[perm])(xq), "1-D y")
    assert_bits_equal(tool.interpolate1D(x[perm], y2[:, perm])(xq), interp1d(x[perm], y2[:, perm])(xq), "2-D y")
    xd, yd = torch.from_numpy(x[perm]).cuda(), torch.from_numpy(y[perm]).cuda()
    got = tool.interpolate1D(xd, yd)(torch.from_numpy(xq).cuda())
    assert_bits_equal(got.cpu().numpy(), interp1d(x, y)(xq), "device arrays")


@pytest.mark.parametrize("kind", ["quadratic", "cubic"])
def test_spline_kinds_go_to_the_stock_tool(pair, kind):
    """The reference's own test (tests/test_computetools.py:71-86) calls interpolate1D(x, y, 'quadratic')."""
    from scipy import interpolate
    _, tp, sim, _, _ = pair
    assert "Interpolators" in tp.stock_tools
    tool = tp.Interpolators(sim, {"type": "Interpolators"})
    x = np.arange(0, 10)
    y = np.exp(-x / 3.0)
    xnew = np.arange(0, 9, 0.1)
    f = tool.interpolate1D(x, y, kind)
    assert np.allclose(f(xnew), interpolate.interp1d(x, y, kind=kind)(xnew))
    with pytest.raises(NotImplementedError):
        tool.interpolate1D(torch.from_numpy(x * 1.0).cuda(), torch.from_numpy(y).cuda(), kind)

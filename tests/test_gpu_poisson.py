"""GPU parity of PoissonSolver1DRadial.solve.  The running sums
are parallel scans: agreement with numpy's sequential cumsum is to rounding.
Tolerance: 1e-12 of max|phi| (stated here; the reference has no test of its own)."""
import numpy as np
import pytest
import torch

from helpers import make_grid, make_owner
from oracle import poisson as opoisson

pytestmark = pytest.mark.gpu

TOL = 1e-12


def source(r):
    return np.exp(-((r - 0.7) / 0.2) ** 2) - 0.3 * np.cos(5 * r)


@pytest.mark.parametrize("tag,n", [("n64", 64), ("n4097", 4097), ("n1m", (1 << 20) + 1)])
def test_poisson_golden(golden_dir, tag, n):
    import turbopy_b200 as tp
    g = np.load(f"{golden_dir}/poisson.npz")
    grid = make_grid(n, 0.0, 2.0)
    tool = tp.PoissonSolver1DRadial(make_owner(grid=grid), {"type": "PoissonSolver1DRadial"})
    src = source(grid.r)
    phi = tool.solve(torch.from_numpy(src).cuda()).cpu().numpy()
    ref = opoisson.solve(src, grid.r, grid.cell_widths)
    scale = np.abs(ref).max()
    assert np.abs(phi - ref).max() <= TOL * scale
    if n <= 4097:
        assert np.array_equal(src, g[f"{tag}_src"])
        assert np.abs(phi - g[f"{tag}_phi"]).max() <= TOL * scale
    else:
        assert np.abs(phi[g[f"{tag}_sel"]] - g[f"{tag}_phi"]).max() <= TOL * float(g[f"{tag}_scale"])
    assert phi[-1] == 0.0                                   # I2 - I2[-1]
    again = tool.solve(torch.from_numpy(src).cuda()).cpu().numpy()
    assert np.array_equal(phi, again)                       # deterministic
    out_np = tool.solve(src)                                # numpy in -> numpy out
    assert isinstance(out_np, np.ndarray) and np.array_equal(out_np, phi)


@pytest.mark.parametrize("n", [3, 5, 2047, 2048, 2049, 300_001])
def test_poisson_ragged_sizes(n):
    import turbopy_b200 as tp
    grid = make_grid(n, 0.0, 1.0)
    tool = tp.PoissonSolver1DRadial(make_owner(grid=grid), {"type": "PoissonSolver1DRadial"})
    src = np.random.default_rng(n).normal(size=n)
    phi = tool.solve(torch.from_numpy(src).cuda()).cpu().numpy()
    ref = opoisson.solve(src, grid.r, grid.cell_widths)
    assert np.abs(phi - ref).max() <= TOL * max(np.abs(ref).max(), 1e-300) * 10

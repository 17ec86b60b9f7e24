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


@pytest.mark.parametrize("n", [3, 5, 2047, 2048, 2049, 16_383, 16_384, 16_385, 300_001])
def test_poisson_ragged_sizes(n):
    import turbopy_b200 as tp
    grid = make_grid(n, 0.0, 1.0)
    tool = tp.PoissonSolver1DRadial(make_owner(grid=grid), {"type": "PoissonSolver1DRadial"})
    src = np.random.default_rng(n).normal(size=n)
    phi = tool.solve(torch.from_numpy(src).cuda()).cpu().numpy()
    ref = opoisson.solve(src, grid.r, grid.cell_widths)
    assert np.abs(phi - ref).max() <= TOL * max(np.abs(ref).max(), 1e-300) * 10
    assert phi[-1] == 0.0                                   # `last` is formed by the instructions that form out[n-1]


@pytest.mark.parametrize("n", [1000, 16_384, 100_003])
@pytest.mark.parametrize("shift", [1, 2, 3])
def test_poisson_unaligned_pointers_match_aligned(n, shift):
    """The 256-bit path (32-byte aligned arrays) and the scalar path (any 8-byte alignment) and the single-launch
    small-grid path all run the same arithmetic: identical bits.  Straight through the C ABI."""
    from turbopy_b200 import _lib
    L = _lib.lib()
    grid = make_grid(n, 0.0, 1.5)
    src = np.random.default_rng(7 * n + shift).normal(size=n)
    dr = float(np.mean(grid.cell_widths))
    scratch = torch.empty(int(L.tp_poisson_scratch_doubles(n)), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def run(off):
        buf = torch.zeros((3, n + 8), dtype=torch.float64, device="cuda")
        s_, r_, o_ = buf[0, off:off + n], buf[1, off:off + n], buf[2, off:off + n]
        s_.copy_(torch.from_numpy(src)); r_.copy_(torch.from_numpy(grid.r))
        _lib.check(L.tp_poisson_radial_solve_f64(s_.data_ptr(), r_.data_ptr(), dr, o_.data_ptr(), n, scratch.data_ptr(), stream))
        torch.cuda.synchronize()
        assert float(buf[2, :off].abs().sum()) == 0.0 and float(buf[2, off + n:].abs().sum()) == 0.0   # no stray writes
        return o_.cpu().numpy()

    aligned, shifted = run(0), run(shift)
    assert np.array_equal(aligned, shifted)
    assert aligned[-1] == 0.0
    ref = opoisson.solve(src, grid.r, grid.cell_widths)
    assert np.abs(aligned - ref).max() <= TOL * np.abs(ref).max() * 10


@pytest.mark.parametrize("n,stretched", [(100_003, False), (100_003, True), (1_000_000, False), (16_385, True)])
def test_poisson_interior_tiles_match_the_generic_tile_path(n, stretched, monkeypatch):
    """Tiles strictly inside (0, n-1) run without the per-element index tests (csrc/scan_kernels.cu: IN); the
    arithmetic is the same, so the bits are: TURBOPY_B200_SCAN_INTERIOR=0 sends every tile down the generic path."""
    import turbopy_b200 as tp
    grid = make_grid(n, 0.0, 1.5)
    if stretched:
        grid.r = grid.r ** 1.25
        grid.cell_widths = grid.r[1:] - grid.r[:-1]
    src = torch.from_numpy(np.random.default_rng(n).normal(size=n)).cuda()
    tool = tp.PoissonSolver1DRadial(make_owner(grid=grid), {"type": "PoissonSolver1DRadial"})
    fast = tool.solve(src).cpu().numpy()
    monkeypatch.setenv("TURBOPY_B200_SCAN_INTERIOR", "0")
    generic = tool.solve(src).cpu().numpy()
    assert np.array_equal(fast.view(np.uint64), generic.view(np.uint64))
    ref = opoisson.solve(src.cpu().numpy(), grid.r, grid.cell_widths)
    assert np.abs(fast - ref).max() <= TOL * np.abs(ref).max() * 10

#!/usr/bin/env python
"""Tuning sweep for the linspace-grid kernels (upwind_left, del2_radial, Poisson) at a given CTAs-per-SM."""
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turbopy_b200 as tp  # noqa: E402
from oracle import grid as ogrid  # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = []
for n in ((1 << 24) + 1, 1 << 28):
    r = ogrid.grid_points(0.0, 1.0, n)
    grid = SimpleNamespace(r=r, dr=ogrid.grid_dr(0.0, 1.0, n), num_points=n, r_min=0.0, r_max=1.0, cell_widths=ogrid.cell_widths(r))
    ow = SimpleNamespace(clock=SimpleNamespace(dt=1e-12, num_steps=1), grid=grid)
    fd = tp.FiniteDifference(ow, {"type": "FiniteDifference", "method": "centered"})
    ps = tp.PoissonSolver1DRadial(ow, {"type": "PoissonSolver1DRadial"})
    y = torch.rand(n, dtype=torch.float64, device="cuda")
    D = fd.del2_radial()
    assert tp.GridOnDevice.of(grid, y.device).linspace is not None or os.environ.get("TURBOPY_B200_LINSPACE") == "0"
    out.append(f"upwind@{n >> 20}Mi {timed(lambda: fd.upwind_left(y)):.4f}ms  del2r {timed(lambda: D @ y):.4f}ms  poisson {timed(lambda: ps.solve(y)):.4f}ms")
    del y, fd, ps, D
print(f"fd_blocks={os.environ.get('TURBOPY_B200_FD_BLOCKS', 'def')} scan={os.environ.get('TURBOPY_B200_SCAN_BLOCKS', 'occ')}: " + " | ".join(out))

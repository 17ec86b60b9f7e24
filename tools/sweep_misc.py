#!/usr/bin/env python
"""Tuning sweep: moments reduction and Poisson solve at the CTAs-per-SM given by
TURBOPY_B200_RED_BLOCKS / TURBOPY_B200_SCAN_BLOCKS."""
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
    return e0.elapsed_time(e1) * 1e-3 / iters


out = []
for n in (1 << 24, 1 << 27):
    mom = tp.soa_particles(n, fill=1e-22)
    o8 = torch.empty(8, dtype=torch.float64, device="cuda")
    sc = torch.empty(int(tp._lib.lib().tp_reduce_scratch_doubles()), dtype=torch.float64, device="cuda")
    t = timed(lambda: tp.reduce_moments(mom, 9.1094e-31, 2.9979e8 ** 2, out=o8, scratch=sc))
    out.append(f"moments@{n >> 20}Mi {24 * n / t / 1e9 / 6545.3 * 100:.1f}")
    del mom
for n in ((1 << 24) + 1, 1 << 28):
    r = ogrid.grid_points(0.0, 1.0, n)
    grid = SimpleNamespace(r=r, dr=ogrid.grid_dr(0.0, 1.0, n), num_points=n, r_min=0.0, r_max=1.0, cell_widths=ogrid.cell_widths(r))
    ps = tp.PoissonSolver1DRadial(SimpleNamespace(clock=None, grid=grid), {"type": "PoissonSolver1DRadial"})
    y = torch.rand(n, dtype=torch.float64, device="cuda")
    t = timed(lambda: ps.solve(y))
    out.append(f"poisson@{n >> 20}Mi {56 * n / t / 1e9 / 6545.3 * 100:.1f}")
    del y, ps
print(f"red={os.environ.get('TURBOPY_B200_RED_BLOCKS', 'occ')} scan={os.environ.get('TURBOPY_B200_SCAN_BLOCKS', 'occ')}: " + "  ".join(out))

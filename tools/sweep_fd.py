#!/usr/bin/env python
"""Tuning sweep: FD kernels at a given blocks-per-SM (TURBOPY_B200_FD_BLOCKS), two grid sizes."""
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
for n in ((1 << 24) + 1, 1 << 28):
    r = ogrid.grid_points(0.0, 1.0, n)
    grid = SimpleNamespace(r=r, dr=ogrid.grid_dr(0.0, 1.0, n), num_points=n, r_min=0.0, r_max=1.0, cell_widths=ogrid.cell_widths(r))
    ow = SimpleNamespace(clock=SimpleNamespace(dt=1e-12, num_steps=1), grid=grid)
    fd = tp.FiniteDifference(ow, {"type": "FiniteDifference", "method": "centered"})
    y = torch.rand(n, dtype=torch.float64, device="cuda")
    ops = {"centered": (fd.centered_difference, 16), "upwind": (fd.upwind_left, 24), "del2_radial": (fd.del2_radial().dot, 24),
           "ddx": (fd.ddx().dot, 16), "del2": (fd.del2().dot, 16), "BC_left": (fd.BC_left_extrap().dot, 16),
           "radial_curl": (fd.radial_curl().dot, 40)}
    for name, (f, b) in ops.items():
        t = timed(lambda: f(y))
        out.append(f"{name}@{n >> 20}Mi {b * n / t / 1e9 / 6545.3 * 100:.1f}")
    del y, fd, ops
print(f"blocks/SM={os.environ.get('TURBOPY_B200_FD_BLOCKS', 'occ')}: " + "  ".join(out))

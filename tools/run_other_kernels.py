#!/usr/bin/env python
"""Profiling driver: a few launches of every kernel other than the fused
gather+push (which bench.py drives), at sizes larger than L2."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import turbopy_b200 as tp  # noqa: E402
from oracle import grid as ogrid  # noqa: E402

Q_E, M_E = -1.6022e-19, 9.1094e-31


def owner_for(ng, dt=1e-12):
    r = ogrid.grid_points(0.0, 1.0, ng)
    grid = SimpleNamespace(r=r, dr=ogrid.grid_dr(0.0, 1.0, ng), num_points=ng, r_min=0.0, r_max=1.0,
                           cell_widths=ogrid.cell_widths(r))
    return SimpleNamespace(clock=SimpleNamespace(dt=dt, num_steps=10), grid=grid)


n = 16 << 20
reps = 3
tool = tp.BorisPush(owner_for(10), {"type": "BorisPush"})
pos, mom = tp.soa_particles(n), tp.soa_particles(n)
pos.copy_(torch.rand((n, 3), dtype=torch.float64, device="cuda"))
mom.copy_(torch.randn((n, 3), dtype=torch.float64, device="cuda") * (M_E * 2.9979e8 * 0.1))
for _ in range(reps):
    tool.push(pos, mom, Q_E, M_E, np.array([0.0, 1e5, 0.0]), np.array([0.0, 0.0, 1.0]))      # push_tma_kernel
Ep, Bp = tp.soa_particles(n, fill=1e4), tp.soa_particles(n, fill=0.5)
for _ in range(reps):
    tool.push(pos, mom, Q_E, M_E, Ep, Bp)                                                    # push_fields_tma_kernel
del Ep, Bp
out8 = None
for _ in range(reps):
    out8 = tp.reduce_moments(mom, M_E, 2.9979e8 ** 2)                                          # moments_partial_kernel
ng = (1 << 26) + 1
ow = owner_for(ng)
fdt = tp.FiniteDifference(ow, {"type": "FiniteDifference", "method": "centered"})
y = torch.rand(ng, dtype=torch.float64, device="cuda")
D, Dx = fdt.del2_radial(), fdt.ddx()
for _ in range(reps):
    fdt.centered_difference(y); fdt.upwind_left(y); D @ y; Dx @ y                              # fd_* kernels
ps = tp.PoissonSolver1DRadial(ow, {"type": "PoissonSolver1DRadial"})
for _ in range(reps):
    ps.solve(y)                                                                                # poisson_* kernels
ow2 = owner_for(4097)
f = tp.Interpolators(ow2, {"type": "Interpolators"}).interpolate1D(torch.from_numpy(ow2.grid.r).cuda(),
                                                                     torch.cos(torch.from_numpy(ow2.grid.r).cuda()))
f.deferred_check = True
xq = 0.05 + 0.9 * torch.rand(n, dtype=torch.float64, device="cuda")
for _ in range(reps):
    f(xq)                                                                                      # interp_kernel
torch.cuda.synchronize()
f.check()
print("ok", float(out8[0]))

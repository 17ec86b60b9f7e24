#!/usr/bin/env python
"""Tuning driver: gather_push_steps (static grid field, K steps per pass) on the pif-shaped grid."""
import os
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turbopy_b200 as tp  # noqa: E402
from oracle import grid as ogrid  # noqa: E402

n, ng = 16 << 20, 4097
r = ogrid.grid_points(0.0, 1.0, ng)
owner = SimpleNamespace(clock=SimpleNamespace(dt=1e-13), grid=SimpleNamespace(r=r, dr=ogrid.grid_dr(0.0, 1.0, ng), num_points=ng))
tool = tp.BorisPush(owner, {"type": "BorisPush"})
rd = torch.from_numpy(r).cuda()
E = (None, torch.cos(6.28 * rd), None)
pos, mom = tp.soa_particles(n), tp.soa_particles(n, fill=0.0)
pos.copy_(0.05 + 0.9 * torch.rand((n, 3), dtype=torch.float64, device="cuda"))
for K in (1, 2, 4, 16):
    for _ in range(3):
        tool.gather_push_steps(pos, mom, -1.6022e-19, 9.1094e-31, K, E_grid=E)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tool.gather_push_steps(pos, mom, -1.6022e-19, 9.1094e-31, K, E_grid=E)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"gather_push_steps K={K}: {ms:.4f} ms per pass, {n * K / ms / 1e6:.1f} G pushes/s")
tool.check_status()

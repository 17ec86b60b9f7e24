#!/usr/bin/env python
"""End-to-end host path (numpy arrays in, same arrays updated) with pageable vs pinned host memory."""
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turbopy_b200 as tp  # noqa: E402
from oracle import grid as ogrid  # noqa: E402

n, ng = 16 << 20, 4097
r = ogrid.grid_points(0.0, 1.0, ng)
owner = SimpleNamespace(clock=SimpleNamespace(dt=1e-13), grid=SimpleNamespace(r=r, dr=ogrid.grid_dr(0.0, 1.0, ng), num_points=ng))
tool = tp.BorisPush(owner, {"type": "BorisPush"})
Ey = np.cos(6.28 * r)
rng = np.random.default_rng(0)
for kind in ("pageable", "pinned"):
    if kind == "pinned":
        pos = torch.empty((n, 3), dtype=torch.float64).pin_memory().numpy()
        mom = torch.empty((n, 3), dtype=torch.float64).pin_memory().numpy()
    else:
        pos, mom = np.empty((n, 3)), np.empty((n, 3))
    pos[:] = rng.uniform(0.05, 0.95, (n, 3)); mom[:] = 0.0
    t0 = time.perf_counter()
    tool.gather_push(pos, mom, -1.6022e-19, 9.1094e-31, E_grid=(None, Ey, None))
    first = time.perf_counter() - t0
    tool.gather_push(pos, mom, -1.6022e-19, 9.1094e-31, E_grid=(None, Ey, None))
    from turbopy_b200 import tools as _t
    print(f"{kind}: first call {first * 1e3:.1f} ms, pinned by the library: {[ok for _, ok in _t._pinned.values()]}")
    t0 = time.perf_counter()
    steps = 8
    for _ in range(steps):
        tool.gather_push(pos, mom, -1.6022e-19, 9.1094e-31, E_grid=(None, Ey, None))
    dt = (time.perf_counter() - t0) / steps
    print(f"{kind}: {dt * 1e3:.2f} ms per step, {n / dt / 1e9:.3f} G pushes/s, {48 * n / dt / 1e9:.1f} GB/s each way")

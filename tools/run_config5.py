#!/usr/bin/env python
"""BASELINE configs[4] at full size: centered / upwind ddx and the radial del2 on a 256 Mi-point cylindrical grid
driving a field update f <- f + dt * (del2_radial @ f) for 100 iterations: one iteration = three own kernels
(centered, upwind, the fused del2_radial axpy of tp_fd_del2_radial_axpy_f64); the round-1 variant with a torch axpy
is timed beside it."""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turbopy_b200 as tp  # noqa: E402
from oracle import grid as ogrid  # noqa: E402

n = 256 << 20
r = ogrid.grid_points(0.0, 1.0, n)
grid = SimpleNamespace(r=r, dr=ogrid.grid_dr(0.0, 1.0, n), num_points=n, r_min=0.0, r_max=1.0, cell_widths=ogrid.cell_widths(r))
fd = tp.FiniteDifference(SimpleNamespace(clock=None, grid=grid), {"type": "FiniteDifference", "method": "centered"})
rd = torch.from_numpy(r).cuda()
f = torch.sin(2 * np.pi * rd) + 0.1 * rd * rd
del rd
D = fd.del2_radial()
dtau = 1e-20                                                  # keeps the explicit update stable on this grid
lin = tp.GridOnDevice.of(grid, f.device).linspace is not None


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def iteration_torch_axpy():
    fd.centered_difference(f)
    fd.upwind_left(f)
    f.add_(D @ f, alpha=dtau)


state = {"f": f}


def iteration():
    fd.centered_difference(state["f"])
    fd.upwind_left(state["f"])
    state["f"] = D.axpy(state["f"], dtau)                        # fresh array, like numpy's f + dtau * (D @ f)


out = {"config": "configs[4]: FD centered / upwind ddx + radial del2 on a 256 Mi-point grid, 100 update iterations",
       "points": n, "linspace_grid_detected": lin,
       "centered_ms": timed(lambda: fd.centered_difference(f), 10), "upwind_ms": timed(lambda: fd.upwind_left(f), 10),
       "del2_radial_ms": timed(lambda: D @ f, 10), "del2_radial_axpy_ms": timed(lambda: D.axpy(f, dtau), 10),
       "iteration_torch_axpy_ms": timed(iteration_torch_axpy, 50), "iteration_ms": timed(iteration, 100)}
f = state["f"]
out["points_per_s_per_operator"] = {k[:-3]: n / out[k] * 1e3 for k in ("centered_ms", "upwind_ms", "del2_radial_ms")}
out["finite"] = bool(torch.isfinite(f[1:]).all())
print(json.dumps(out))

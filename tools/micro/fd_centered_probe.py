#!/usr/bin/env python
"""Why does bench_kernels.py time fd centered_difference at 0.90 ms per 256 Mi points when tools/run_config5.py times
the same kernel at 0.66 ms?  Data (smooth sine vs random), order (first kernel after the allocation), output
allocation (torch.empty_like per call vs the C entry point on a preallocated output)."""
import ctypes as C
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import turbopy_b200 as tp  # noqa: E402
from turbopy_b200 import _lib  # noqa: E402
from oracle import grid as ogrid  # noqa: E402


def timed(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


n = 1 << 28
r = ogrid.grid_points(0.0, 1.0, n)
grid = SimpleNamespace(r=r, dr=ogrid.grid_dr(0.0, 1.0, n), num_points=n, r_min=0.0, r_max=1.0, cell_widths=ogrid.cell_widths(r))
fdt = tp.FiniteDifference(SimpleNamespace(clock=None, grid=grid), {"type": "FiniteDifference", "method": "centered"})
L = _lib.lib()
stream = torch.cuda.current_stream().cuda_stream
for label, y in (("sine", torch.sin(torch.linspace(0, 6.28, n, dtype=torch.float64, device="cuda"))),
                 ("random", torch.rand(n, dtype=torch.float64, device="cuda")),
                 ("sine*1e5", 1e5 * torch.sin(torch.linspace(0, 6.28, n, dtype=torch.float64, device="cuda")))):
    out = torch.empty_like(y)
    t_tool = timed(lambda: fdt.centered_difference(y))
    t_capi = timed(lambda: L.tp_fd_centered_f64(C.c_void_p(y.data_ptr()), C.c_void_p(out.data_ptr()), n, 2 * fdt.dr, stream))
    D = fdt.ddx()
    t_ddx = timed(lambda: D @ y)
    t_tool2 = timed(lambda: fdt.centered_difference(y))
    print(f"{label:9s} centered via tool {t_tool:.4f} ms | C entry point, preallocated out {t_capi:.4f} ms | ddx() @ y {t_ddx:.4f} ms | "
          f"tool again {t_tool2:.4f} ms", flush=True)
    del y, out
    torch.cuda.empty_cache()

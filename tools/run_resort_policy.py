#!/usr/bin/env python
"""config-3 shape (125 M particles, 2^20+1-node grid, six components) with BorisPush's sort_every policy: one
global sort, then a block-local re-sort every K steps.  Prints ms per step INCLUDING the re-sorts, block by block."""
import json
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import turbopy_b200 as tp  # noqa: E402
from tools.run_gather6 import C, M_E, Q_E, ogrid, peak  # noqa: E402

ng, n = (1 << 20) + 1, int(os.environ.get("N", 125_000_000))
r = ogrid.grid_points(0.0, 1.0, ng)
dr = ogrid.grid_dr(0.0, 1.0, ng)
owner = SimpleNamespace(clock=SimpleNamespace(dt=0.5 * dr / C),
                        grid=SimpleNamespace(r=r, dr=dr, num_points=ng, r_min=0.0, r_max=1.0))
rd = torch.from_numpy(r).cuda()
E = torch.stack([1e3 * torch.cos(6.28 * (k + 1) * rd) for k in range(3)], dim=1).contiguous()
B = torch.stack([torch.sin(6.28 * (k + 1) * rd) for k in range(3)], dim=1).contiguous()
rows = []
for every in [int(v) for v in os.environ.get("EVERY", "25,50,100").split(",")]:
    tool = tp.BorisPush(owner, {"type": "BorisPush", "sort_every": every})
    g = torch.Generator(device="cuda").manual_seed(7)
    pos, mom = tp.soa_particles(n), tp.soa_particles(n)
    pos.copy_(0.05 + 0.9 * torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=g))
    mom.copy_(torch.randn((n, 3), dtype=torch.float64, device="cuda", generator=g) * (M_E * C * 0.1))
    tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)          # call 0: the global sort
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for blk in range(int(os.environ.get("BLOCKS", 5))):
        e0.record()
        for _ in range(100):
            tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 100
        w = [w[0] for w in tool._windows.values() if w[0] is not None][0].view(-1, 2)[: n // 960, 1].float()
        rows.append({"sort_every": every, "steps": (blk + 1) * 100, "ms_per_step_incl_resorts": ms,
                     "frac_of_hbm_peak": 96.0 * n / (ms * 1e-3) / 1e9 / peak(),
                     "cells_per_tile_mean_max": [float(w.mean()), float(w.max())]})
        print(json.dumps(rows[-1]), flush=True)
    tool.check_status()
    e0.record()
    tp.resort_local(pos, mom, owner.grid)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"resort_local_ms": e0.elapsed_time(e1), "particles": n}), flush=True)
    del pos, mom
    torch.cuda.empty_cache()

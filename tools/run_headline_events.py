#!/usr/bin/env python
"""Per-launch CUDA-event timings of the headline kernel (configs[1] shape), back to back, optionally with an idle gap
or an L2 flush between launches -- to see how consecutive launches interact.  TURBOPY_B200_MASK_KERNELS=0/1 selects
the generic / mask-specialised kernel."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import turbopy_b200 as tp  # noqa: E402

n, ng = 16 << 20, 4097
r, dr = bench.grid_points(ng)
pusher = tp.BorisPush(bench.Owner(r, dr, bench.DT_PIF), {"type": "BorisPush"})
dev = torch.device("cuda", 0)
fields = torch.zeros((64, 4098), dtype=torch.float64, device=dev)
fields[:, :ng] = torch.from_numpy(bench.wave_fields(r, 64)).to(dev)
pos_h, mom_h = bench.host_ensemble(n, 2345)
pos, mom = tp.to_soa(pos_h, dev), tp.to_soa(mom_h, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def step(k):
    pusher.gather_push(pos, mom, bench.Q_E, bench.M_E, E_grid=(None, fields[k % 64, :ng], None), B_grid=None)


for mode in ("back_to_back", "sleep_200us", "l2_flush"):
    for k in range(5):
        step(k)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * 24)]
    for k in range(24):
        if mode == "sleep_200us":
            torch.cuda._sleep(400_000)
        elif mode == "l2_flush":
            flush.zero_()
        ev[2 * k].record()
        step(k)
        ev[2 * k + 1].record()
    torch.cuda.synchronize()
    t = np.array([ev[2 * k].elapsed_time(ev[2 * k + 1]) * 1e3 for k in range(24)])
    print(f"mask={os.environ.get('TURBOPY_B200_MASK_KERNELS', '1')} {mode:13s} per-launch us: median {np.median(t):.1f} min {t.min():.1f} max {t.max():.1f}  first 6: {np.round(t[:6], 1)}")

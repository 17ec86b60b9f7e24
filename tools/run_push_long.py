#!/usr/bin/env python
"""BorisPush.push with uniform E, B (configs[0] / configs[3] shape) over many steps: throughput and NVML clocks."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import turbopy_b200 as tp  # noqa: E402

n = int(os.environ.get("N", 16 << 20))
r, dr = bench.grid_points(30)
tool = tp.BorisPush(bench.Owner(r, dr, 1e-12), {"type": "BorisPush"})
g = torch.Generator(device="cuda").manual_seed(1)
pos, mom = tp.soa_particles(n), tp.soa_particles(n)
pos.copy_(torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=g))
mom.copy_(torch.randn((n, 3), dtype=torch.float64, device="cuda", generator=g) * (bench.M_E * bench.C_LIGHT * 0.1))
E, B = np.array([0.0, 1e5, 0.0]), np.array([0.0, 0.0, 1.0])
for K in [int(k) for k in os.environ.get("KS", "400,8000").split(",")]:
    for _ in range(5):
        tool.push(pos, mom, bench.Q_E, bench.M_E, E, B)
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(0)
    sampler.start()
    time.sleep(0.02)
    first = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        tool.push(pos, mom, bench.Q_E, bench.M_E, E, B)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop(first, sampler.mark())
    ms = e0.elapsed_time(e1) / K
    print(f"push uniform n={n} K={K} delay={os.environ.get('TURBOPY_B200_PUSH_DELAY_NS', '0')}: {ms*1e3:.1f} us/step "
          f"{n/ms/1e6:.2f} G/s {96*n/(ms*1e-3)/1e9/6545.3:.4f} of peak  clocks {clocks['sm_mhz']} {clocks['reasons']}", flush=True)

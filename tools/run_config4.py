#!/usr/bin/env python
"""BASELINE configs[3] at full size: 256 Mi-particle relativistic beam (gamma log-uniform in [1, 1e3]) in a strong
uniform B, E = 0 -- per-step push() and the temporally blocked push_steps(), pushes/s from CUDA events."""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turbopy_b200 as tp  # noqa: E402

Q_E, M_E, C = -1.6022e-19, 9.1094e-31, 2.9979e8
n, B0 = 256 << 20, 10.0
dt = 0.02 * (2 * np.pi * M_E / (abs(Q_E) * B0))
gen = torch.Generator(device="cuda").manual_seed(4567)
pos, mom = tp.soa_particles(n, fill=0.0), tp.soa_particles(n, fill=0.0)
gam = torch.exp(torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * np.log(1e3))
mom[:, 0].copy_(torch.sqrt(gam * gam - 1.0) * (M_E * C))
del gam
for c in (1, 2):
    mom[:, c].copy_(mom[:, 0] * 0.01 * torch.randn(n, dtype=torch.float64, device="cuda", generator=gen))
tool = tp.BorisPush(SimpleNamespace(clock=SimpleNamespace(dt=dt), grid=None), {"type": "BorisPush"})
E, B = np.zeros(3), np.array([0.0, 0.0, B0])
p2_before = float((mom.double() ** 2).sum())


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


out = {"config": "configs[3]: 256 Mi-particle beam, gamma in [1, 1e3], B = 10 T, E = 0", "particles": n, "dt": dt}
ms = timed(lambda: tool.push(pos, mom, Q_E, M_E, E, B), 20)
out["push_per_step"] = {"ms": ms, "pushes_per_s": n / ms * 1e3, "frac_of_hbm_peak": 96 * n / ms / 1e6 / 6545.3}
for K in (10, 100):
    ms = timed(lambda: tool.push_steps(pos, mom, Q_E, M_E, E, B, K), 3)
    out[f"push_steps_{K}"] = {"ms_per_pass": ms, "pushes_per_s": n * K / ms * 1e3}
p2_after = float((mom.double() ** 2).sum())
out["relative_drift_of_sum_p2"] = abs(p2_after - p2_before) / p2_before      # E = 0: a pure rotation
print(json.dumps(out))

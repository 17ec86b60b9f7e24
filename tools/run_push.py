#!/usr/bin/env python
"""Profiling / tuning driver: BorisPush.push with uniform E, B on n particles (SoA or C-order)."""
import argparse
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import turbopy_b200 as tp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16 << 20)
ap.add_argument("--layout", default="soa")
ap.add_argument("--steps", type=int, default=200)
args = ap.parse_args()
Q_E, M_E = -1.6022e-19, 9.1094e-31
n = args.n
tool = tp.BorisPush(SimpleNamespace(clock=SimpleNamespace(dt=1e-12), grid=None), {"type": "BorisPush"})
if args.layout == "soa":
    pos, mom = tp.soa_particles(n), tp.soa_particles(n)
else:
    pos = torch.empty((n, 3), dtype=torch.float64, device="cuda"); mom = torch.empty_like(pos)
pos.copy_(torch.rand((n, 3), dtype=torch.float64, device="cuda"))
mom.copy_(torch.randn((n, 3), dtype=torch.float64, device="cuda") * (M_E * 2.9979e8 * 0.1))
E, B = np.array([0.0, 1e5, 0.0]), np.array([0.0, 0.0, 1.0])
for _ in range(5):
    tool.push(pos, mom, Q_E, M_E, E, B)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    tool.push(pos, mom, Q_E, M_E, E, B)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print(f"n={n} {args.layout} small={os.environ.get('TURBOPY_B200_SMALLTILE', '-')}: {ms:.4f} ms  {n / ms / 1e6:.2f} G/s  "
      f"{96 * n / ms / 1e6 / 6545.3 * 100:.1f}%")

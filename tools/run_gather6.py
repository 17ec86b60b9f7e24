#!/usr/bin/env python
"""Small driver for profiling: N steps of the 6-component fused gather+push
(ng = 1025 staged in shared memory, or a larger grid through L2)."""
import argparse
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import turbopy_b200 as tp  # noqa: E402
from oracle import grid as ogrid  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ng", type=int, default=1025)
ap.add_argument("--n", type=int, default=16 << 20)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--formula", default="interp")
args = ap.parse_args()
r = ogrid.grid_points(0.0, 1.0, args.ng)
owner = SimpleNamespace(clock=SimpleNamespace(dt=1e-13),
                        grid=SimpleNamespace(r=r, dr=ogrid.grid_dr(0.0, 1.0, args.ng), num_points=args.ng, r_min=0.0, r_max=1.0))
tool = tp.BorisPush(owner, {"type": "BorisPush"})
rd = torch.from_numpy(r).cuda()
E = torch.stack([1e3 * torch.cos(6.28 * (k + 1) * rd) for k in range(3)], dim=1).contiguous()
B = torch.stack([torch.sin(6.28 * (k + 1) * rd) for k in range(3)], dim=1).contiguous()
pos, mom = tp.soa_particles(args.n), tp.soa_particles(args.n, fill=0.0)
pos.copy_(0.05 + 0.9 * torch.rand((args.n, 3), dtype=torch.float64, device="cuda"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    tool.gather_push(pos, mom, -1.6022e-19, 9.1094e-31, E_grid=E, B_grid=B, formula=args.formula)
e0.record()
for _ in range(args.steps):
    tool.gather_push(pos, mom, -1.6022e-19, 9.1094e-31, E_grid=E, B_grid=B, formula=args.formula)
e1.record()
torch.cuda.synchronize()
tool.check_status()
ms = e0.elapsed_time(e1) / args.steps
print(f"ng={args.ng} 6 comps {args.formula}: {ms:.4f} ms/step  {args.n / ms / 1e6:.2f} G pushes/s")

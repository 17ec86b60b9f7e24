#!/usr/bin/env python
"""PoissonSolver1DRadial.solve: CUDA-event timings on linspace and stretched (array-reading) grids, plus the error
against the sequential oracle on a size the oracle finishes in seconds.  `--reps 2` is the ncu launch-list driver.
Usage: python tools/run_poisson.py [--reps 20] [--out gpurun_out/poisson.json]"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import turbopy_b200 as tp  # noqa: E402
from oracle import grid as ogrid, poisson as opoisson  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def owner_for(ng, stretched):
    r = ogrid.grid_points(0.0, 1.0, ng)
    if stretched:
        r = r ** 1.25                                  # not a linspace: the kernels read the r array
    grid = SimpleNamespace(r=r, dr=ogrid.grid_dr(0.0, 1.0, ng), num_points=ng, r_min=0.0, r_max=1.0,
                           cell_widths=ogrid.cell_widths(r))
    return SimpleNamespace(clock=SimpleNamespace(dt=1e-12, num_steps=10), grid=grid)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rows = []
    for ng in ((1 << 24) + 1, 1 << 28):
        for stretched in (False, True):
            ow = owner_for(ng, stretched)
            ps = tp.PoissonSolver1DRadial(ow, {"type": "PoissonSolver1DRadial"})
            lin = tp.GridOnDevice.of(ow.grid, torch.device("cuda", 0)).linspace is not None
            assert lin != stretched
            src = np.random.default_rng(ng & 0xffff).normal(size=ng)
            y = torch.from_numpy(src).cuda()
            for _ in range(3):
                out = ps.solve(y)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                out = ps.solve(y)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            bpp = 24 if lin else 40
            row = {"points": ng, "grid": "linspace (r rebuilt)" if lin else "stretched (r read)", "ms": ms,
                   "G_points_per_s": ng / ms / 1e6, "bytes_per_point_moved": bpp,
                   "frac_of_measured_peak": ng * bpp / ms / 1e6 / PEAK,
                   "frac_algorithmic_16B": ng * 16 / ms / 1e6 / PEAK, "last_is_zero": float(out[-1]) == 0.0}
            if ng <= (1 << 24) + 1:
                ref = opoisson.solve(src, ow.grid.r, ow.grid.cell_widths)
                row["max_err_over_max_phi"] = float(np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max())
            rows.append(row)
            print(json.dumps(row), flush=True)
            del y, out, ps, ow
            tp.GridOnDevice._cache.clear()
            torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""interpolate1D(x, y)(x_new) at several sizes: CUDA-event time per call against the 16 B/point bound of a 1-D y
(8 + 8 k B/point for a (k, ng) y: --comps k).

    python tools/run_interp.py [--n 16777216 134217728] [--ng 1025 4097]
    TURBOPY_B200_INTERP_RING=0 python tools/run_interp.py     # the register-path kernel
"""
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


def peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[1 << 24, 1 << 27])
    ap.add_argument("--ng", type=int, nargs="+", default=[1025, 4097])
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--stretched", action="store_true", help="non-linspace abscissas ({x0, x1} table)")
    ap.add_argument("--comps", type=int, default=0, help="k > 0: a (k, ng) y (scipy _call_linear rounding)")
    args = ap.parse_args()
    owner = SimpleNamespace(clock=SimpleNamespace(dt=1e-12), grid=None)
    tool = tp.Interpolators(owner, {"type": "Interpolators"})
    for ng in args.ng:
        x = np.linspace(0.0, 1.0, ng)
        if args.stretched:
            x = x ** 1.3
        xd = torch.from_numpy(x).cuda()
        y = torch.cos(3.0 * xd)
        if args.comps > 0:
            y = torch.stack([torch.cos((c + 3.0) * xd) for c in range(args.comps)])
        f = tool.interpolate1D(xd, y)
        bpp = 16.0 if args.comps == 0 else 8.0 + 8.0 * args.comps
        f.deferred_check = True
        for n in args.n:
            xq = 0.05 + 0.9 * torch.rand(n, dtype=torch.float64, device="cuda")
            for _ in range(3):
                f(xq)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                f(xq)
            e1.record()
            torch.cuda.synchronize()
            f.check()
            ms = e0.elapsed_time(e1) / args.iters
            print(json.dumps({"ng": ng, "points": n, "ms": ms, "g_points_per_s": n / ms / 1e6,
                              "frac_of_hbm_peak": bpp * n / (ms * 1e-3) / 1e9 / peak(), "stretched": args.stretched,
                              "comps": args.comps, "bytes_per_point": bpp, "cells": os.environ.get("TURBOPY_B200_INTERP_CELLS", "1") != "0",
                              "ring": os.environ.get("TURBOPY_B200_INTERP_RING", "1") != "0",
                              "stages": os.environ.get("TURBOPY_B200_INTERP_STAGES", "6")}), flush=True)
            del xq


if __name__ == "__main__":
    main()

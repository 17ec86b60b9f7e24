#!/usr/bin/env python
"""Driver for profiling / sweeping the 6-component fused gather+push: a grid staged whole in shared memory
(ng <= ~1.5 k), or a larger grid through per-tile record windows / L2.

    python tools/run_gather6.py --ng 1048577 --n 16777216 --order sorted --formula interp
    python tools/run_gather6.py --sweep          # the table of profiles/r2_gather6.json
"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import turbopy_b200 as tp  # noqa: E402
from oracle import grid as ogrid  # noqa: E402  (Grid.r restatement, setup only)

Q_E, M_E, C = -1.6022e-19, 9.1094e-31, 2.9979e8


def peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


def run(ng, n, order, formula, steps, comps=6, quiet=False):
    r = ogrid.grid_points(0.0, 1.0, ng)
    dr = ogrid.grid_dr(0.0, 1.0, ng)
    owner = SimpleNamespace(clock=SimpleNamespace(dt=0.5 * dr / C),           # config 3: at most half a cell per step
                            grid=SimpleNamespace(r=r, dr=dr, num_points=ng, r_min=0.0, r_max=1.0))
    tool = tp.BorisPush(owner, {"type": "BorisPush"})
    rd = torch.from_numpy(r).cuda()
    E = torch.stack([1e3 * torch.cos(6.28 * (k + 1) * rd) for k in range(3)], dim=1).contiguous()
    B = torch.stack([torch.sin(6.28 * (k + 1) * rd) for k in range(3)], dim=1).contiguous() if comps == 6 else None
    g = torch.Generator(device="cuda").manual_seed(7)
    pos, mom = tp.soa_particles(n), tp.soa_particles(n)
    pos.copy_(0.05 + 0.9 * torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=g))
    mom.copy_(torch.randn((n, 3), dtype=torch.float64, device="cuda", generator=g) * (M_E * C * 0.1))
    if order == "sorted":
        tp.sort_by_cell(pos, mom, owner.grid)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B, formula=formula)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B, formula=formula)
    e1.record()
    torch.cuda.synchronize()
    tool.check_status()
    ms = e0.elapsed_time(e1) / steps
    row = {"ng": ng, "particles": n, "order": order, "formula": formula, "components": comps, "ms_per_step": ms,
           "g_pushes_per_s": n / ms / 1e6, "frac_of_hbm_peak": 96.0 * n / (ms * 1e-3) / 1e9 / peak(),
           "windows": os.environ.get("TURBOPY_B200_WINDOWS", "1") != "0"}
    wins = [w[0] for w in tool._windows.values() if w[0] is not None]
    if wins:
        b = wins[0].view(-1, 2)[: n // 960, 1].float()
        row["cells_per_tile_mean_max"] = [float(b.mean()), float(b.max())]
    if not quiet:
        print(json.dumps(row), flush=True)
    return row


def decay(ng, n, formula, blocks, steps_per_block):
    """Throughput block by block after ONE sort: how long a cell order lasts (dt = dr/(2c), thermal 0.1c momenta)."""
    r = ogrid.grid_points(0.0, 1.0, ng)
    dr = ogrid.grid_dr(0.0, 1.0, ng)
    owner = SimpleNamespace(clock=SimpleNamespace(dt=0.5 * dr / C),
                            grid=SimpleNamespace(r=r, dr=dr, num_points=ng, r_min=0.0, r_max=1.0))
    tool = tp.BorisPush(owner, {"type": "BorisPush"})
    rd = torch.from_numpy(r).cuda()
    E = torch.stack([1e3 * torch.cos(6.28 * (k + 1) * rd) for k in range(3)], dim=1).contiguous()
    B = torch.stack([torch.sin(6.28 * (k + 1) * rd) for k in range(3)], dim=1).contiguous()
    g = torch.Generator(device="cuda").manual_seed(7)
    pos, mom = tp.soa_particles(n), tp.soa_particles(n)
    pos.copy_(0.05 + 0.9 * torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=g))
    mom.copy_(torch.randn((n, 3), dtype=torch.float64, device="cuda", generator=g) * (M_E * C * 0.1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tp.sort_by_cell(pos, mom, owner.grid)
    e1.record()
    torch.cuda.synchronize()
    sort_ms = e0.elapsed_time(e1)
    rows = []
    import threading
    import time
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    samples, stop = [], threading.Event()

    def poll():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h),
                            pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU)))
            time.sleep(0.005)
    threading.Thread(target=poll, daemon=True).start()
    for b in range(blocks):
        mark = len(samples)
        e0.record()
        for _ in range(steps_per_block):
            tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B, formula=formula)
        e1.record()
        torch.cuda.synchronize()
        win = samples[mark:] or [(0, 0.0, 0, 0)]
        clk = {"sm_mhz_median": sorted(c for c, _, _, _ in win)[len(win) // 2], "power_w_max": max(p for _, p, _, _ in win),
               "reasons_or": hex(__import__("functools").reduce(lambda a, b: a | b, (r for _, _, r, _ in win), 0)),
               "temp_c": max(t for _, _, _, t in win)}
        ms = e0.elapsed_time(e1) / steps_per_block
        wins = [w[0] for w in tool._windows.values() if w[0] is not None]
        cells = float(wins[0].view(-1, 2)[: n // 960, 1].float().mean()) if wins else None
        rows.append({"steps_since_sort": (b + 1) * steps_per_block, "ms_per_step": ms,
                     "frac_of_hbm_peak": 96.0 * n / (ms * 1e-3) / 1e9 / peak(), "cells_per_tile_mean": cells, **clk})
        print(json.dumps(rows[-1]), flush=True)
    stop.set()
    tool.check_status()
    return {"ng": ng, "particles": n, "formula": formula, "sort_ms": sort_ms, "blocks": rows,
            "windows": os.environ.get("TURBOPY_B200_WINDOWS", "1") != "0"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--decay", type=int, default=0, help="blocks of 50 steps to time after one sort")
    ap.add_argument("--ng", type=int, default=1025)
    ap.add_argument("--n", type=int, default=16 << 20)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--formula", default="interp")
    ap.add_argument("--order", default="random", choices=["random", "sorted"])
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--json")
    a = ap.parse_args()
    if a.decay:
        out = decay(a.ng, a.n, a.formula, a.decay, 50)
        if a.json:
            with open(a.json, "w") as f:
                json.dump(out, f, indent=1)
        return
    if not a.sweep:
        run(a.ng, a.n, a.order, a.formula, a.steps)
        return
    rows = []
    for ng in (1025, (1 << 20) + 1):
        for order in ("random", "sorted"):
            for formula in ("interp", "create_interpolator", "interp1d_nd"):
                rows.append(run(ng, a.n, order, formula, a.steps))
    rows.append(run((1 << 20) + 1, 64 << 20, "sorted", "interp", a.steps))
    rows.append(run((1 << 20) + 1, 125_000_000, "sorted", "interp", a.steps))
    rows.append(run((1 << 20) + 1, 125_000_000, "random", "interp", a.steps))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()

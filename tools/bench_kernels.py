#!/usr/bin/env python
"""Roofline numbers for every kernel of the hot path (the driver's headline
bench is ../bench.py; this script feeds DESIGN.md / profiles/).

    python tools/bench_kernels.py [--quick] [--json out.json]

Each row: algorithmic bytes per unit x units / CUDA-event time, against the
measured HBM copy peak.  Inputs are larger than L2 unless the row says so.
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
from oracle import grid as ogrid  # noqa: E402  (Grid.r restatement, setup only)

Q_E, M_E = -1.6022e-19, 9.1094e-31


def peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


def owner_for(ng, dt=1e-12, rmax=1.0):
    r = ogrid.grid_points(0.0, rmax, ng)
    grid = SimpleNamespace(r=r, dr=ogrid.grid_dr(0.0, rmax, ng), num_points=ng, r_min=0.0, r_max=rmax,
                           cell_widths=ogrid.cell_widths(r))
    return SimpleNamespace(clock=SimpleNamespace(dt=dt, num_steps=10), grid=grid)


def particles(n, seed=0, lo=0.05, hi=0.95, sigma=0.1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    pos, mom = tp.soa_particles(n), tp.soa_particles(n)
    pos.copy_(lo + (hi - lo) * torch.rand((n, 3), dtype=torch.float64, device="cuda", generator=g))
    mom.copy_(torch.randn((n, 3), dtype=torch.float64, device="cuda", generator=g) * (M_E * 2.9979e8 * sigma))
    return pos, mom


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    pk = peak()
    rows = []

    def add(name, units, bytes_per_unit, seconds, note=""):
        gbs = units * bytes_per_unit / seconds / 1e9
        rows.append({"kernel": name, "units": units, "bytes_per_unit": bytes_per_unit, "ms": seconds * 1e3,
                     "units_per_s": units / seconds, "GBps": gbs, "frac_of_measured_peak": gbs / pk, "note": note})
        print(f"{name:58s} {units:>11d} units  {seconds*1e3:9.4f} ms  {units/seconds/1e9:8.2f} G/s  "
              f"{gbs:8.1f} GB/s  {100*gbs/pk:5.1f}%  {note}", flush=True)

    it = 5 if args.quick else 30
    n16 = 16 << 20

    # ---- kernel 1a: BorisPush.push, uniform E x B
    for n, note in ((1_000_000, "config 1 size: 48 MB of state, L2-resident"), (n16, ""), (64 << 20, "")):
        if args.quick and n > n16:
            continue
        tool = tp.BorisPush(owner_for(10), {"type": "BorisPush"})
        pos, mom = particles(n)
        E, B = np.array([0.0, 1e5, 0.0]), np.array([0.0, 0.0, 1.0])
        add("push uniform ExB (SoA)", n, 96, timed(lambda: tool.push(pos, mom, Q_E, M_E, E, B), it), note)
        if n == 1_000_000:
            with tp.StepGraph() as g:
                for _ in range(100):
                    tool.push(pos, mom, Q_E, M_E, E, B)
            add("push uniform ExB (SoA), CUDA graph of 100 steps", n, 96, timed(g.replay, 5) / 100, note)
            g.close()
        del pos, mom
    # AoS (numpy's layout) and per-particle fields (the un-fused path: 144 B/particle)
    tool = tp.BorisPush(owner_for(10), {"type": "BorisPush"})
    pos, mom = particles(n16)
    pa, ma = pos.contiguous(), mom.contiguous()
    for K in (4, 50):
        t = timed(lambda: tool.push_steps(pos, mom, Q_E, M_E, np.array([0.0, 1e5, 0.0]), np.array([0.0, 0.0, 1.0]), K), max(3, it // K))
        add(f"push_steps: {K} pushes per pass (static uniform ExB, temporal blocking)", n16 * K, 96, t,
            f"moves 96/{K} B per push: fp64-bound, not on the HBM roofline")
    add("push uniform ExB (AoS (n,3) C-order)", n16, 96,
        timed(lambda: tool.push(pa, ma, Q_E, M_E, np.array([0.0, 1e5, 0.0]), np.array([0.0, 0.0, 1.0])), it))
    Ep, Bp = tp.soa_particles(n16, fill=1e4), tp.soa_particles(n16, fill=0.5)
    add("push per-particle E,B (SoA, un-fused)", n16, 144, timed(lambda: tool.push(pos, mom, Q_E, M_E, Ep, Bp), it))
    Ea, Ba = Ep.contiguous(), Bp.contiguous()
    add("push per-particle E,B (SoA particles, C-order (n,3) fields)", n16, 144,
        timed(lambda: tool.push(pos, mom, Q_E, M_E, Ea, Ba), it))
    add("push per-particle E,B (all four arrays C-order (n,3))", n16, 144,
        timed(lambda: tool.push(pa, ma, Q_E, M_E, Ea, Ba), it))
    del pa, ma, Ep, Bp, Ea, Ba

    # ---- kernel 1b: fused gather + push
    for ng, ncomp, formula in ((30, 1, "interp"), (4097, 1, "interp"), (1025, 6, "interp"), (1025, 6, "create_interpolator"),
                               (1025, 6, "interp1d_nd"), ((1 << 20) + 1, 6, "interp")):
        ow = owner_for(ng, dt=1e-13)
        tool = tp.BorisPush(ow, {"type": "BorisPush"})
        r = torch.from_numpy(ow.grid.r).cuda()
        if ncomp == 1:
            E = (None, torch.cos(6.28 * r), None); B = None
        else:
            E = torch.stack([1e3 * torch.cos(6.28 * (k + 1) * r) for k in range(3)], dim=1).contiguous()
            B = torch.stack([torch.sin(6.28 * (k + 1) * r) for k in range(3)], dim=1).contiguous()
        pos.copy_(0.05 + 0.9 * torch.rand((n16, 3), dtype=torch.float64, device="cuda"))
        mom.zero_()
        t = timed(lambda: tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B, formula=formula), it)
        tool.check_status()
        add(f"gather+push ng={ng} comps={ncomp} formula={formula}", n16, 96, t,
            "grid in L2 (not staged)" if ng > 5000 else "grid staged in smem by TMA")
    # static field map: K fused gather+push steps per pass
    ow = owner_for(4097, dt=1e-13)
    tool = tp.BorisPush(ow, {"type": "BorisPush"})
    r = torch.from_numpy(ow.grid.r).cuda()
    Es = (None, torch.cos(6.28 * r), None)
    pos.copy_(0.05 + 0.9 * torch.rand((n16, 3), dtype=torch.float64, device="cuda")); mom.zero_()
    t = timed(lambda: tool.gather_push_steps(pos, mom, Q_E, M_E, 16, E_grid=Es), max(3, it // 8))
    tool.check_status()
    add("gather_push_steps: 16 steps per pass (static E_y map, ng=4097)", n16 * 16, 96, t, "fp64 / shared-memory bound, not on the HBM roofline")
    # sorted particles on the big grid (cell-coherent gathers)
    ow = owner_for((1 << 20) + 1, dt=1e-13)
    tool = tp.BorisPush(ow, {"type": "BorisPush"})
    r = torch.from_numpy(ow.grid.r).cuda()
    E = torch.stack([1e3 * torch.cos(6.28 * (k + 1) * r) for k in range(3)], dim=1).contiguous()
    B = torch.stack([torch.sin(6.28 * (k + 1) * r) for k in range(3)], dim=1).contiguous()
    xs = torch.sort(0.05 + 0.9 * torch.rand(n16, dtype=torch.float64, device="cuda")).values
    pos[:, 0].copy_(xs); mom.zero_()
    t = timed(lambda: tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B), it)
    add("gather+push ng=1048577 comps=6, particles sorted by x", n16, 96, t, "grid in L2")
    # the same after tp.sort_by_cell on a random ensemble, and the cost of the sort itself
    pos.copy_(0.05 + 0.9 * torch.rand((n16, 3), dtype=torch.float64, device="cuda")); mom.zero_()
    keep = pos.clone()

    def resort():
        pos.copy_(keep)
        tp.sort_by_cell(pos, mom, ow.grid)
    t_sort = timed(resort, max(3, it // 10)) - timed(lambda: pos.copy_(keep), max(3, it // 10))
    resort()                                    # leave the ensemble sorted for the next line
    add("sort_by_cell (20 kernels: stable 2-pass radix sort + in-place permute)", n16, 96, t_sort,
        "one-off; amortised over the steps until particles have drifted across bins")
    t = timed(lambda: tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B), it)
    tool.check_status()
    add("gather+push ng=1048577 comps=6, after sort_by_cell", n16, 96, t, "grid in L2, warps share record lines")
    del pos, mom, keep

    # ---- gather only: Interpolators.interpolate1D(x, y)(x_new), bounds check deferred (no sync per call)
    for ng in (1025, 4097):
        ow = owner_for(ng)
        it_tool = tp.Interpolators(ow, {"type": "Interpolators"})
        rr = torch.from_numpy(ow.grid.r).cuda()
        Y = torch.stack([torch.cos((k + 1) * rr) for k in range(6)])
        xq = 0.05 + 0.9 * torch.rand(n16, dtype=torch.float64, device="cuda")
        for label, f, bpp in ((f"interpolate1D ng={ng}, 1-D y (numpy.interp rounding)", it_tool.interpolate1D(rr, Y[0].contiguous()), 16),
                              (f"interpolate1D ng={ng}, (6,ng) y (scipy _call_linear rounding)", it_tool.interpolate1D(rr, Y), 56)):
            f.deferred_check = True
            add(label, n16, bpp, timed(lambda: f(xq), it))
            f.check()
        del xq
        if not args.quick:      # the 16 Mi-point call is a 50 us kernel (launch, table build and drain are ~9 us of it)
            xq = 0.05 + 0.9 * torch.rand(1 << 27, dtype=torch.float64, device="cuda")
            f = it_tool.interpolate1D(rr, Y[0].contiguous())
            f.deferred_check = True
            add(f"interpolate1D ng={ng}, 1-D y (numpy.interp rounding)", 1 << 27, 16, timed(lambda: f(xq), it))
            f.check()
            del xq

    # ---- kernel 2: FiniteDifference stencils
    for n in ((1 << 24) + 1, (1 << 28)):
        if args.quick and n > (1 << 25):
            continue
        ow = owner_for(n)
        fdt = tp.FiniteDifference(ow, {"type": "FiniteDifference", "method": "centered"})
        y = torch.sin(torch.linspace(0, 6.28, n, dtype=torch.float64, device="cuda"))
        add("fd centered_difference", n, 16, timed(lambda: fdt.centered_difference(y), it), "incl. torch.empty_like")
        lin = tp.GridOnDevice.of(ow.grid, y.device).linspace is not None     # r rebuilt in registers: 8 B/pt less
        lin_note = "linspace grid: r / cell_widths not read" if lin else "reads the r / cell_widths array"
        add("fd upwind_left", n, 16 if lin else 24, timed(lambda: fdt.upwind_left(y), it), lin_note)
        D = fdt.del2_radial()
        add("fd del2_radial @ f", n, 16 if lin else 24, timed(lambda: D @ y, it), lin_note)
        for name, bpp in (("ddx", 16), ("del2", 16), ("BC_left_extrap", 16), ("radial_curl", 40)):
            D = getattr(fdt, name)()
            add(f"fd {name}() @ f", n, bpp, timed(lambda: D @ y, it),
                "stored diagonals" if name == "radial_curl" else "constant diagonals in kernel args")
        ps = tp.PoissonSolver1DRadial(ow, {"type": "PoissonSolver1DRadial"})
        add("PoissonSolver1DRadial.solve (3 launches)", n, 24 if lin else 40, timed(lambda: ps.solve(y), it),
            ("24" if lin else "40") + " B/pt over the two reduce-then-scan passes; 16 B/pt algorithmic; " + lin_note)
        del y, fdt, D, ow, ps
        tp.GridOnDevice._cache.clear()
        torch.cuda.empty_cache()

    # ---- kernel 3: moment reduction
    for n in (n16, 128 << 20):
        if args.quick and n > n16:
            continue
        _, mom = particles(n)
        out = torch.empty(8, dtype=torch.float64, device="cuda")
        scratch = torch.empty(16384, dtype=torch.float64, device="cuda")
        add("moments reduction (KE, P, |p|^2)", n, 24,
            timed(lambda: tp.reduce_moments(mom, M_E, 2.9979e8 ** 2, out=out, scratch=scratch), it))
        del mom

    # ---- the whole cycle on the device: field update + gather/push + clock, CUDA graph of 20 steps
    ow = owner_for(4097, dt=5e-10)
    tool = tp.BorisPush(ow, {"type": "BorisPush"})
    pos, mom = particles(n16)
    mom.zero_()
    clock = tp.DeviceClock(0.0, 5e-10)
    wave = tp.PlaneWave(ow.grid, 1.0, 2e8)

    def cycle():
        wave.update(clock)
        tool.gather_push(pos, mom, Q_E, M_E, E_grid=(None, wave.E, None))
        clock.advance()

    cycle()
    with tp.StepGraph() as g:
        for _ in range(20):
            cycle()
    add("pif cycle (wave update + gather/push + clock), graph of 20 steps", n16, 96, timed(g.replay, 5) / 20)
    g.close()
    add("pif cycle, eager (3 launches per step)", n16, 96, timed(cycle, it))
    tool.check_status()

    if args.json:
        with open(args.json, "w") as f:
            json.dump({"peak_gbs": pk, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""[KineticEnergyDiagnostic.diagnose(); BorisPush.gather_push()] per step on the pif shape (configs[1]: 16 Mi particles,
E_y on a 4097-node grid) and with six components on a 1025-node grid, diagnostic every step: the standalone reduction
in front of the push vs the diagnostic riding on the push kernel (fuse_with_push).  CUDA events, 100 steps."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import turbopy_b200 as tp  # noqa: E402

dev = torch.device("cuda", 0)
peak, _ = bench.measured_peak()
n = 16 << 20
out = []
ONLY = os.environ.get("FUSED_KE_ONLY")          # e.g. "pif:fused" under ncu
for shape, ng in (("pif (E_y only)", 4097), ("six components", 1025), ("uniform E x B push()", 1025)):
    if ONLY and not shape.startswith(ONLY.split(":")[0]):
        continue
    r, dr = bench.grid_points(ng)
    pif = shape.startswith("pif")
    uniform = shape.startswith("uniform")
    owner = bench.Owner(r, dr, bench.DT_PIF if pif else bench.dt_of("config3", dr))
    pos_h, mom_h = bench.host_ensemble(n, 2345, thermal=not pif)
    if uniform:
        E, B = np.array([0.0, 1e5, 0.0]), np.array([0.0, 0.0, 1.0])
    elif pif:
        Ey = torch.from_numpy(bench.wave_fields(r, 1)[0]).to(dev)
        E, B = (None, Ey, None), None
    else:
        comps = bench.config3_fields(r)
        E = torch.from_numpy(np.stack(comps[:3], axis=1)).to(dev)
        B = torch.from_numpy(np.stack(comps[3:], axis=1)).to(dev)
    for fuse in (False, True, None):
        if ONLY and {"fused": True, "standalone": False, "none": None}[ONLY.split(":")[1]] is not fuse:
            continue
        pusher = tp.BorisPush(owner, {"type": "BorisPush"})
        pos, mom = tp.to_soa(pos_h, dev), tp.to_soa(mom_h, dev)
        diag = None
        if fuse is not None:
            diag = tp.KineticEnergyDiagnostic(owner, {"momentum": "p", "mass": bench.M_E, "fuse_with_push": fuse})
            diag.inspect_resource({"p": mom})
            diag.initialize()

        def step():
            if diag is not None:
                diag.diagnose()
            if uniform:
                pusher.push(pos, mom, bench.Q_E, bench.M_E, E, B)
            else:
                pusher.gather_push(pos, mom, bench.Q_E, bench.M_E, E_grid=E, B_grid=B)
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = int(os.environ.get("FUSED_KE_STEPS", "100"))
        e0.record()
        for _ in range(K):
            step()
        e1.record()
        torch.cuda.synchronize()
        if not uniform:
            pusher.check_status()
        ms = e0.elapsed_time(e1) / K
        row = {"shape": shape, "ng": ng, "particles": n,
               "diagnostic": "none" if fuse is None else "fused into the push" if fuse else "standalone reduction + push",
               "ms_per_step": ms, "G_steps_per_s": n / ms / 1e6,
               "frac_of_hbm_peak_at_96B": 96 * n / (ms * 1e-3) / 1e9 / peak,
               "ke_last": float(diag.values()[-1][0]) if diag is not None else None}
        out.append(row)
        print(json.dumps(row), flush=True)

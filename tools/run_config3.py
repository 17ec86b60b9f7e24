#!/usr/bin/env python
"""BASELINE configs[2] at full size under torchrun: 1 B particles sharded over the ranks (125 M each at 8 ranks),
2^20+1-node replicated grid with six smooth field components, fused gather+push every step, kinetic-energy
diagnostic + NCCL all-reduce every 10 steps.  Reports whole-job pushes/s (CUDA events, max over ranks) for the
particles in random order and after tp.sort_by_cell (dt at the grid's CFL limit, so the order persists).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_config3.py
"""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import turbopy_b200 as tp  # noqa: E402
from turbopy_b200.sharding import init_process_group  # noqa: E402
from oracle import grid as ogrid  # noqa: E402

Q_E, M_E, C = -1.6022e-19, 9.1094e-31, 2.9979e8


def main():
    rank, world, local = init_process_group()
    dev = torch.device("cuda", local)
    n_total = int(os.environ.get("CONFIG3_PARTICLES", 1_000_000_000))
    n = n_total // world
    ng = (1 << 20) + 1
    steps = int(os.environ.get("CONFIG3_STEPS", 40))
    r = ogrid.grid_points(0.0, 1.0, ng)
    dr = ogrid.grid_dr(0.0, 1.0, ng)
    dt = 0.5 * dr / C                                       # CFL: at most half a cell per step
    owner = SimpleNamespace(clock=SimpleNamespace(dt=dt, num_steps=steps),
                            grid=SimpleNamespace(r=r, dr=dr, num_points=ng, r_min=0.0, r_max=1.0))
    gen = torch.Generator(device=dev).manual_seed(3456 + rank)
    pos, mom = tp.soa_particles(n, dev), tp.soa_particles(n, dev)
    for c in range(3):
        pos[:, c].copy_(0.05 + 0.9 * torch.rand(n, dtype=torch.float64, device=dev, generator=gen))
        mom[:, c].copy_(torch.randn(n, dtype=torch.float64, device=dev, generator=gen) * (M_E * C * 0.1))
    rd = torch.from_numpy(r).to(dev)
    E = torch.stack([1e5 * torch.cos(2 * np.pi * (k + 1) * rd) for k in range(3)], dim=1).contiguous()
    B = torch.stack([torch.cos(2 * np.pi * (k + 4) * rd) for k in range(3)], dim=1).contiguous()
    pusher = tp.BorisPush(owner, {"type": "BorisPush"})
    diag = tp.KineticEnergyDiagnostic(owner, {"momentum": "p", "mass": M_E, "interval": 10})
    diag.inspect_resource({"p": mom})
    diag.initialize()

    def run(label):
        for _ in range(3):
            pusher.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            diag.diagnose()                                  # every 10th call: reduction + NCCL all-reduce of 8 doubles
            pusher.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        pusher.check_status()
        return {"order": label, "ms_per_step": float(ms) / steps, "pushes_per_s": n * world * steps / (float(ms) * 1e-3)}

    out = [run("random")]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tp.sort_by_cell(pos, mom, owner.grid)
    e1.record(); torch.cuda.synchronize()
    sort_ms = e0.elapsed_time(e1)
    out.append(run("after sort_by_cell"))
    ke = diag.values()
    if rank == 0:
        print(json.dumps({"config": "configs[2]: 1B particles sharded, 2^20+1-node replicated grid, 6 components, KE all-reduce every 10 steps",
                          "n_gpus": world, "particles_total": n * world, "particles_per_gpu": n, "steps": steps, "dt": dt,
                          "sort_ms_per_gpu": sort_ms, "runs": out,
                          "kinetic_energy_first_last": [float(ke[0][0]), float(ke[-1][0])] if len(ke) else None}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""End-to-end host path under torchrun: every rank streams its own pinned (n,3) arrays through its GPU at the same
time; prints pushes/s against the host-link ceiling (tp_host_link_probe, all ranks at once) for both bench workloads.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_e2e_scaling.py
Environment: TURBOPY_B200_HOST_CHUNKS (pipeline depth), E2E_PARTICLES, E2E_STEPS."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import turbopy_b200 as tp  # noqa: E402
from turbopy_b200 import _lib  # noqa: E402
from turbopy_b200.sharding import init_process_group  # noqa: E402

rank, world, local = init_process_group()
dev = torch.device("cuda", local)
n = int(os.environ.get("E2E_PARTICLES", 16 << 20))
steps = int(os.environ.get("E2E_STEPS", 10))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def reduce_max(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


up, down, both = C.c_double(), C.c_double(), C.c_double()
barrier()
_lib.check(_lib.lib().tp_host_link_probe(256 << 20, 4, C.byref(up), C.byref(down), C.byref(both)))
barrier()
duplex_min = -reduce_max(-both.value)
out = {"ranks": world, "chunks": os.environ.get("TURBOPY_B200_HOST_CHUNKS", "16"), "probe_h2d": up.value, "probe_d2h": down.value,
       "probe_duplex_min": duplex_min}
for workload, ng in (("pif", 4097), ("config3", (1 << 20) + 1)):
    r, dr = bench.grid_points(ng)
    pusher = tp.BorisPush(bench.Owner(r, dr, bench.dt_of(workload, dr)), {"type": "BorisPush"})
    pos_h, mom_h = bench.host_ensemble(n, 2345 + rank, thermal=workload != "pif")
    hp = torch.empty((n, 3), dtype=torch.float64).pin_memory(); hp.copy_(torch.from_numpy(pos_h)); hp = hp.numpy()
    hm = torch.empty((n, 3), dtype=torch.float64).pin_memory(); hm.copy_(torch.from_numpy(mom_h)); hm = hm.numpy()
    if workload == "pif":
        grid = dict(E_grid=(None, bench.wave_fields(r, 1)[0], None), B_grid=None)
    else:
        comps = bench.config3_fields(r)
        grid = dict(E_grid=np.stack(comps[:3], axis=1), B_grid=np.stack(comps[3:], axis=1))
    for _ in range(2):      # two warm-up calls: the library page-locks large host arrays the SECOND time it sees them
        pusher.gather_push(hp, hm, bench.Q_E, bench.M_E, **grid)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        pusher.gather_push(hp, hm, bench.Q_E, bench.M_E, **grid)
    barrier()
    dt = reduce_max(time.perf_counter() - t0)
    value = n * world * steps / dt
    ceiling = world * duplex_min * 1e9 / 48
    out[workload] = {"pushes_per_s": value, "frac_of_ceiling": value / ceiling, "gbs_per_direction_per_gpu": value / world * 48 / 1e9}
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()

#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU): sharded particles,
replicated grid, fused gather+push per rank, kinetic-energy diagnostic with the
NCCL all-reduce -- global sums against the oracle evaluated on the whole ensemble.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_nccl_check.py
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import turbopy_b200 as tp  # noqa: E402
from turbopy_b200.sharding import init_process_group, shard_range  # noqa: E402
from oracle import boris as oboris, gather as ogather, grid as ogrid, moments as omom  # noqa: E402

Q_E, M_E = -1.6022e-19, 9.1094e-31


def main():
    rank, world, local = init_process_group()
    dev = torch.device("cuda", local)
    n_total, ng, dt, steps = 400_003, 1025, 1e-12, 3
    r = ogrid.grid_points(0.0, 1.0, ng)
    owner = SimpleNamespace(clock=SimpleNamespace(dt=dt, num_steps=steps),
                            grid=SimpleNamespace(r=r, dr=ogrid.grid_dr(0.0, 1.0, ng), num_points=ng, r_min=0.0, r_max=1.0))
    rng = np.random.default_rng(3456)                       # every rank builds the same global ensemble
    pos0 = rng.uniform(0.1, 0.9, (n_total, 3))
    mom0 = M_E * 2.9979e8 * rng.normal(0, 0.1, (n_total, 3))
    comps = [1e5 * np.cos(2 * np.pi * (k + 1) * r) for k in range(3)] + [np.sin(2 * np.pi * (k + 1) * r) for k in range(3)]
    lo, hi = shard_range(n_total, rank, world)
    pos, mom = tp.to_soa(pos0[lo:hi], dev), tp.to_soa(mom0[lo:hi], dev)
    E = torch.from_numpy(np.stack(comps[:3], axis=1)).to(dev)
    B = torch.from_numpy(np.stack(comps[3:], axis=1)).to(dev)
    pusher = tp.BorisPush(owner, {"type": "BorisPush"})
    diag = tp.KineticEnergyDiagnostic(owner, {"momentum": "p", "mass": M_E})
    diag.inspect_resource({"p": mom})
    diag.initialize()
    for _ in range(steps):
        diag.diagnose()                                     # reduction + NCCL all-reduce of 8 doubles
        pusher.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)
    diag.diagnose()
    pusher.check_status()
    rows = diag.values()
    # oracle on the whole ensemble
    rp, rm = pos0.copy(), mom0.copy()
    ref_rows = []
    for s in range(steps + 1):
        ref_rows.append(omom.moments(rm[:, 0], rm[:, 1], rm[:, 2], M_E, 2.9979e8 ** 2)[:6])
        if s < steps:
            F = [ogather.interp_b(rp[:, 0], r, c)[0] for c in comps]
            oboris.push_aos(rp, rm, Q_E, M_E, np.stack(F[:3], axis=1), np.stack(F[3:], axis=1), dt)
    ref_rows = np.array(ref_rows)
    scale = np.maximum(np.abs(ref_rows), np.abs(mom0).sum() * np.array([0, 1, 1, 1, 0, 0])[None, :] + 1e-300)
    assert rows.shape == ref_rows.shape, (rows.shape, ref_rows.shape)
    assert np.all(np.abs(rows - ref_rows) <= 1e-12 * scale), (rows, ref_rows)     # sum order differs: 1e-12
    assert rows[0, 5] == n_total
    got_p, got_m = pos.cpu().numpy(), mom.cpu().numpy()
    assert np.array_equal(got_p.view(np.uint64), rp[lo:hi].view(np.uint64))       # shard = exact subset of the whole
    assert np.array_equal(got_m.view(np.uint64), rm[lo:hi].view(np.uint64))
    torch.distributed.barrier() if world > 1 else None
    if rank == 0:
        print(f"dist check ok: world={world}, {n_total} particles sharded, KE row0={rows[0,0]:.6e} J "
              f"(ref {ref_rows[0,0]:.6e}), shards bit-identical to the oracle")
    graph_check(rank, world, dev, owner, pos0[lo:hi], mom0[lo:hi], E, B)
    fused_check(rank, world, dev, owner, pos0[lo:hi], mom0[lo:hi], E, B, ref_rows, rp[lo:hi], rm[lo:hi], n_total)
    if world > 1:
        torch.distributed.destroy_process_group()


def graph_check(rank, world, dev, owner, pos0, mom0, E, B):
    """`[10 x gather_push, reduce_moments, all-reduce]` is ONE stream-ordered sequence -- the all-reduce
    runs on the compute stream through the path's own NCCL communicator -- so it can be captured in a StepGraph and
    replayed.  Three replays (30 steps, 3 all-reduces) against the same 30 steps issued eagerly: particles
    bit-identical, reduced rows identical (same per-rank partials, same NCCL reduction order)."""
    from turbopy_b200.diagnostics import reduce_moments
    from turbopy_b200.sharding import allreduce_sum_

    def fresh():
        pusher = tp.BorisPush(owner, {"type": "BorisPush"})
        return pusher, tp.to_soa(pos0, dev), tp.to_soa(mom0, dev)

    def block(pusher, pos, mom, row, scratch):
        for _ in range(10):
            pusher.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)
        reduce_moments(mom, M_E, pusher.c2, out=row, scratch=scratch)
        allreduce_sum_(row)                                  # tp_nccl_allreduce_sum_f64 on the current stream

    from turbopy_b200 import _lib
    scratch = torch.empty(int(_lib.lib().tp_reduce_scratch_doubles()), dtype=torch.float64, device=dev)
    # eager
    pusher, pe, me = fresh()
    rows_e = torch.zeros((3, 8), dtype=torch.float64, device=dev)
    for k in range(3):
        block(pusher, pe, me, rows_e[k], scratch)
    torch.cuda.synchronize()
    # captured once, replayed three times (the graph writes its row to one buffer; copy it out between replays)
    pusher_g, pg, mg = fresh()
    row = torch.zeros(8, dtype=torch.float64, device=dev)
    allreduce_sum_(row.clone())                              # communicator exists before the capture starts
    pusher_g.gather_push(pg, mg, 0.0, M_E, E_grid=E, B_grid=B)   # warm every lazily allocated buffer ...
    pg.copy_(tp.to_soa(pos0, dev)); mg.copy_(tp.to_soa(mom0, dev))   # ... and restore the start state
    with tp.StepGraph(dev) as graph:
        block(pusher_g, pg, mg, row, scratch)
    rows_g = []
    for k in range(3):
        graph.replay()
        rows_g.append(row.clone())
    torch.cuda.synchronize()
    pusher_g.check_status()
    assert torch.equal(pe, pg) and torch.equal(me, mg), "graph replay differs from eager steps"
    got = torch.stack(rows_g).cpu().numpy()
    want = rows_e.cpu().numpy()
    assert np.array_equal(got, want), (got, want)
    assert got[0, 5] == 400_003                              # the count over all ranks
    if world > 1:
        torch.distributed.barrier()
    if rank == 0:
        print(f"graph check ok: world={world}, [10 x gather_push, reduce, NCCL all-reduce on the compute stream] captured "
              f"once, replayed 3x == eager; KE after 30 steps {got[2, 0]:.6e} J")


def fused_check(rank, world, dev, owner, pos0, mom0, E, B, ref_rows, rp, rm, n_total):
    """The same with the diagnostic riding on the push (fuse_with_push): rows against the oracle on the whole
    ensemble, shards bit-identical; then `[diagnose(); 10 x gather_push]` (the first push delivers the row, the NCCL
    all-reduce follows on the compute stream) captured in a StepGraph and replayed 3x against the eager sequence."""
    steps = ref_rows.shape[0] - 1
    pusher = tp.BorisPush(owner, {"type": "BorisPush"})
    pos, mom = tp.to_soa(pos0, dev), tp.to_soa(mom0, dev)
    diag = tp.KineticEnergyDiagnostic(owner, {"momentum": "p", "mass": M_E, "fuse_with_push": True})
    diag.inspect_resource({"p": mom})
    diag.initialize()
    for _ in range(steps):
        diag.diagnose()                                     # posts; the push below delivers and all-reduces
        pusher.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)
    diag.diagnose()                                         # the last row: flushed by values() (standalone + all-reduce)
    pusher.check_status()
    rows = diag.values()
    assert diag.fused_rows == steps, diag.fused_rows
    scale = np.maximum(np.abs(ref_rows), np.abs(rm).sum() * world * np.array([0, 1, 1, 1, 0, 0])[None, :] + 1e-300)
    assert rows.shape == ref_rows.shape and np.all(np.abs(rows - ref_rows) <= 1e-12 * scale), (rows, ref_rows)
    assert rows[0, 5] == n_total
    assert np.array_equal(pos.cpu().numpy().view(np.uint64), rp.view(np.uint64))
    assert np.array_equal(mom.cpu().numpy().view(np.uint64), rm.view(np.uint64))

    def fresh():
        pusher = tp.BorisPush(owner, {"type": "BorisPush"})
        p, m = tp.to_soa(pos0, dev), tp.to_soa(mom0, dev)
        d = tp.KineticEnergyDiagnostic(owner, {"momentum": "p", "mass": M_E, "fuse_with_push": True})
        d.inspect_resource({"p": m})
        d.initialize()
        return pusher, p, m, d

    def block(pusher, p, m, d):
        d.diagnose()
        for _ in range(10):
            pusher.gather_push(p, m, Q_E, M_E, E_grid=E, B_grid=B)

    pusher_e, pe, me, de = fresh()
    for _ in range(3):
        block(pusher_e, pe, me, de)
    rows_e = de.values()
    assert de.fused_rows == 3
    pusher_g, pg, mg, dg = fresh()
    block(pusher_g, pg, mg, dg)                              # warm: buffers, communicator, the diagnostic's row table
    pg.copy_(tp.to_soa(pos0, dev)); mg.copy_(tp.to_soa(mom0, dev))
    with tp.StepGraph(dev) as graph:
        block(pusher_g, pg, mg, dg)
    captured_row = dg._used - 1
    rows_g = []
    for _ in range(3):
        graph.replay()
        rows_g.append(dg.rows[captured_row, :6].clone())
    torch.cuda.synchronize()
    pusher_g.check_status()
    assert torch.equal(pe, pg) and torch.equal(me, mg), "fused graph replay differs from eager steps"
    got = torch.stack(rows_g).cpu().numpy()
    assert np.array_equal(got, rows_e), (got, rows_e)
    if world > 1:
        torch.distributed.barrier()
    if rank == 0:
        print(f"fused check ok: world={world}, diagnostic delivered by the push kernel ({steps} rows vs the oracle on the "
              f"whole ensemble, 1e-12), [diagnose, 10 x gather_push] with the NCCL all-reduce behind the first push captured "
              f"once, replayed 3x == eager; KE rows {got[:, 0]}")


if __name__ == "__main__":
    main()

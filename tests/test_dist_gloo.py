"""CPU, world_size 2 over gloo: the N>1 path of the hot path is "shard the
particles, replicate the grid, all-reduce the diagnostic scalars" -- the host
logic of exactly that."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import moments as omom


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from turbopy_b200.sharding import allreduce_sum_, init_process_group, rank_seed, shard_range
    r, w, _ = init_process_group(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = shard_range(n_total, rank, world)
    # every rank builds the same global ensemble and keeps its block (particles are independent)
    rng = np.random.default_rng(1234)
    mom = 9.1094e-31 * 2.9979e8 * rng.normal(0, 0.1, (n_total, 3))
    mine = mom[lo:hi]
    part = omom.moments(mine[:, 0], mine[:, 1], mine[:, 2], 9.1094e-31, 2.9979e8 ** 2)
    t = torch.from_numpy(part.copy())
    allreduce_sum_(t)                                     # what the diagnostic does with its 8-double row
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.concatenate([t.numpy(), part, [lo, hi, rank_seed(3456, rank)]]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_moments_allreduce(tmp_path):
    world, n_total = 2, 100_001
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    rows = [np.load(tmp_path / f"rank{r}.npy") for r in range(world)]
    reduced = [row[:8] for row in rows]
    parts = [row[8:16] for row in rows]
    assert np.array_equal(reduced[0], reduced[1])          # every rank holds the same global sums
    from turbopy_b200.sharding import combine_moments
    assert np.array_equal(reduced[0], combine_moments([torch.from_numpy(p.copy()) for p in parts]).numpy())
    rng = np.random.default_rng(1234)
    mom = 9.1094e-31 * 2.9979e8 * rng.normal(0, 0.1, (n_total, 3))
    whole = omom.moments(mom[:, 0], mom[:, 1], mom[:, 2], 9.1094e-31, 2.9979e8 ** 2)
    assert abs(reduced[0][0] - whole[0]) <= 1e-12 * whole[0]   # global KE to 1e-12 (sum order differs)
    assert reduced[0][5] == n_total
    assert [int(r[16]) for r in rows] == [0, 50001] and [int(r[17]) for r in rows] == [50001, 100001]
    assert [int(r[18]) for r in rows] == [3456, 3457]


def _worker_deferred(rank, world, port, n_total, out_dir):
    """The diagnostic's deferred rows at world size 2: a row delivered by the push and a row flushed by values() are
    both all-reduced (the same collective sequence on every rank)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from types import SimpleNamespace
    import turbopy_b200 as tp
    from turbopy_b200 import diagnostics as D
    from turbopy_b200.sharding import init_process_group, shard_range
    init_process_group(backend="gloo")
    m, c2 = 9.1094e-31, 2.9979e8 ** 2
    lo, hi = shard_range(n_total, rank, world)
    rng = np.random.default_rng(77)
    mom = torch.from_numpy((m * 2.9979e8 * rng.normal(0, 0.1, (n_total, 3)))[lo:hi].copy())

    def host_reduce(momentum, mass, c2_, out=None, scratch=None):      # stand-in for the CUDA reduction
        a = momentum.numpy()
        out[:] = torch.from_numpy(omom.moments(a[:, 0], a[:, 1], a[:, 2], mass, c2_))
        return out

    D.reduce_moments = host_reduce
    D._can_ride_on_push = lambda t: True
    D._lib.lib = lambda: SimpleNamespace(tp_reduce_scratch_doubles=lambda: 8)
    owner = SimpleNamespace(clock=SimpleNamespace(dt=1e-12, num_steps=4), grid=None)
    d = tp.KineticEnergyDiagnostic(owner, {"momentum": "p", "mass": m, "fuse_with_push": True})
    d.inspect_resource({"p": mom})
    d.initialize()
    d.diagnose()
    taken = D.take_moments_request(mom)                    # what BorisPush.gather_push does ...
    assert taken is d
    mass, mass_sq, c2_d, row, scratch = d.moments_request()
    host_reduce(mom, mass, c2_d, out=row)                  # ... the kernel fills the row ...
    part0 = row.clone()
    d.moments_delivered()                                  # ... and the all-reduce follows
    mom.mul_(0.5)
    d.diagnose()                                           # posted, never taken:
    rows = d.values()                                      # flushed here, all-reduced as well
    a = mom.numpy()
    part1 = omom.moments(a[:, 0], a[:, 1], a[:, 2], m, c2)
    np.save(os.path.join(out_dir, f"def{rank}.npy"), np.concatenate([rows.ravel(), part0.numpy()[:6], part1[:6]]))
    dist.barrier()
    dist.destroy_process_group()


def test_deferred_diagnostic_rows_are_allreduced(tmp_path):
    world, n_total = 2, 20_001
    mp.spawn(_worker_deferred, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    out = [np.load(tmp_path / f"def{r}.npy") for r in range(world)]
    rows = [o[:12].reshape(2, 6) for o in out]
    assert np.array_equal(rows[0], rows[1])
    assert np.array_equal(rows[0][0], out[0][12:18] + out[1][12:18])      # delivered row: sum of the ranks' partials
    assert np.array_equal(rows[0][1], out[0][18:24] + out[1][18:24])      # flushed row likewise
    assert rows[0][0][5] == n_total and rows[0][1][5] == n_total
    assert abs(rows[0][1][0] / rows[0][0][0] - 0.25) < 1e-2               # momenta halved: KE ~ quartered (0.1c thermal)

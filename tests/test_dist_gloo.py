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

"""Particle sharding across the GPUs of one node.

Particles are independent given the fields (turbopy/computetools.py:427-441 has
no cross-particle term and the reference has no deposition step), so the path
shards by particle index with no data-path collective: rank r owns
``[r*N/P, (r+1)*N/P)``, the small 1-D grid is replicated, and the only exchange
is the all-reduce of the <= 8 diagnostic scalars.
"""
import os

import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Half-open index range of ``rank``'s block (sizes differ by at most 1)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(base_seed, rank):
    """Per-rank RNG seed for synthetic ensembles (config 3: 3456 + rank)."""
    return int(base_seed) + int(rank)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), \
        int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """One process per GPU (torchrun); NCCL on GPUs, gloo for the CPU tests."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def allreduce_sum_(t):
    """In-place SUM all-reduce of the diagnostic scalars (NCCL over NVLink on
    GPUs); a no-op in a single-process run."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def combine_moments(parts):
    """Fold per-rank moment vectors ``[KE, Px, Py, Pz, |p|^2, n, 0, 0]`` the way
    the all-reduce does (plain sums)."""
    out = parts[0].clone()
    for p in parts[1:]:
        out += p
    return out

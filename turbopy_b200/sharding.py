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


class DiagnosticComm:
    """The path's own NCCL communicator (one rank per GPU), used for exactly one thing: the in-place SUM all-reduce
    of the diagnostic's 8 doubles, enqueued on the CURRENT (compute) stream through the C ABI
    (tp_nccl_allreduce_sum_f64).  That keeps ``reduce_moments`` -> all-reduce stream-ordered with the push kernels
    and capturable in a ``StepGraph``; ``torch.distributed.all_reduce`` runs on the process group's
    own stream behind event waits instead.  torch.distributed is only the bootstrap: it carries the 128-byte
    NCCL unique id from rank 0 to the other ranks."""

    _instance = None

    def __init__(self):
        import ctypes as C

        from . import _lib
        self._C, self._lib = C, _lib
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        lib = _lib.lib()
        ident = (C.c_char * 128)()
        if self.rank == 0:
            _lib.check(lib.tp_nccl_unique_id(ident))
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0)                # pickled through the process group (CPU or GPU backend)
        ident = (C.c_char * 128).from_buffer_copy(box[0])
        self._comm = C.c_void_p()
        _lib.check(lib.tp_nccl_comm_init(ident, self.world, self.rank, C.byref(self._comm)))

    @classmethod
    def get(cls):
        """The process-wide communicator, created collectively on first use (every rank must get here)."""
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def allreduce_sum_(self, t):
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise TypeError("DiagnosticComm reduces contiguous CUDA float64 tensors")
        C = self._C
        self._lib.check(self._lib.lib().tp_nccl_allreduce_sum_f64(
            self._comm, C.c_void_p(t.data_ptr()), t.numel(), C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)))
        return t

    def close(self):
        if self._comm:
            self._lib.lib().tp_nccl_comm_destroy(self._comm)
            self._comm = self._C.c_void_p()
        if DiagnosticComm._instance is self:
            DiagnosticComm._instance = None


def allreduce_sum_(t):
    """In-place SUM all-reduce of the diagnostic scalars; a no-op in a single-process run.  CUDA tensors go
    through the path's own NCCL communicator on the current stream (:class:`DiagnosticComm`, NVLink / NVSwitch);
    host tensors (the gloo tests) through ``torch.distributed``.  ``TURBOPY_B200_TORCH_ALLREDUCE=1`` forces the
    latter for CUDA tensors too (A-B switch)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if t.is_cuda and os.environ.get("TURBOPY_B200_TORCH_ALLREDUCE", "0") != "1":
            DiagnosticComm.get().allreduce_sum_(t)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def combine_moments(parts):
    """Fold per-rank moment vectors ``[KE, Px, Py, Pz, |p|^2, n, 0, 0]`` the way
    the all-reduce does (plain sums)."""
    out = parts[0].clone()
    for p in parts[1:]:
        out += p
    return out

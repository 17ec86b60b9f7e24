"""Particle storage helpers: the SoA layout the kernels are fastest on, behind
the (n,3) logical shape turboPy's callers and diagnostics index
(e.g. ``data[0, :]``, tests/fixtures/particle_in_field/particle_in_field.py:78).
"""
import numpy as np
import torch


#: bumped by every sort_by_cell: tools that keep per-tile state about an ensemble (the windowed large-grid gather)
#: re-derive it when the particle order changed under them
order_generation = 0


def soa_particles(n, device="cuda", fill=None):
    """An ``(n,3)`` float64 view whose memory is three contiguous component
    arrays (strides ``(1, ld)``, ``ld`` = ``n`` rounded up to even so every
    component starts 16-byte aligned for 128-bit accesses)."""
    ld = max(2, (int(n) + 1) // 2 * 2)
    buf = torch.empty((3, ld), dtype=torch.float64, device=device)
    if fill is not None:
        buf.fill_(fill)
    return buf.t()[:n]


def to_soa(array, device="cuda"):
    """Copy an ``(n,3)`` numpy array / tensor into a fresh SoA view on ``device``."""
    src = torch.as_tensor(np.asarray(array) if not isinstance(array, torch.Tensor) else array,
                          dtype=torch.float64)
    out = soa_particles(src.shape[0], device)
    out.copy_(src.to(device, non_blocking=False))
    return out


def components(view):
    """The three contiguous component tensors (x, y, z) of an SoA view."""
    return view[:, 0], view[:, 1], view[:, 2]


def sort_by_cell(position, momentum, grid, axis=0, cells_per_bin_log2=0, extra=(), return_perm=False):
    """Reorder the particles in place (same storage) by the grid cell they sit
    in along ``axis``.

    Additive helper (the reference never reorders particles; its physics does
    not depend on particle order).  It pays off for the fused gather on grids
    too large for shared memory: with particles in random order every lane of
    a warp reads its node records from its own cache line and the kernel is
    bound by L1 wavefronts; ordered by cell, a warp shares a few lines.  The
    sort is *stable* (key = cell index ``>> cells_per_bin_log2``, ties keep
    their order), so the result is reproducible:
    ``perm == numpy.argsort(key, kind="stable")`` (oracle/sort.py).  When
    particles move a fraction of a cell per step (dt at the grid's CFL limit)
    the order decays slowly and one sort every ~50-100 steps is enough.

    ``extra``: further per-particle float64 CUDA tensors (1-D, or (n,k) whose
    columns are permuted one by one) to reorder with the same permutation.
    Returns ``perm`` (``new[i] = old[perm[i]]``, int64 CUDA tensor) if
    ``return_perm``.  Asynchronous on the current stream.
    """
    import ctypes as C

    from . import _lib
    from .tools import _is_cuda, _on_device, _particles_struct, _stream_ptr

    if not (_is_cuda(position) and _is_cuda(momentum)) or position.device != momentum.device:
        raise TypeError("position and momentum must be CUDA float64 tensors on one device")
    if axis not in (0, 1, 2):
        raise ValueError("axis must be 0, 1 or 2")
    p = _particles_struct(position, momentum)
    n = int(p.n)
    dev = position.device
    ng = int(grid.num_points) if hasattr(grid, "num_points") else int(len(grid.r))
    ncells = ng - 1
    lib = _lib.lib()
    need = int(lib.tp_sort_scratch_bytes(n, ncells, int(cells_per_bin_log2)))
    if need < 0:
        raise ValueError("sort_by_cell: need 0 <= n < 2**32, at least one cell and 0 <= cells_per_bin_log2 <= 40")
    global order_generation
    order_generation += 1
    with _on_device(dev):
        scratch = torch.empty(max(need, 8), dtype=torch.uint8, device=dev)
        perm = torch.empty(n, dtype=torch.int64, device=dev) if (return_perm or len(extra)) else None
        _lib.check(lib.tp_sort_by_cell_f64(C.byref(p), int(axis), float(grid.r[0]), float(grid.dr), ncells,
                                           int(cells_per_bin_log2), C.c_void_p(scratch.data_ptr()),
                                           C.c_void_p(perm.data_ptr()) if perm is not None else None,
                                           _stream_ptr(dev)))
        if len(extra):
            tmp = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
            for t in extra:
                if not _is_cuda(t) or t.device != dev or t.dtype != torch.float64 or t.shape[0] != n or t.ndim > 2:
                    raise TypeError("extra arrays must be float64 CUDA tensors of shape (n,) or (n,k) on the particles' device")
                cols = [t] if t.ndim == 1 else [t[:, c] for c in range(t.shape[1])]
                for col in cols:
                    _lib.check(lib.tp_permute_f64(C.c_void_p(col.data_ptr()), int(col.stride(0)) if n > 1 else 1, n,
                                                  C.c_void_p(perm.data_ptr()), C.c_void_p(tmp.data_ptr()), _stream_ptr(dev)))
    return perm if return_perm else None


def resort_local(position, momentum, grid, axis=0, phase=0):
    """Restore the cell order of an ALREADY nearly ordered ensemble in place, at the cost of one push step: a stable
    sort by cell inside blocks of 1920 consecutive particles (tp_resort_blocks_f64; the global :func:`sort_by_cell`
    costs ~18 push steps).  ``phase`` alternates the block boundaries (even: blocks start at particle 0, odd: at
    particle 960) so that repeated calls let particles migrate between blocks.  SoA ensembles
    (:func:`soa_particles`) only.  Additive helper, asynchronous on the current stream."""
    import ctypes as C

    from . import _lib
    from .tools import _is_cuda, _on_device, _particles_struct, _stream_ptr

    if not (_is_cuda(position) and _is_cuda(momentum)) or position.device != momentum.device:
        raise TypeError("position and momentum must be CUDA float64 tensors on one device")
    if axis not in (0, 1, 2):
        raise ValueError("axis must be 0, 1 or 2")
    p = _particles_struct(position, momentum)
    ng = int(grid.num_points) if hasattr(grid, "num_points") else int(len(grid.r))
    lib = _lib.lib()
    offset = (int(lib.tp_resort_block_particles()) // 2) if (int(phase) & 1) else 0
    global order_generation
    order_generation += 1
    with _on_device(position.device):
        _lib.check(lib.tp_resort_blocks_f64(C.byref(p), int(axis), float(grid.r[0]), float(grid.dr), ng - 1, offset,
                                            _stream_ptr(position.device)))

"""Particle storage helpers: the SoA layout the kernels are fastest on, behind
the (n,3) logical shape turboPy's callers and diagnostics index
(e.g. ``data[0, :]``, tests/fixtures/particle_in_field/particle_in_field.py:78).
"""
import numpy as np
import torch


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

"""Oracle of the particle reordering (test infrastructure, not product code).

The reference never reorders particles, so there is no reference function to
follow: the specification is the one stated in include/turbopy_b200.h -- a
STABLE sort by the cell index (optionally coarsened by 2**cells_per_bin_log2) --
restated here with numpy's stable argsort.  The key uses the same two rounded
fp64 operations as the kernel.
"""
import numpy as np


def keys(x, r_first, dr, ncells, cells_per_bin_log2=0):
    inv_dr = 1.0 / dr
    t = (np.asarray(x, dtype=np.float64) - r_first) * inv_dr
    with np.errstate(invalid="ignore"):
        ok = t >= 0.0                                   # NaN and negatives -> cell 0
    c = np.zeros(t.shape, dtype=np.int64)
    c[ok] = np.minimum(t[ok], 9.0e18).astype(np.int64)  # trunc == floor for t >= 0
    c = np.minimum(c, ncells - 1)
    return c >> cells_per_bin_log2


def permutation(x, r_first, dr, ncells, cells_per_bin_log2=0):
    """new[i] = old[perm[i]]"""
    return np.argsort(keys(x, r_first, dr, ncells, cells_per_bin_log2), kind="stable")

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


def resort_blocks_permutation(x, r_first, dr, ncells, offset=0, block=1920):
    """tp_resort_blocks_f64: the stable order by cell applied independently inside blocks of ``block`` consecutive
    particles (block 0 = [0, offset), then [offset + k*block, offset + (k+1)*block)); inside a block the key is
    ``min(cell - lowest cell of the block, 2**20 - 1)``.  new[i] = old[perm[i]]."""
    k = keys(x, r_first, dr, ncells)
    n = len(k)
    perm = np.arange(n)
    edges = ([0] if offset > 0 else []) + list(range(min(offset, n), n, block)) + [n]
    for lo, hi in zip(edges[:-1], edges[1:]):
        if hi > lo:
            kb = np.minimum(k[lo:hi] - k[lo:hi].min(), (1 << 20) - 1)
            perm[lo:hi] = lo + np.argsort(kb, kind="stable")
    return perm

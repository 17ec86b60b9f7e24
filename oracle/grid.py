"""Oracle restatement of the ``Grid`` attributes the hot path reads -- TEST
INFRASTRUCTURE ONLY.  turbopy/core.py:610-673,709."""
import numpy as np


def grid_points(r_min, r_max, n):
    """``Grid.r``: core.py:666-667 with ``generate_linear`` = np.linspace(0,1,n)
    (:709).  numpy's linspace is ``arange(n)*step`` with ``step = 1/(n-1)`` and
    the last element overwritten by the stop value, hence:"""
    step = 1.0 / (n - 1)
    lin = np.arange(0, n).astype(np.float64) * step
    lin[-1] = 1.0
    return r_min + (r_max - r_min) * lin


def grid_dr(r_min, r_max, n):
    """``Grid.dr`` for the ``"N"`` input form, core.py:625."""
    return (r_max - r_min) / (n - 1)


def cell_widths(r):
    """``Grid.cell_widths``, core.py:670."""
    return r[1:] - r[:-1]

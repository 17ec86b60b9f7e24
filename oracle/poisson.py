"""Oracle restatement of ``PoissonSolver1DRadial.solve`` -- TEST INFRASTRUCTURE
ONLY.  turbopy/computetools.py:35-59.  The reference has no test for it
(nothing under tests/ calls it): parity is pinned by ``make_golden.py`` against
the imported reference only.

numpy's ``cumsum`` adds strictly left to right; a parallel scan associates
differently, so the CUDA result is compared with a tolerance (stated in the
test), not bit for bit.
"""
import numpy as np


def solve(sources, r, widths):
    """Two running sums: I1 = cumsum(r*s*dr), then cumsum of I1*dr/r with the
    r = 0 entry replaced by a linear extrapolation that is subtracted again
    (so the derivative vanishes on the axis), finally shifted so that the last
    point is zero."""
    r = np.asarray(r, dtype=np.float64)
    s = np.asarray(sources, dtype=np.float64)
    dr = np.mean(widths)                                  # :49
    first = np.cumsum((r * s) * dr)                       # :50
    with np.errstate(divide="ignore", invalid="ignore"):
        g = (first * dr) / r                              # :51
    axis_value = 2 * g[1] - g[2]                          # :53
    g[0] = axis_value                                     # :54
    g = g - axis_value                                    # :56
    second = np.cumsum(g)                                 # :57
    return second - second[-1]                            # :58

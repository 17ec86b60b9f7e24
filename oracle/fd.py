"""Oracle restatement of the ``FiniteDifference`` stencils on the hot path --
TEST INFRASTRUCTURE ONLY.  turbopy/computetools.py:104-138,189-223."""
import numpy as np


def centered_difference(y, dr):
    """computetools.py:118-120: ``2*dr`` is formed first, then the divide; the
    two end points stay 0 (``Grid.generate_field`` zeros)."""
    y = np.asarray(y)
    d = np.zeros(len(y))
    d[1:-1] = (y[2:] - y[:-2]) / (2 * dr)
    return d


def upwind_left(y, widths):
    """computetools.py:136-138: divides by the per-cell ``Grid.cell_widths``
    array (core.py:670), not by a constant ``dr``; ``d[0]`` stays 0."""
    y = np.asarray(y)
    d = np.zeros(len(y))
    d[1:] = (y[1:] - y[:-1]) / widths
    return d


def del2_radial_rows(r, dr):
    """The three coefficient columns of ``del2_radial()`` (computetools.py:197-223)
    as (below, diag, above) with ``below[i] = A[i,i-1]``, ``above[i] = A[i,i+1]``.

    D = D1 + D2 (scipy adds the stored diagonals entry by entry):
    D1: ``A[i,i-1] = (-g1)/r[i]``, ``A[i,i+1] = g1/r[i]`` with ``A[0,1] = 0`` (:206);
    D2: ``g2, -2*g2, g2`` with ``A[0,1] = 2*g2`` (:217).
    """
    r = np.asarray(r, dtype=np.float64)
    n = len(r)
    g1 = 1 / (2.0 * dr)
    g2 = 1 / (dr ** 2)
    below = np.zeros(n)
    above = np.zeros(n)
    with np.errstate(divide="ignore"):
        below[1:] = (-g1) / r[1:] + g2
        above[:-1] = g1 / r[:-1] + g2
    above[0] = 0 + 2 * g2
    diag = np.full(n, -2 * g2)
    return below, diag, above


def del2_radial_apply(f, r, dr):
    """``del2_radial() @ f``.  scipy's dia matvec starts from zeros and adds
    one diagonal at a time in offset order [-1, 0, 1], so each row is
    ``((0 + A[i,i-1]*f[i-1]) + A[i,i]*f[i]) + A[i,i+1]*f[i+1]`` -- this order
    is kept (second differences cancel; reassociation shows at 1e-12)."""
    f = np.asarray(f, dtype=np.float64)
    below, diag, above = del2_radial_rows(r, dr)
    out = np.zeros(len(f))
    out[1:] = out[1:] + below[1:] * f[:-1]
    out = out + diag * f
    out[:-1] = out[:-1] + above[:-1] * f[1:]
    return out


# ---------------------------------------------------------------------------
# The remaining banded operators, turbopy/computetools.py:140-187,
# 225-381.  Each builder returns (data, offsets) in scipy's dia layout as the
# reference stores it: data[k][j] is the entry of column j on diagonal
# offsets[k] (entries that fall outside the matrix are stored but unused).
# ---------------------------------------------------------------------------
def banded_operator(name, r, dr):
    r = np.asarray(r, dtype=np.float64)
    n = len(r)
    zeros, ones = np.zeros(n), np.ones(n)
    if name == "ddx":                                   # :140-155
        g = 1 / (2.0 * dr)
        return np.array([zeros - g, zeros + g]), [-1, 1]
    if name == "ddr":                                   # :246-263
        g1 = 1 / (2.0 * dr)
        above = g1 * ones
        above[1] = 0
        return np.array([-g1 * ones, above]), [-1, 1]
    if name == "del2":                                  # :225-244
        g2 = 1 / (dr ** 2)
        above = g2 * ones
        above[1] = 2 * above[1]
        return np.array([g2 * ones, -2 * (g2 * ones), above]), [-1, 0, 1]
    if name == "radial_curl":                           # :157-187
        g = 1 / (2.0 * dr)
        below, diag, above = np.zeros(n), np.zeros(n), np.zeros(n)
        with np.errstate(divide="ignore", invalid="ignore"):
            below[:-1] = -g * (r[:-1] / r[1:])
            above[1:] = g * (r[1:] / r[:-1])
        above[1] = 2.0 / dr
        diag[-1] = 1.0 / dr
        below[-2] = 2.0 * below[-1]
        return np.array([below, diag, above]), [-1, 0, 1]
    if name in ("BC_left_extrap", "BC_left_avg", "BC_left_quad"):    # :265-335
        diag, above, above2 = np.ones(n), np.zeros(n), np.zeros(n)
        diag[0] = 0
        if name == "BC_left_extrap":
            above[1], above2[2] = 2, -1
        elif name == "BC_left_avg":
            above[1], above2[2] = 1.5, -0.5
        else:
            R2 = (r[1] ** 2 + r[2] ** 2) / (r[2] ** 2 - r[1] ** 2) / 2
            above[1], above2[2] = 0.5 + R2, 0.5 - R2
        return np.array([diag, above, above2]), [0, 1, 2]
    if name == "BC_left_flat":                          # :337-357
        diag, above = np.ones(n), np.zeros(n)
        diag[0] = 0
        above[1] = 1
        return np.array([diag, above]), [0, 1]
    if name == "BC_right_extrap":                       # :359-381
        diag, below, below2 = np.ones(n), np.zeros(n), np.zeros(n)
        diag[-1] = 0
        below[-2] = 2
        below2[-3] = -1
        return np.array([below2, below, diag]), [-2, -1, 0]
    raise KeyError(name)


BANDED_NAMES = ("ddx", "ddr", "del2", "radial_curl", "BC_left_extrap", "BC_left_avg", "BC_left_quad",
                "BC_left_flat", "BC_right_extrap")


def dia_apply(data, offsets, x):
    """``dia_matrix((data, offsets)) @ x`` the way scipy's dia_matvec sums it:
    start from zeros, add one stored diagonal at a time in the given order,
    ``y[i] += data[k][i+off] * x[i+off]``."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    y = np.zeros(n)
    for k, off in enumerate(offsets):
        j = np.arange(max(0, off), min(n, n + off))
        y[j - off] = y[j - off] + data[k][j] * x[j]
    return y


def dia_toarray(data, offsets, n):
    A = np.zeros((n, n))
    for k, off in enumerate(offsets):
        j = np.arange(max(0, off), min(n, n + off))
        A[j - off, j] = data[k][j]
    return A

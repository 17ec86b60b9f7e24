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

"""Oracle restatement of the linear field gather -- TEST INFRASTRUCTURE ONLY.

The reference has three "linear" gathers that round differently:

A  ``Grid.create_interpolator(r0)(y)``              turbopy/core.py:711-755
B  ``Interpolators.interpolate1D(x, y)(x_new)``, 1-D fp64 ``y``
       turbopy/computetools.py:459-481 -> scipy.interpolate.interp1d
       (scipy 1.18.1, not vendored, unpinned by the reference's setup.py:24-26)
       -> ``numpy.interp`` (numpy 2.3.5)
C  the same call with N-D ``y``  -> scipy ``interp1d._call_linear``

scipy/numpy are third-party code that is absent from /root/reference; their
published algorithms are restated here and pinned by comparing against the
installed scipy/numpy on seeded inputs (``make_golden.py``,
``tests/test_oracle_goldens.py``, ``tests/test_oracle_vs_reference.py``), and against the reference's own call sites
(tests/test_core.py:164-170, tests/test_computetools.py:71-79, the golden
tests/fixtures/particle_in_field/output/e_0.5.csv).
"""
import numpy as np

#: status codes shared with the CUDA kernels (include/turbopy_b200.h)
OK, BELOW, ABOVE, BAD_WINDOW = 0, 1, 2, 3


def cell_index(xq, xp):
    """Index j with ``xp[j] <= xq < xp[j+1]`` (numpy.interp's binary search);
    ``len(xp)-1`` for ``xq >= xp[-1]``, ``-1`` for ``xq < xp[0]``."""
    return np.searchsorted(xp, xq, side="right") - 1


def interp_b(xq, xp, fp):
    """Formula B: numpy.interp's arithmetic for in-range points.

    ``slope = (fp[j+1]-fp[j])/(xp[j+1]-xp[j]);  out = slope*(xq-xp[j]) + fp[j]``;
    exact ``fp[j]`` when ``xq == xp[j]`` and ``fp[-1]`` when ``xq == xp[-1]``;
    if that is NaN (an infinite ``fp``) the other end is tried, then the
    ``fp[j]==fp[j+1]`` shortcut.  Returns ``(values, status)``; out-of-range
    points (interp1d raises ValueError for them, bounds_error=True) get NaN and
    status BELOW / ABOVE.
    """
    xq = np.asarray(xq, dtype=np.float64)
    xp = np.asarray(xp, dtype=np.float64)
    fp = np.asarray(fp, dtype=np.float64)
    n = len(xp)
    status = np.zeros(xq.shape, dtype=np.int32)
    status[xq < xp[0]] = BELOW
    status[xq > xp[-1]] = ABOVE
    j = np.clip(cell_index(xq, xp), 0, n - 1)
    jn = np.minimum(j + 1, n - 1)
    with np.errstate(all="ignore"):
        slope = (fp[jn] - fp[j]) / (xp[jn] - xp[j])
        out = slope * (xq - xp[j]) + fp[j]
        bad = np.isnan(out)
        if bad.any():
            alt = slope * (xq - xp[jn]) + fp[jn]
            alt = np.where(np.isnan(alt) & (fp[j] == fp[jn]), fp[j], alt)
            out = np.where(bad, alt, out)
    out = np.where((j == n - 1) | (xp[j] == xq), fp[j], out)
    out = np.where(np.isnan(xq), xq, out)               # numpy.interp hands a NaN abscissa straight back
    out = np.where(status != OK, np.nan, out)
    return out, status


def interp_c(xq, xp, fp):
    """Formula C: scipy ``interp1d._call_linear`` (N-D ``y`` path).

    ``i = clip(searchsorted(xp, xq, 'left'), 1, n-1); lo=i-1; hi=i``
    ``out = ((xq-x_lo)/(x_hi-x_lo))*y_hi + ((x_hi-xq)/(x_hi-x_lo))*y_lo``.
    ``fp`` is ``(n,)`` or ``(k, n)`` (interpolation along the last axis, the
    interp1d default axis=-1); returns ``(values, status)`` with values shaped
    ``xq.shape`` or ``(k,) + xq.shape``.
    """
    xq = np.asarray(xq, dtype=np.float64)
    xp = np.asarray(xp, dtype=np.float64)
    fp = np.asarray(fp, dtype=np.float64)
    n = len(xp)
    status = np.zeros(xq.shape, dtype=np.int32)
    status[xq < xp[0]] = BELOW
    status[xq > xp[-1]] = ABOVE
    hi = np.clip(np.searchsorted(xp, xq, side="left"), 1, n - 1)
    lo = hi - 1
    x_lo, x_hi = xp[lo], xp[hi]
    with np.errstate(all="ignore"):
        w_hi = (xq - x_lo) / (x_hi - x_lo)
        w_lo = (x_hi - xq) / (x_hi - x_lo)
        out = w_hi * fp[..., hi] + w_lo * fp[..., lo]
    out = np.where(status != OK, np.nan, out)
    return out, status


def interp_a(xq, r, dr, fp, r_min=None, r_max=None):
    """Formula A: ``Grid.create_interpolator`` evaluated at many points.

    core.py:728 selects the nodes with ``r0 - dr < r[i] < r0 + dr`` (strict,
    with ``r0 -/+ dr`` rounded once).  One node -> ``fp[i]`` (:731-732); two
    nodes -> ``y0 + ((r0-r_lo)*(y1-y0))/(r_hi-r_lo)`` (:752-753); the asserts at
    :726-727,:729 become status BELOW / ABOVE / BAD_WINDOW with NaN values.
    """
    xq = np.asarray(xq, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    fp = np.asarray(fp, dtype=np.float64)
    n = len(r)
    r_min = r[0] if r_min is None else r_min
    r_max = r[-1] if r_max is None else r_max
    status = np.zeros(xq.shape, dtype=np.int32)
    lo_edge = xq - dr
    hi_edge = xq + dr
    # first node strictly above lo_edge, first node >= hi_edge
    first = np.searchsorted(r, lo_edge, side="right")
    last = np.searchsorted(r, hi_edge, side="left")       # exclusive
    count = last - first
    i0 = np.clip(first, 0, n - 1)
    i1 = np.clip(first + 1, 0, n - 1)
    with np.errstate(all="ignore"):
        y0, y1 = fp[i0], fp[i1]
        two = y0 + ((xq - r[i0]) * (y1 - y0)) / (r[i1] - r[i0])
    out = np.where(count == 1, fp[i0], two)
    status[(count != 1) & (count != 2)] = BAD_WINDOW
    status[xq > r_max] = ABOVE
    status[xq < r_min] = BELOW
    out = np.where(status != OK, np.nan, out)
    return out, status

"""Oracle for the kinetic-energy / momentum diagnostic -- TEST INFRASTRUCTURE
ONLY.

PARITY UNPINNED: the reference has no energy or moment diagnostic (nothing in
turbopy/diagnostics.py or tests/ computes one); BASELINE.json's north_star asks
for it as a new reduction.  The definition below uses the same ``gamma*m`` that
``BorisPush.push`` forms at turbopy/computetools.py:430-431:

    m1  = sqrt(mass**2 + |p|^2 / c2)
    KE  = sum_i |p_i|^2 / (m1_i + mass)        [ = (m1 - mass) * c2, cancellation-free ]
    P   = sum_i p_i  (three components),  N = particle count

The sums are order-dependent, so the CUDA result is compared with a relative
tolerance of 1e-12, stated in the tests, not bit-for-bit.
"""
import math

import numpy as np


def moments(px, py, pz, mass, c2):
    """Returns ``[KE, Px, Py, Pz, sum|p|^2, N, 0, 0]`` (the 8 doubles the CUDA
    reduction writes), accumulated with ``math.fsum`` (exactly rounded)."""
    p2 = (px * px + py * py) + pz * pz
    m1 = np.sqrt(mass ** 2 + p2 / c2)
    ke = p2 / (m1 + mass)
    return np.array([math.fsum(ke), math.fsum(px), math.fsum(py),
                     math.fsum(pz), math.fsum(p2), float(len(px)), 0.0, 0.0])

"""Shared test scaffolding: a stand-in `owner` exposing exactly the attributes
the tools read from a turboPy Simulation (owner.clock.dt, owner.grid.{r, dr,
cell_widths, num_points, r_min, r_max}), built from the oracle's Grid
restatement.  The oracle is the checker only; it is never the code under test."""
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import grid as ogrid  # noqa: E402

Q_E, M_E = -1.6022e-19, 9.1094e-31
Q_P, M_P = 1.6022e-19, 1.6726e-27
C_LIGHT = 2.9979e8


def make_grid(n, r_min=0.0, r_max=1.0):
    r = ogrid.grid_points(r_min, r_max, n)
    return SimpleNamespace(r=r, dr=ogrid.grid_dr(r_min, r_max, n), num_points=n, r_min=r_min, r_max=r_max,
                           cell_widths=ogrid.cell_widths(r))


def make_owner(dt=1e-12, grid=None, num_steps=10):
    return SimpleNamespace(clock=SimpleNamespace(dt=dt, num_steps=num_steps, time=0.0, this_step=0),
                           grid=grid if grid is not None else make_grid(10))


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = a.view(np.uint64) == b.view(np.uint64)
    if not same.all():
        idx = np.argwhere(~same)
        first = tuple(idx[0])
        raise AssertionError(f"{what}: {len(idx)} of {a.size} values differ in their bits; first at {first}: "
                             f"{a[first]!r} vs {b[first]!r}")


def thermal_ensemble(n, seed=1234, sigma=0.1):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0.0, 1.0, (n, 3))
    mom = M_E * C_LIGHT * rng.normal(0.0, sigma, (n, 3))
    return pos, mom


def beam_ensemble(n, seed=4567):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0.0, 1.0, (n, 3))
    g = np.exp(rng.uniform(0.0, np.log(1e3), n))
    bg = np.sqrt(g * g - 1.0)
    mom = np.zeros((n, 3))
    mom[:, 0] = bg * M_E * C_LIGHT
    mom[:, 1:] = 0.01 * mom[:, :1] * rng.normal(0.0, 1.0, (n, 2))
    return pos, mom

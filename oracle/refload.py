"""Import the UNMODIFIED reference from /root/reference -- TEST INFRASTRUCTURE
ONLY, and only usable in the build container (the GPU box has no
/root/reference; nothing that runs there may call this).

The three shims in ``oracle/shims`` (qtoml -> tomllib, np.int/np.float aliases,
a minimal xarray.DataArray) are the only environment changes; no reference
source is modified or copied.
"""
import os
import sys

REFERENCE = os.environ.get("TURBOPY_REFERENCE", "/root/reference")
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available():
    return os.path.isdir(os.path.join(REFERENCE, "turbopy"))


def load():
    """Returns the imported reference ``turbopy`` package."""
    if not available():
        raise ImportError(f"reference not present at {REFERENCE}")
    sys.dont_write_bytecode = True          # /root/reference is read-only
    for p in (REFERENCE, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int                         # turbopy/core.py:531,633
    if not hasattr(np, "float"):
        np.float = float                     # tests/test_computetools.py:29-30
    import turbopy
    return turbopy


def make_owner(turbopy, grid=None, clock=None):
    """A prepared reference ``Simulation`` with the given Grid / Clock dicts."""
    sim = turbopy.Simulation({
        "Grid": grid or {"N": 10, "min": 0.0, "max": 1.0},
        "Clock": clock or {"start_time": 0.0, "end_time": 1.0, "num_steps": 10},
        "PhysicsModules": {}, "Tools": {}, "Diagnostics": {},
    })
    sim.read_grid_from_input()
    sim.read_clock_from_input()
    return sim

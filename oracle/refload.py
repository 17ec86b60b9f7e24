"""Import the UNMODIFIED reference -- TEST INFRASTRUCTURE ONLY.

Where it is looked for, in order:
1. ``$TURBOPY_REFERENCE`` if set;
2. ``/root/reference`` (the build container);
3. ``oracle/_ref`` -- the byte-for-byte copy ``oracle/stage_reference.py`` makes in the build container; it is
   git-ignored but travels to the GPU box with the snapshot, so the oracle-vs-reference tests, the drop-in
   ``Simulation.run()`` tests and ``bench.py``'s CPU arms run the real reference there as well.

The three shims in ``oracle/shims`` (qtoml -> tomllib, np.int/np.float aliases, a minimal xarray.DataArray) are the
only environment changes; no reference source is modified.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU arms may import this module; the product never does.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(HERE, "shims")
STAGED = os.path.join(HERE, "_ref")


def _candidates():
    env = os.environ.get("TURBOPY_REFERENCE")
    if env:
        nested = os.path.join(env, "tests", "fixtures")
        yield env, nested if os.path.isdir(nested) else os.path.join(env, "fixtures")
    yield "/root/reference", "/root/reference/tests/fixtures"
    yield STAGED, os.path.join(STAGED, "fixtures")


def locate():
    """(directory holding the ``turbopy`` package, directory holding ``particle_in_field/``) or (None, None)."""
    for root, fixtures in _candidates():
        if os.path.isfile(os.path.join(root, "turbopy", "core.py")):
            return root, fixtures
    return None, None


REFERENCE, FIXTURES = locate()


def available():
    return REFERENCE is not None


def kind():
    """Where the reference comes from: 'source' (/root/reference), 'staged' (oracle/_ref) or None."""
    if REFERENCE is None:
        return None
    return "staged" if os.path.abspath(REFERENCE) == os.path.abspath(STAGED) else "source"


def tests_dir():
    """The reference's own ``tests/`` directory (conftest, test modules, fixtures), or None: ``<reference>/tests``
    in the build container, ``oracle/_ref/tests`` on the GPU box (staged by oracle/stage_reference.py)."""
    if REFERENCE is None:
        return None
    d = os.path.join(REFERENCE, "tests")
    return d if os.path.isfile(os.path.join(d, "test_computetools.py")) else None


def paths():
    """sys.path entries (shims first) a subprocess needs to import the reference."""
    return [SHIMS, REFERENCE]


def load():
    """Returns the imported reference ``turbopy`` package."""
    if not available():
        raise ImportError("reference not present (/root/reference or oracle/_ref; run python -m oracle.stage_reference)")
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


def load_pif():
    """The reference's particle-in-field app module (tests/fixtures/particle_in_field/particle_in_field.py),
    imported unmodified (its ``import xarray`` / ``import pytest`` resolve to the shim / the installed pytest)."""
    load()
    import importlib.util
    path = os.path.join(FIXTURES, "particle_in_field", "particle_in_field.py")
    spec = importlib.util.spec_from_file_location("reference_pif", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


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

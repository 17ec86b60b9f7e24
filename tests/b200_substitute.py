"""pytest plugin (``-p b200_substitute``) for running the reference's OWN test-suite against the B200 tools --
TEST INFRASTRUCTURE ONLY, loaded by tests/test_gpu_reference_suite.py in a subprocess.

The reference's tests construct the tools directly (``from turbopy.computetools import *`` and then
``BorisPush(Simulation(...), {...})``, tests/test_computetools.py:4,9-10), so overriding the registry alone would
leave them on the stock classes.  The plugin is imported before pytest collects anything: it imports the unmodified
reference, imports ``turbopy_b200`` (which takes the registry names, turbopy/core.py:302-312), and rebinds the four
tool names in ``turbopy.computetools`` and ``turbopy`` to the B200 classes, so the star-imports of the reference's
test modules pick them up.  No reference file is changed.  At session end it writes how many kernels the C-ABI
library launched and which class each name resolved to, for the calling test to assert on.
"""
import json
import os

import numpy as np

if not hasattr(np, "int"):
    np.int = int                                  # turbopy/core.py:531,633
if not hasattr(np, "float"):
    np.float = float                              # tests/test_computetools.py:29-30

import turbopy                                    # noqa: E402  the unmodified reference (PYTHONPATH set by the caller)
import turbopy.computetools as _ct                # noqa: E402

SUBSTITUTE = os.environ.get("B200_SUBSTITUTE", "1") != "0"
NAMES = ("BorisPush", "Interpolators", "FiniteDifference", "PoissonSolver1DRadial")
_tp = None
if SUBSTITUTE:
    import turbopy_b200 as _tp                    # noqa: E402  registers over the stock names
    for _name in NAMES:
        setattr(_ct, _name, getattr(_tp, _name))
        setattr(turbopy, _name, getattr(_tp, _name))


def pytest_sessionfinish(session, exitstatus):
    report = {"substituted": SUBSTITUTE, "exitstatus": int(exitstatus),
              "classes": {n: f"{getattr(_ct, n).__module__}.{getattr(_ct, n).__qualname__}" for n in NAMES},
              "registry": {n: f"{turbopy.ComputeTool.lookup(n).__module__}.{turbopy.ComputeTool.lookup(n).__qualname__}"
                           for n in NAMES},
              "launches": None}
    if _tp is not None:
        from turbopy_b200 import _lib
        report["launches"] = int(_lib.launch_count())
    out = os.environ.get("B200_SUBSTITUTE_REPORT")
    if out:
        with open(out, "w") as f:
            json.dump(report, f)

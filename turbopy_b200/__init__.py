"""turbopy_b200 -- B200-native implementation of turboPy's particle hot path.

``BorisPush.push`` fused with the linear 1-D field gather, the
``FiniteDifference`` stencils and the energy / moment reductions, as hand-written
sm_100a CUDA kernels behind a C ABI (include/turbopy_b200.h), exposed as drop-in
``ComputeTool``s under the reference's own registry names.

    import turbopy, turbopy_b200          # the second import overrides the stock tools
    sim = turbopy.Simulation(config)      # "BorisPush", "Interpolators", "FiniteDifference"
                                          # in config["Tools"] now resolve to the B200 tools

There is no CPU fallback: calls raise if libturbopy_b200.so or a B200 is missing.
"""
import os

from . import _lib
from ._host import ComputeTool, Diagnostic, HAVE_TURBOPY, stock_tools
from .diagnostics import (DeviceFieldDiagnostic, DeviceParticleDiagnostic, DevicePointDiagnostic,
                          KineticEnergyDiagnostic, reduce_moments)
from .cycle import DeviceClock, PlaneWave
from .graph import StepGraph
from .particles import resort_local, soa_particles, sort_by_cell, to_soa
from .tools import (BandedOperator, BorisPush, FiniteDifference, GridOnDevice, Interpolators,
                    PoissonSolver1DRadial)

__version__ = "0.1.0"

TOOL_CLASSES = {"BorisPush": BorisPush, "Interpolators": Interpolators, "FiniteDifference": FiniteDifference,
                "PoissonSolver1DRadial": PoissonSolver1DRadial}
DIAGNOSTIC_CLASSES = {"kinetic_energy": KineticEnergyDiagnostic, "device_point": DevicePointDiagnostic,
                      "device_field": DeviceFieldDiagnostic, "device_particle": DeviceParticleDiagnostic}


def register_tools(override=True):
    """Register the B200 tools under the stock names.  The names are already
    taken once ``turbopy`` is imported (turbopy/computetools.py:484-487), hence
    ``override=True`` (turbopy/core.py:302-312).  The displaced stock classes are
    remembered in ``stock_tools``: calls outside the GPU path's scope (spline
    kinds of ``interpolate1D``) are handed back to them, so a Simulation that
    worked before the import keeps working after it."""
    if not HAVE_TURBOPY:
        raise RuntimeError("turboPy is not importable: there is no ComputeTool registry to register with")
    for name, cls in TOOL_CLASSES.items():
        if ComputeTool.is_valid_name(name) and ComputeTool.lookup(name) is not cls:
            stock_tools.setdefault(name, ComputeTool.lookup(name))
        ComputeTool.register(name, cls, override=override)
    for name, cls in DIAGNOSTIC_CLASSES.items():
        Diagnostic.register(name, cls, override=override)


def unregister_tools():
    """Put the stock tools back under their names (the new diagnostics keep theirs)."""
    for name, cls in stock_tools.items():
        ComputeTool.register(name, cls, override=True)


if HAVE_TURBOPY and not os.environ.get("TURBOPY_B200_NO_AUTOREGISTER"):
    register_tools(override=True)

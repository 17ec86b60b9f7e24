"""The reference's plugin base classes.

turbopy_b200 is a plugin for turboPy: with turboPy importable (the deployment:
``import turbopy; import turbopy_b200``) the tools subclass the reference's own
``turbopy.core.ComputeTool`` / ``Diagnostic``, so ``ComputeTool.register(...,
override=True)`` accepts them (turbopy/core.py:302-312 checks ``issubclass``) and
the registry, ``Simulation``, ``Grid`` and ``Clock`` are the reference's, untouched.

Without turboPy on ``sys.path`` there is no registry to plug into; the tool
classes can still be constructed and called directly (kernel benchmarks do that),
on two bare bases that only keep the constructor contract
``cls(owner=..., input_data=...)`` (core.py:456-459, :837-839).  Nothing of the
reference's control plane is rebuilt here.
"""

#: the stock tool classes the B200 tools displaced in the registry (name -> class), filled by register_tools();
#: the B200 tools hand calls that are outside the GPU path's scope (spline interpolation kinds) back to them
stock_tools = {}

try:                                        # pragma: no cover - depends on environment
    from turbopy.core import ComputeTool, Diagnostic
    HAVE_TURBOPY = True
except Exception:                           # turboPy (or one of its dependencies) is absent
    HAVE_TURBOPY = False

    class ComputeTool:
        def __init__(self, owner, input_data):
            self.owner = owner
            self.input_data = input_data
            self.name = input_data["type"]

        def initialize(self):
            pass

    class Diagnostic:
        def __init__(self, owner, input_data):
            self.owner = owner
            self.input_data = input_data

        def inspect_resource(self, resource):
            pass

        def initialize(self):
            pass

        def finalize(self):
            pass

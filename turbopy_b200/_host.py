"""The reference's plugin base classes, or a mirror of them.

When turboPy is importable (the normal deployment: `import turbopy; import
turbopy_b200`) the tools below subclass the *reference's own*
``turbopy.core.ComputeTool`` / ``Diagnostic`` so that
``ComputeTool.register(..., override=True)`` accepts them
(turbopy/core.py:302-312 checks ``issubclass``).

When it is not (the GPU test box has no /root/reference), a minimal mirror of
the same interface is used so the tools can still be constructed and driven by
the parity tests: same ``register`` / ``lookup`` / ``is_valid_name`` semantics
and error types (turbopy/core.py:282-327), same constructor
``cls(owner=..., input_data=...)`` storing ``owner``, ``input_data``,
``name = input_data["type"]`` (:456-459), same no-op ``initialize`` (:461-463).
Nothing of the reference's control plane (Simulation, Grid, Clock) is rebuilt.
"""

try:                                        # pragma: no cover - depends on environment
    from turbopy.core import ComputeTool, Diagnostic
    HAVE_TURBOPY = True
except Exception:                           # turboPy (or one of its deps) is absent
    HAVE_TURBOPY = False

    class _Factory:
        """Mirror of turbopy.core.DynamicFactory (core.py:282-327)."""
        _factory_type_name = "Factory"
        _registry = None

        @classmethod
        def register(cls, name_to_register, class_to_register, override=False):
            if name_to_register in cls._registry and not override:
                raise ValueError("{0} '{1}' already registered".format(
                    cls._factory_type_name, name_to_register))
            if not issubclass(class_to_register, cls):
                raise TypeError("{0} is not a subclass of {1}".format(class_to_register, cls))
            cls._registry[name_to_register] = class_to_register

        @classmethod
        def lookup(cls, name):
            try:
                return cls._registry[name]
            except KeyError:
                raise KeyError("{0} '{1}' not found in registry".format(cls._factory_type_name, name))

        @classmethod
        def is_valid_name(cls, name):
            return name in cls._registry

    class ComputeTool(_Factory):
        """Mirror of turbopy.core.ComputeTool (core.py:424-463)."""
        _factory_type_name = "Compute Tool"
        _registry = {}

        def __init__(self, owner, input_data):
            self.owner = owner
            self.input_data = input_data
            self.name = input_data["type"]

        def initialize(self):
            pass

    class Diagnostic(_Factory):
        """Mirror of turbopy.core.Diagnostic (core.py:811-884)."""
        _factory_type_name = "Diagnostic"
        _registry = {}

        def __init__(self, owner, input_data):
            self.owner = owner
            self.input_data = input_data

        def inspect_resource(self, resource):
            pass

        def diagnose(self):
            raise NotImplementedError

        def initialize(self):
            pass

        def finalize(self):
            pass

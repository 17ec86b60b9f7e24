"""TEST INFRASTRUCTURE ONLY -- the few lines of `xarray.DataArray` that the
reference's particle_in_field fixture touches
(tests/fixtures/particle_in_field/particle_in_field.py:18-23): an ndarray that
carries `.attrs` and exposes its coordinates as attributes."""
import numpy as np


class _Coord(np.ndarray):
    def __new__(cls, data):
        obj = np.asarray(data).view(cls)
        obj.attrs = {}
        return obj

    def __array_finalize__(self, obj):
        self.attrs = getattr(obj, "attrs", {})


class DataArray(np.ndarray):
    def __new__(cls, data, dims=None, coords=None):
        obj = np.asarray(data).view(cls)
        obj.attrs = {}
        obj.dims = dims
        obj._coords = {k: _Coord(v) for k, v in (coords or {}).items()}
        return obj

    def __array_finalize__(self, obj):
        self.attrs = getattr(obj, "attrs", {})
        self.dims = getattr(obj, "dims", None)
        self._coords = getattr(obj, "_coords", {})

    def __getattr__(self, name):
        coords = self.__dict__.get("_coords", {})
        if name in coords:
            return coords[name]
        raise AttributeError(name)

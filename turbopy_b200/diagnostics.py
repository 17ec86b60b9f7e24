"""Energy / moment diagnostics on device-resident particles (kernel 3).

The reference has no energy diagnostic (turbopy/diagnostics.py holds only
point / field / grid / clock diagnostics), so this is a NEW ``Diagnostic``
registered through ``Diagnostic.register`` (turbopy/core.py:835-836) that
follows the stock plumbing: ``inspect_resource`` grabs the shared momentum
array, ``diagnose()`` runs every step, ``finalize()`` writes a CSV the way
``CSVOutputUtility`` does (turbopy/diagnostics.py:14-59: one row per diagnose
call, ``np.savetxt(..., delimiter=",")``).

Per diagnose: one deterministic fp64 reduction (tp_reduce_moments_f64) into a
row of a device buffer and, when torch.distributed is initialised, one
all-reduce (NCCL) of that 8-double row -- no host sync until finalize().
"""
import ctypes as C
import os
import weakref

import numpy as np
import torch

from . import _lib
from ._host import Diagnostic
from .sharding import allreduce_sum_

MOMENT_NAMES = ("kinetic_energy", "px", "py", "pz", "p_squared", "count")


def reduce_moments(momentum, mass, c2, out=None, scratch=None):
    """``[KE, Px, Py, Pz, sum|p|^2, n, 0, 0]`` of an (n,3) CUDA momentum tensor,
    with KE = sum |p|^2/(m1+mass), m1 = sqrt(mass**2 + |p|^2/c2) (the gamma*m
    of turbopy/computetools.py:430-431).  Asynchronous; returns the 8-double
    device tensor."""
    if not (isinstance(momentum, torch.Tensor) and momentum.is_cuda and momentum.dtype == torch.float64):
        raise TypeError("momentum must be a CUDA float64 tensor")
    if momentum.ndim != 2 or momentum.shape[1] != 3:
        raise ValueError("momentum must have shape (n,3)")
    dev = momentum.device
    lib = _lib.lib()
    if out is None:
        out = torch.empty(8, dtype=torch.float64, device=dev)
    if scratch is None:
        scratch = torch.empty(int(lib.tp_reduce_scratch_doubles()), dtype=torch.float64, device=dev)
    sn, sc = momentum.stride()
    if momentum.shape[0] <= 1:
        sn = max(sn, 1)
    with torch.cuda.device(dev):
        _lib.check(lib.tp_reduce_moments_f64(
            C.c_void_p(momentum.data_ptr()), sn, sc, momentum.shape[0],
            float(mass), float(mass ** 2), float(c2),
            C.c_void_p(out.data_ptr()), C.c_void_p(scratch.data_ptr()),
            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


#: Deferred moment requests, keyed by (device index, data_ptr) of the momentum tensor: a KineticEnergyDiagnostic with
#: ``fuse_with_push`` posts its row here instead of launching the reduction, and the next ``BorisPush.gather_push`` on
#: that momentum tensor takes it along (tp_gather_push_moments_f64: the sums are accumulated from the momenta
#: the push kernel has just loaded).  Anything that needs the row earlier flushes it with the standalone reduction.
_PENDING = weakref.WeakValueDictionary()      # (a diagnostic that is dropped un-finalized takes its request with it)


def _pending_key(momentum):
    return (momentum.device.index, momentum.data_ptr())


def _can_ride_on_push(momentum):
    """Only a resident CUDA tensor is pushed by the kernels that can deliver a request."""
    return isinstance(momentum, torch.Tensor) and momentum.is_cuda


def take_moments_request(momentum):
    """The diagnostic waiting for the moments of ``momentum`` (and forget the request), or None."""
    return _PENDING.pop(_pending_key(momentum), None)


def flush_moments_request(momentum):
    """A caller is about to change ``momentum`` through a path that cannot deliver a posted request (``push_steps``,
    multi-step passes): run the standalone reduction now, while the momenta are still the ones
    the diagnostic asked about."""
    diag = _PENDING.get(_pending_key(momentum))
    if diag is not None:
        diag._flush()


class KineticEnergyDiagnostic(Diagnostic):
    """``input_data``: ``"momentum"`` (resource name of the shared (n,3)
    momentum tensor), ``"mass"``, optional ``"c2"`` (default BorisPush's
    2.9979e8**2), ``"interval"`` (steps between reductions, default 1),
    ``"filename"`` / ``"directory"`` for the CSV, ``"allreduce"`` (default
    True: sum over ranks when torch.distributed is initialised),
    ``"fuse_with_push"`` (default False).

    ``fuse_with_push``: the Simulation loop runs ``diagnose()`` before the
    modules' ``update()`` (turbopy/core.py:152-158), so the momenta this
    diagnostic sums are the ones the step's push is about to load anyway.
    With the option set, ``diagnose()`` only posts a request; the next
    ``BorisPush.gather_push`` on the same momentum tensor delivers the row
    from inside its kernel (no second pass over the momentum arrays) and
    enqueues the all-reduce behind it.  Contract: nothing else writes the
    momentum between ``diagnose()`` and that push.  ``values()``,
    ``finalize()`` and the next ``diagnose()`` flush an undelivered request
    with the standalone reduction, so the rows are the same either way (to
    summation-order rounding, 1e-12)."""

    def __init__(self, owner, input_data):
        super().__init__(owner, input_data)
        self.momentum = None
        self.mass = float(input_data["mass"])
        self.c2 = float(input_data.get("c2", 2.9979e8 ** 2))
        self.interval = int(input_data.get("interval", 1))
        self.allreduce = bool(input_data.get("allreduce", True))
        self.fuse_with_push = bool(input_data.get("fuse_with_push", False))
        self.rows = None
        self._scratch = None
        self._calls = 0
        self._used = 0
        self._pending_row = None
        self.fused_rows = 0              # rows delivered by a push kernel (the rest came from the standalone reduction)

    def inspect_resource(self, resource):
        name = self.input_data["momentum"]
        if name in resource:
            self.momentum = resource[name]

    def initialize(self):
        steps = int(self.owner.clock.num_steps)
        self._capacity = steps // self.interval + 2

    # ---- deferred rows (fuse_with_push)
    def moments_request(self):
        """(mass, mass**2, c2, out8 row, scratch) of the posted request -- read by BorisPush.gather_push."""
        return self.mass, self.mass ** 2, self.c2, self._pending_row, self._scratch

    def moments_delivered(self):
        """BorisPush.gather_push has enqueued the kernels that fill the posted row (on the current stream)."""
        row, self._pending_row = self._pending_row, None
        self.fused_rows += 1
        if self.allreduce:
            allreduce_sum_(row)

    def _flush(self):
        """An undelivered request: the standalone reduction, now."""
        if self._pending_row is None:
            return
        _PENDING.pop(_pending_key(self.momentum), None)
        row, self._pending_row = self._pending_row, None
        reduce_moments(self.momentum, self.mass, self.c2, out=row, scratch=self._scratch)
        if self.allreduce:
            allreduce_sum_(row)

    def diagnose(self):
        self._calls += 1
        if (self._calls - 1) % self.interval:
            return
        if self.momentum is None:
            raise RuntimeError("Diagnostic field {} not found".format(self.input_data["momentum"]))
        self._flush()
        if self.rows is None:
            dev = self.momentum.device
            cap = getattr(self, "_capacity", 1024)
            self.rows = torch.zeros((cap, 8), dtype=torch.float64, device=dev)
        if self._used >= self.rows.shape[0]:
            self.rows = torch.cat([self.rows, torch.zeros_like(self.rows)])
        row = self.rows[self._used]
        if self._scratch is None:
            self._scratch = torch.empty(int(_lib.lib().tp_reduce_scratch_doubles()), dtype=torch.float64,
                                        device=self.momentum.device)
        self._used += 1
        if self.fuse_with_push and _can_ride_on_push(self.momentum):
            self._pending_row = row                      # filled by the next gather_push on this tensor (or a flush)
            _PENDING[_pending_key(self.momentum)] = self
            return
        reduce_moments(self.momentum, self.mass, self.c2, out=row, scratch=self._scratch)
        if self.allreduce:
            allreduce_sum_(row)

    def values(self):
        """Host copy of the rows recorded so far: (k, 6) [KE, Px, Py, Pz, |p|^2, n]."""
        if self.rows is None:
            return np.zeros((0, 6))
        self._flush()
        return self.rows[:self._used, :6].cpu().numpy()

    def finalize(self):
        self.diagnose()
        self._flush()
        filename = self.input_data.get("filename")
        if filename:
            directory = self.input_data.get("directory", "")
            if directory:
                os.makedirs(directory, exist_ok=True)
                filename = os.path.join(directory, filename)
            with open(filename, "wb") as f:
                np.savetxt(f, self.values(), delimiter=",")


# --------------------------------------------------------------------------- SURVEY 8f-3
# Device-aware versions of the stock point / field diagnostics
# (turbopy/diagnostics.py:61-303).  They read CUDA field tensors, keep their
# rows in a device buffer (no host sync per step) and write, at finalize(), the
# same CSV bytes the stock diagnostics write for the same data
# (CSVOutputUtility: np.savetxt(f, buffer, delimiter=","), diagnostics.py:55-59).
# Registered under NEW names ("device_point", "device_field"); the stock "point"
# and "field" stay untouched and keep serving host numpy fields.
def _savetxt(filename, rows):
    with open(filename, "wb") as f:
        np.savetxt(f, rows, delimiter=",")


class DevicePointDiagnostic(Diagnostic):
    """``input_data``: ``"field"`` (resource name of a (ng,) CUDA field),
    ``"location"``, ``"output_type"`` ("csv" | "stdout"), ``"filename"``.  The
    value is ``Grid.create_interpolator(location)(field)`` (formula A,
    turbopy/core.py:711-755) evaluated by tp_interp1d_f64 straight into the
    row buffer."""

    def __init__(self, owner, input_data):
        super().__init__(owner, input_data)
        self.location = input_data["location"]
        self.field_name = input_data["field"]
        self.output = input_data["output_type"]
        self.field = None
        self.rows = None
        self._used = 0

    def inspect_resource(self, resource):
        if self.field_name in resource:
            self.field = resource[self.field_name]

    def initialize(self):
        from .tools import GatherStatus, GridOnDevice
        if self.field is None:
            raise RuntimeError(f"Diagnostic field {self.field_name} was not found")
        grid = self.owner.grid
        assert (self.location >= grid.r_min), "Requested point is not in the grid"       # core.py:726
        assert (self.location <= grid.r_max), "Requested point is not in the grid"       # core.py:727
        dev = self.field.device
        self._grid = GridOnDevice.of(grid, dev)
        self._status = GatherStatus(dev)
        self._xq = torch.tensor([float(self.location)], dtype=torch.float64, device=dev)
        self.rows = torch.zeros((int(self.owner.clock.num_steps) + 1, 1), dtype=torch.float64, device=dev)

    def diagnose(self):
        f = self.field
        if f.ndim != 1 or not f.is_cuda or f.dtype != torch.float64:
            raise TypeError("device_point needs a (ng,) CUDA float64 field")
        if self._used >= self.rows.shape[0]:
            self.rows = torch.cat([self.rows, torch.zeros_like(self.rows)])
        g = self._grid
        row = self.rows[self._used]
        with torch.cuda.device(f.device):
            _lib.check(_lib.lib().tp_interp1d_f64(
                C.c_void_p(g.r.data_ptr()), g.ng, g.dr, g.r_first, g.r_last,
                C.c_void_p(f.data_ptr()), 1, g.ng, f.stride(0),
                C.c_void_p(self._xq.data_ptr()), 1, 1,
                C.c_void_p(row.data_ptr()), 1, _lib.TP_INTERP_A,
                C.c_void_p(self._status.words.data_ptr()),
                C.c_void_p(torch.cuda.current_stream(f.device).cuda_stream)))
        self._used += 1
        if self.output == "stdout":
            print(row.cpu().numpy())

    def finalize(self):
        self.diagnose()
        count, _, _ = self._status.read()
        assert count == 0, "Error finding requested point in the grid"                   # core.py:729
        if self.output == "csv":
            _savetxt(self.input_data["filename"], self.rows.cpu().numpy())


class DeviceFieldDiagnostic(Diagnostic):
    """``input_data``: ``"field"``, ``"component"``, ``"output_type"``,
    ``"filename"``, optional ``"dump_interval"`` -- the stock FieldDiagnostic's
    keys and timing (turbopy/diagnostics.py:164-303) for a CUDA field of shape
    (ng,) or (ng, k)."""

    def __init__(self, owner, input_data):
        super().__init__(owner, input_data)
        self.component = input_data["component"]
        self.field_name = input_data["field"]
        self.output = input_data["output_type"]
        self.field = None
        self.dump_interval = None
        self.last_dump = None
        self.diagnose = self.do_diagnostic
        self.diagnostic_size = None
        self.field_was_found = False
        self.rows = None
        self._used = 0

    def inspect_resource(self, resource):
        if self.field_name in resource:
            self.field_was_found = True
            self.field = resource[self.field_name]

    def check_step(self):
        if self.owner.clock.time >= self.last_dump + self.dump_interval:
            self.do_diagnostic()
            self.last_dump = self.owner.clock.time

    def do_diagnostic(self):
        f = self.field
        data = f[:, self.component] if f.ndim > 1 else f
        if self.output == "stdout":
            print(self.field_name, data.cpu().numpy())
            return
        if self._used < self.rows.shape[0]:
            self.rows[self._used].copy_(data)               # device-to-device, asynchronous
        self._used += 1

    def initialize(self):
        if not self.field_was_found:
            raise RuntimeError(f"Diagnostic field {self.field_name} was not found")
        self.diagnostic_size = (int(self.owner.clock.num_steps) + 1, int(self.field.shape[0]))
        if "dump_interval" in self.input_data:
            self.dump_interval = self.input_data["dump_interval"]
            self.diagnose = self.check_step
            self.last_dump = 0
            self.diagnostic_size = (int(np.ceil(self.owner.clock.end_time / self.dump_interval) + 1),
                                    int(self.field.shape[0]))
        if self.output == "csv":
            self.rows = torch.zeros(self.diagnostic_size, dtype=torch.float64, device=self.field.device)

    def finalize(self):
        self.do_diagnostic()
        if self.output == "csv":
            if self._used > self.rows.shape[0]:
                # CSVOutputUtility.append would have raised IndexError on the extra row
                raise IndexError(f"index {self.rows.shape[0]} is out of bounds for axis 0 with size {self.rows.shape[0]}")
            _savetxt(self.input_data["filename"], self.rows.cpu().numpy())


class DeviceParticleDiagnostic(Diagnostic):
    """One particle's row of a shared (n,3) CUDA tensor, every step -- the
    device-side counterpart of the particle-in-field example's
    ``ParticleDiagnostic`` (tests/fixtures/particle_in_field/particle_in_field.py:66-99:
    ``output_function(self.data[0, :])`` into a ``(num_steps + 1, 3)`` CSV).

    ``input_data``: ``"resource"`` (name of the shared tensor, e.g.
    ``"ChargedParticle:position"``) or the example's ``"component"``
    (``"position"`` / ``"momentum"``, looked up as ``"ChargedParticle:<component>"``);
    ``"index"`` (which particle, default 0); ``"output_type"`` ("csv" | "stdout");
    ``"filename"``.  Rows are copied device-to-device on the current stream; the
    only host transfer is in ``finalize()``."""

    def __init__(self, owner, input_data):
        super().__init__(owner, input_data)
        self.resource_name = input_data.get("resource", "ChargedParticle:" + str(input_data.get("component", "position")))
        self.index = int(input_data.get("index", 0))
        self.output = input_data["output_type"]
        self.data = None
        self.rows = None
        self._used = 0

    def inspect_resource(self, resource):
        if self.resource_name in resource:
            self.data = resource[self.resource_name]

    def initialize(self):
        if self.data is None:
            raise RuntimeError(f"Diagnostic resource {self.resource_name} was not found")
        d = self.data
        if not (isinstance(d, torch.Tensor) and d.is_cuda and d.dtype == torch.float64 and d.ndim == 2):
            raise TypeError("device_particle needs an (n,k) CUDA float64 tensor")
        if not -d.shape[0] <= self.index < d.shape[0]:
            raise IndexError(f"index {self.index} is out of bounds for axis 0 with size {d.shape[0]}")
        if self.output == "csv":
            self.rows = torch.zeros((int(self.owner.clock.num_steps) + 1, d.shape[1]), dtype=torch.float64, device=d.device)

    def diagnose(self):
        row = self.data[self.index, :]
        if self.output == "stdout":
            print(row.cpu().numpy())
            return
        if self._used < self.rows.shape[0]:
            self.rows[self._used].copy_(row)                # device-to-device, asynchronous
        self._used += 1

    def finalize(self):
        self.diagnose()
        if self.output == "csv":
            if self._used > self.rows.shape[0]:
                # CSVOutputUtility.append would have raised IndexError on the extra row
                raise IndexError(f"index {self.rows.shape[0]} is out of bounds for axis 0 with size {self.rows.shape[0]}")
            _savetxt(self.input_data["filename"], self.rows.cpu().numpy())

"""Drop-in ComputeTools for turboPy's particle hot path, backed by the sm_100a
kernels in libturbopy_b200.so.

Same registry names, constructor and call signatures as the stock tools in
turbopy/computetools.py, so Simulation / PhysicsModule / Grid / Clock /
Diagnostic code is untouched:

* ``BorisPush.push(position, momentum, charge, mass, E, B)``   (computetools.py:407)
* ``Interpolators.interpolate1D(x, y, kind='linear')``           (computetools.py:459)
* ``FiniteDifference.setup_ddx / centered_difference / upwind_left / del2_radial``
                                                               (computetools.py:87-223)

plus one additive method, ``BorisPush.gather_push``, the fused field gather +
push of ``ChargedParticle.update`` (tests/fixtures/particle_in_field/
particle_in_field.py:61-63) for N particles.

Arrays: resident ``torch.float64`` CUDA tensors (any of the layouts the C ABI
accepts; SoA from :func:`turbopy_b200.particles.soa_particles` is the fast one)
are updated in place on the current CUDA stream, asynchronously.  numpy arrays
(what the reference's callers pass) take the end-to-end host path: they are
streamed through the GPU by the library and hold the result on return.  Either
way the arithmetic runs in the CUDA kernels; there is no CPU implementation in
this package, and calls raise if the library or a B200 is missing.
"""
import ctypes as C
import os
import weakref

import numpy as np
import torch

from . import _lib
from . import diagnostics as _diagnostics
from ._host import ComputeTool, stock_tools

_FORMULAS = {
    "A": _lib.TP_INTERP_A, "create_interpolator": _lib.TP_INTERP_A,
    "B": _lib.TP_INTERP_B, "interp": _lib.TP_INTERP_B, "linear": _lib.TP_INTERP_B,
    "C": _lib.TP_INTERP_C, "interp1d_nd": _lib.TP_INTERP_C,
}


# --------------------------------------------------------------------------- helpers
def _is_cuda(t):
    return isinstance(t, torch.Tensor) and t.is_cuda


try:                                        # raw handle of torch's current stream, without building a Stream object
    _raw_stream = torch._C._cuda_getCurrentRawStream
except AttributeError:                      # pragma: no cover - older / newer torch
    _raw_stream = None


def _device_index(device):
    return device.index if device.index is not None else torch.cuda.current_device()


def _stream_ptr(device):
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(_device_index(device)))
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _on_device:
    """``with torch.cuda.device(dev)`` that costs nothing when ``dev`` is already
    current (one process per GPU: always)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if _device_index(device) == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _require_f64(t, what):
    dt = t.dtype
    if dt not in (torch.float64, np.dtype(np.float64)):
        raise TypeError(f"{what} must be float64, got {dt}")


def _n3_strides(t, what):
    """(stride_n, stride_c) in elements of a logical (n,3) array."""
    if t.ndim != 2:
        # the reference fails the same way on (3,) particles (m1[:, np.newaxis])
        raise IndexError(f"{what} must be a 2-D (n,3) array, got shape {tuple(t.shape)}")
    if t.shape[1] != 3:
        raise ValueError(f"{what} must have shape (n,3), got {tuple(t.shape)}")
    _require_f64(t, what)
    if isinstance(t, torch.Tensor):
        sn, sc = t.stride()
    else:
        if t.strides[0] % 8 or t.strides[1] % 8:
            raise ValueError(f"{what}: strides must be multiples of 8 bytes")
        sn, sc = t.strides[0] // 8, t.strides[1] // 8
    if t.shape[0] <= 1:
        sn = max(sn, 1)
    if t.shape[0] == 0:                     # strides of an empty array carry no information
        sc = max(sc, 1)
    if sn <= 0 or sc <= 0:
        raise ValueError(f"{what}: negative or zero strides are not supported")
    return int(sn), int(sc)


def _ptr(t):
    if isinstance(t, torch.Tensor):
        return t.data_ptr()
    return t.ctypes.data


# ---- page-locking of caller-owned numpy arrays (host path) -------------------
# A Simulation passes the same position / momentum arrays every step.  Pageable
# memory is copied through the driver's staging buffers (0.14 G pushes/s end to end
# at 16 Mi particles); page-locked once, the same arrays move by asynchronous DMA
# (0.93).  The registration lives exactly as long as the ndarray that OWNS the
# memory: a weakref finalizer unpins it before numpy frees the buffer, so a later
# allocation at the same address is never mistaken for the pinned one.
_PIN_MIN_BYTES = 1 << 20
_pinned = {}


def _unpin(key, ptr):
    _pinned.pop(key, None)
    try:
        _lib.lib().tp_host_unpin(C.c_void_p(ptr))
    except Exception:                       # interpreter shutdown: the process is going away anyway
        pass


def _pin_host(array):
    """Page-lock the buffer behind ``array`` the SECOND time it is passed in (page-locking costs ~4 ms per MB,
    which a one-off call should not pay; a time loop amortises it within a few steps).  No-op when disabled
    with TURBOPY_B200_PIN=0, for small arrays, foreign buffers, or if the driver refuses."""
    if os.environ.get("TURBOPY_B200_PIN", "1") == "0":
        return False
    owner = array
    while isinstance(owner.base, np.ndarray):
        owner = owner.base
    if owner.base is not None or not owner.flags.owndata or owner.nbytes < _PIN_MIN_BYTES:
        return False                        # memory owned by something else (mmap, another library): leave it alone
    key = id(owner)
    hit = _pinned.get(key)
    if hit is None:                         # first sight: only remember the owner (forgotten again when it dies)
        _pinned[key] = (weakref.finalize(owner, _pinned.pop, key, None), None)
        return False
    fin, state = hit
    if state is not None:
        return state
    ptr = owner.ctypes.data
    ok = _lib.lib().tp_host_pin(C.c_void_p(ptr), owner.nbytes) == 0
    if ok:
        fin.detach()
        fin = weakref.finalize(owner, _unpin, key, ptr)
    _pinned[key] = (fin, ok)
    return ok


_struct_cache = {}


def _particles_struct(position, momentum):
    """tp_particles for the two (n,3) arrays.  A Simulation passes the same two
    arrays every step, so the validated struct is cached on (pointers, shapes,
    strides, dtypes)."""
    if isinstance(position, torch.Tensor) and isinstance(momentum, torch.Tensor):
        key = (position.data_ptr(), momentum.data_ptr(), position.shape, momentum.shape,
               position.stride(), momentum.stride(), position.dtype, momentum.dtype)
        hit = _struct_cache.get(key)
        if hit is not None:
            return hit
    else:
        key = None
    psn, psc = _n3_strides(position, "position")
    msn, msc = _n3_strides(momentum, "momentum")
    if position.shape[0] != momentum.shape[0]:
        raise ValueError("position and momentum must have the same number of particles")
    out = _lib.Particles(_ptr(position), psn, psc, _ptr(momentum), msn, msc, int(position.shape[0]))
    if key is not None:
        if len(_struct_cache) > 64:
            _struct_cache.clear()
        _struct_cache[key] = out
    return out


def _push_consts(charge, mass, dt, c2):
    # mass ** 2 exactly as turbopy/computetools.py:430 evaluates it (Python /
    # numpy scalar power), so the kernel starts from the reference's own value.
    return _lib.PushConsts(float(charge), float(mass ** 2), float(dt), float(c2))


class _FieldArg:
    """E or B of ``push`` lowered to (pointer, mode, strides) + keep-alives."""

    def __init__(self, F, name, n, device):
        self.keep = None
        self.sn = self.sc = 0
        if type(F) is np.ndarray and F.dtype == np.float64 and F.shape == (3,) and F.flags.c_contiguous:
            # the common case (a module's E / B vector): the library reads the three
            # host doubles at call time, no copy needed
            self.mode = _lib.TP_FIELD_UNIFORM_HOST
            self.keep = F
            self.ptr = F.__array_interface__["data"][0]
            return
        if _is_cuda(F):
            _require_f64(F, name)
            if F.ndim == 2 and F.shape == (n, 3) and n != 1:
                self.mode = _lib.TP_FIELD_PER_PARTICLE
                self.sn, self.sc = (int(s) for s in F.stride())
                if self.sn <= 0 or self.sc <= 0:
                    raise ValueError(f"{name}: unsupported strides")
                self.keep = F
            elif F.shape in ((3,), (1, 3)):
                self.mode = _lib.TP_FIELD_UNIFORM_DEV
                self.keep = F.reshape(3).contiguous()
            else:
                raise ValueError(f"{name} must have shape (3,), (1,3) or (n,3); got {tuple(F.shape)}")
            self.ptr = self.keep.data_ptr()
            return
        A = np.asarray(F.cpu() if isinstance(F, torch.Tensor) else F, dtype=np.float64)
        if A.ndim == 0:
            if name == "B":
                # np.cross(vminus, t) with a scalar t (computetools.py:436)
                raise ValueError("incompatible dimensions for cross product\n(dimension must be 2 or 3)")
            A = np.full(3, float(A))        # dt*E*charge/2 broadcasts a scalar E
        if A.shape in ((3,), (1, 3)):
            self.mode = _lib.TP_FIELD_UNIFORM_HOST
            self.keep = (C.c_double * 3)(*A.reshape(3))
            self.ptr = C.addressof(self.keep)
        elif A.shape == (n, 3):
            self.mode = _lib.TP_FIELD_PER_PARTICLE
            if device is None:              # host path: C-order per-particle array
                self.keep = np.ascontiguousarray(A)
                self.ptr = self.keep.ctypes.data
            else:
                self.keep = torch.from_numpy(np.ascontiguousarray(A)).to(device)
                self.ptr = self.keep.data_ptr()
            self.sn, self.sc = 3, 1
        else:
            raise ValueError(f"{name} must have shape (3,), (1,3) or (n,3); got {A.shape}")


class GridOnDevice:
    """Device copies of the ``Grid`` attributes the kernels read (uploaded once
    per grid and device): ``r`` (turbopy/core.py:666-667), ``cell_widths`` (:670),
    and the host scalars ``dr``, ``r[0]``, ``r[-1]``."""

    _cache = {}

    def __init__(self, r_host, dr, device):
        r_host = np.ascontiguousarray(np.asarray(r_host, dtype=np.float64))
        self.ng = int(r_host.shape[0])
        self.dr = float(dr)
        self.r_first = float(r_host[0])
        self.r_last = float(r_host[-1])
        self.r = torch.from_numpy(r_host).to(device)
        self._widths = None
        self._mean_width = None
        self._r_host = r_host
        self._linspace = False              # not probed yet

    @staticmethod
    def probe_linspace(r, bounds=(None, None)):
        """``_lib.Linspace`` if the host array ``r`` equals ``start + span*(i*step)`` (last node
        ``start + span*1.0``) bit for bit, else None."""
        n = r.shape[0]
        if n < 2 or os.environ.get("TURBOPY_B200_LINSPACE", "1") == "0":
            return None
        step = 1.0 / (n - 1)
        t = np.arange(n, dtype=np.float64) * step
        t[-1] = 1.0
        start = float(r[0])
        spans = [float(r[-1]) - start]
        lo, hi = bounds
        if lo is not None and hi is not None:
            spans.insert(0, float(hi) - float(lo))      # the reference's own (r_max - r_min)
        for span in spans:
            if np.array_equal((start + span * t).view(np.uint64), r.view(np.uint64)):
                return _lib.Linspace(start, span, step, n)
        return None

    @property
    def linspace(self):
        """A ``tp_linspace`` describing ``Grid.r`` exactly, or None.  turboPy builds
        ``r = r_min + (r_max - r_min) * np.linspace(0, 1, N)`` (turbopy/core.py:666-667,709),
        i.e. ``r[i] = r_min + span*(i*step)`` with ``step = 1/(N-1)`` and the last node
        ``r_min + span*1.0``, every operation rounded once.  When the grid's array matches
        that formula bit for bit (checked here on the host, once per grid) the kernels that
        only need ``r`` rebuild it in registers instead of reading it."""
        if self._linspace is False:
            self._linspace = self.probe_linspace(self._r_host, getattr(self, "_bounds", (None, None)))
        return self._linspace

    @property
    def mean_cell_width(self):
        """``np.mean(Grid.cell_widths)`` (turbopy/computetools.py:49), formed once
        on the host with numpy so that it carries numpy's own pairwise rounding."""
        if self._mean_width is None:
            self._mean_width = float(np.mean(self._r_host[1:] - self._r_host[:-1]))
        return self._mean_width

    def widths_follow_r(self, grid):
        """True when ``grid.cell_widths`` is what the reference computes from ``r``
        (``r[1:] - r[:-1]``, turbopy/core.py:670) -- checked once per grid."""
        if getattr(self, "_widths_ok", None) is None:
            w = np.asarray(getattr(grid, "cell_widths", None))
            self._widths_ok = bool(w.shape == (self.ng - 1,) and
                                   np.array_equal(w.view(np.uint64), (self._r_host[1:] - self._r_host[:-1]).view(np.uint64)))
        return self._widths_ok

    @property
    def cell_widths(self):
        if self._widths is None:
            self._widths = torch.from_numpy(self._r_host[1:] - self._r_host[:-1]).to(self.r.device)
        return self._widths

    @classmethod
    def of(cls, grid, device):
        device = torch.device(device)
        key = (id(grid), device.type, device.index if device.index is not None else torch.cuda.current_device())
        hit = cls._cache.pop(key, None)
        if hit is None or hit[0] is not grid:
            while len(cls._cache) >= 32:    # grids are few and long-lived: drop the least recently used one only
                cls._cache.pop(next(iter(cls._cache)))
            hit = (grid, cls(grid.r, grid.dr, device))
            hit[1]._bounds = (getattr(grid, "r_min", None), getattr(grid, "r_max", None))
        cls._cache[key] = hit               # (re-)inserted last = most recently used
        return hit[1]


class GatherStatus:
    """The 4-word device status block of the gather kernels, polled lazily."""

    def __init__(self, device):
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            self.words = torch.zeros(4, dtype=torch.int64, device=self.device)
            self.reset()

    def reset(self):
        _lib.check(_lib.lib().tp_status_reset(C.c_void_p(self.words.data_ptr()), _stream_ptr(self.device)))

    def read(self):
        w = self.words.cpu().numpy().view(np.uint64)
        return int(w[0]), int(w[1]), int(w[2])

    @staticmethod
    def message(count, first, bits):
        why = []
        if bits & _lib.TP_ST_BELOW:
            why.append("below the interpolation range's minimum value")
        if bits & _lib.TP_ST_ABOVE:
            why.append("above the interpolation range's maximum value")
        if bits & _lib.TP_ST_BAD_WINDOW:
            why.append("Error finding requested point in the grid")
        return (f"{count} particle coordinate(s) outside the grid (first index {first}): "
                + "; ".join(why) + " -- those particles were left unchanged")

    def check(self):
        """Raise like the reference would have: ValueError for the interp1d
        bounds check (formulas B, C), AssertionError for create_interpolator's
        asserts (formula A, turbopy/core.py:726-729)."""
        count, first, bits = self.read()
        if count:
            self.reset()
            exc = AssertionError if bits & _lib.TP_ST_BAD_WINDOW else ValueError
            raise exc(self.message(count, first, bits))


def _field_key(F):
    """Hashable identity of a grid-field argument (None if it cannot be cached)."""
    if F is None:
        return 0
    if isinstance(F, torch.Tensor):
        return (F.data_ptr(), tuple(F.shape), F.stride(), F.dtype)
    if isinstance(F, (tuple, list)) and len(F) == 3 and all(c is None or isinstance(c, torch.Tensor) for c in F):
        return tuple(0 if c is None else (c.data_ptr(), tuple(c.shape), c.stride(), c.dtype) for c in F)
    return None


def _grid_components(F, name, ng):
    """Lower a grid field to three (tensor-or-None, node stride) pairs."""
    if F is None or (np.isscalar(F) and F == 0):
        return [(None, 1)] * 3
    if isinstance(F, (tuple, list)):
        if len(F) != 3:
            raise ValueError(f"{name}: a component tuple needs 3 entries (None = zero component)")
        out = []
        for c, comp in enumerate(F):
            if comp is None:
                out.append((None, 1))
                continue
            if comp.ndim != 1 or comp.shape[0] != ng:
                raise ValueError(f"{name}[{c}] must have shape ({ng},), got {tuple(comp.shape)}")
            _require_f64(comp, f"{name}[{c}]")
            st = comp.stride(0) if isinstance(comp, torch.Tensor) else comp.strides[0] // 8
            out.append((comp, int(st)))
        return out
    if getattr(F, "ndim", None) == 2 and tuple(F.shape) == (ng, 3):
        # Grid.generate_field(3) layout (turbopy/core.py:675-698)
        _require_f64(F, name)
        return [(F[:, c], int(F[:, c].stride(0) if isinstance(F, torch.Tensor) else F[:, c].strides[0] // 8))
                for c in range(3)]
    raise ValueError(f"{name} must be None, an ({ng},3) array, or a 3-tuple of ({ng},) arrays / None; "
                     "a bare (ng,) array is ambiguous (which component?)")


# --------------------------------------------------------------------------- BorisPush
class BorisPush(ComputeTool):
    """Calculate charged particle motion in electric and magnetic fields
    (relativistic Boris push) on a B200.

    Replaces ``turbopy.computetools.BorisPush`` (computetools.py:384-441) under
    the same name; ``input_data`` takes no options the stock tool does not have.
    As in the reference, the constructor touches neither the clock, the grid
    nor the device; ``owner.clock.dt`` is read at call time (:427).
    """

    def __init__(self, owner, input_data):
        super().__init__(owner, input_data)
        self.c2 = 2.9979e8 ** 2             # computetools.py:405 (pinned by tests/test_computetools.py:12)
        self._status = {}
        self._grid_cache = {}
        self._windows = {}                  # per ensemble: tile bounds of the windowed large-grid gather
        self._records = {}                  # per (grid size, device): scratch for the packed node records
        # additive option: re-order the particles by cell every K gather_push calls (0 = never).  The reference
        # never reorders particles and its physics does not depend on their order, but callers that index
        # particles by position in the array would notice, hence opt-in.
        self.sort_every = int(input_data.get("sort_every", 0) or 0)
        # every full_sort_every-th of those re-orderings is a global sort again (0 = only the first one): a block-local
        # re-sort moves a particle at most 960 positions, so over thousands of steps the tiles slowly widen
        # (profiles/r2_window_kernel.md: 10 -> 54 cells per tile after 500 steps at K = 100)
        self.full_sort_every = int(input_data.get("full_sort_every", 10) or 0)
        self._sort_calls = {}

    # -- the reference method ------------------------------------------------
    def push(self, position, momentum, charge, mass, E, B):
        """Update position and momentum in place (computetools.py:407-441).

        ``position``/``momentum``: (n,3) float64, CUDA tensors (resident path)
        or numpy arrays (end-to-end host path).  ``E``/``B``: anything the
        reference broadcasts against (n,3): (3,), (1,3) or (n,3).
        """
        k = _push_consts(charge, mass, self.owner.clock.dt, self.c2)
        lib = _lib.lib()
        if _is_cuda(position):
            if not _is_cuda(momentum) or momentum.device != position.device:
                raise TypeError("position and momentum must live on the same CUDA device")
            p = _particles_struct(position, momentum)
            dev = position.device
            # a KineticEnergyDiagnostic (fuse_with_push) waiting for the moments of exactly these momenta?
            diag = _diagnostics.take_moments_request(momentum) if _diagnostics._PENDING else None
            e = _FieldArg(E, "E", p.n, dev)
            b = _FieldArg(B, "B", p.n, dev)
            with _on_device(dev):
                if diag is not None:
                    mass_d, mass_sq, c2_d, row, scratch = diag.moments_request()
                    _lib.check(lib.tp_boris_push_moments_f64(C.byref(p), C.byref(k),
                                                             C.c_void_p(e.ptr), e.mode, e.sn, e.sc,
                                                             C.c_void_p(b.ptr), b.mode, b.sn, b.sc,
                                                             mass_d, mass_sq, c2_d, C.c_void_p(row.data_ptr()),
                                                             C.c_void_p(scratch.data_ptr()), _stream_ptr(dev)))
                    diag.moments_delivered()
                else:
                    _lib.check(lib.tp_boris_push_f64(C.byref(p), C.byref(k),
                                                     C.c_void_p(e.ptr), e.mode, e.sn, e.sc,
                                                     C.c_void_p(b.ptr), b.mode, b.sn, b.sc, _stream_ptr(dev)))
            return None
        if isinstance(position, np.ndarray) and isinstance(momentum, np.ndarray):
            p = _particles_struct(position, momentum)
            _pin_host(position); _pin_host(momentum)
            e = _FieldArg(E, "E", p.n, None)
            b = _FieldArg(B, "B", p.n, None)
            _lib.check(lib.tp_boris_push_host_f64(C.byref(p), C.byref(k),
                                                  C.c_void_p(e.ptr), int(e.mode == _lib.TP_FIELD_PER_PARTICLE),
                                                  C.c_void_p(b.ptr), int(b.mode == _lib.TP_FIELD_PER_PARTICLE)))
            return None
        raise TypeError("position/momentum must both be CUDA float64 tensors or both numpy float64 arrays")

    def push_steps(self, position, momentum, charge, mass, E, B, steps):
        """``steps`` consecutive :meth:`push` calls with unchanged ``E`` and ``B``
        (what a time loop over a static field does, e.g. the uniform E x B of
        tests/test_computetools.py:40-62) in one pass over the particles:
        bit-identical to calling :meth:`push` ``steps`` times, but each particle
        is read and written once, so the kernel is bound by fp64 issue instead
        of HBM bandwidth.  CUDA tensors only (additive; not a reference method)."""
        steps = int(steps)
        if steps < 0:
            raise ValueError("steps must be >= 0")
        if not _is_cuda(position):
            raise TypeError("push_steps needs CUDA float64 tensors (use push() for numpy arrays)")
        if not _is_cuda(momentum) or momentum.device != position.device:
            raise TypeError("position and momentum must live on the same CUDA device")
        k = _push_consts(charge, mass, self.owner.clock.dt, self.c2)
        p = _particles_struct(position, momentum)
        dev = position.device
        if _diagnostics._PENDING:
            _diagnostics.flush_moments_request(momentum)
        e = _FieldArg(E, "E", p.n, dev)
        b = _FieldArg(B, "B", p.n, dev)
        with _on_device(dev):
            _lib.check(_lib.lib().tp_boris_push_steps_f64(C.byref(p), C.byref(k),
                                                          C.c_void_p(e.ptr), e.mode, e.sn, e.sc,
                                                          C.c_void_p(b.ptr), b.mode, b.sn, b.sc, steps, _stream_ptr(dev)))
        return None

    # -- the fused path --------------------------------------------------------
    def gather_push_steps(self, position, momentum, charge, mass, steps, E_grid=None, B_grid=None,
                          axis=0, formula="interp", grid=None):
        """``steps`` consecutive :meth:`gather_push` calls on unchanged grid fields (a
        static field map) in one pass over the particles: bit-identical to calling
        :meth:`gather_push` ``steps`` times, with one read and one write of the
        particle state.  A particle that leaves the grid during the pass is flagged
        once and keeps the state it left with.  CUDA tensors only (additive)."""
        if not _is_cuda(position):
            raise TypeError("gather_push_steps needs CUDA float64 tensors")
        steps = int(steps)
        if steps < 0:
            raise ValueError("steps must be >= 0")
        return self.gather_push(position, momentum, charge, mass, E_grid, B_grid, axis, formula, grid, _steps=steps)

    def gather_push(self, position, momentum, charge, mass, E_grid=None, B_grid=None,
                    axis=0, formula="interp", grid=None, _steps=None):
        """Gather E and B at ``position[:, axis]`` by linear interpolation of the
        grid fields and push, in one kernel.

        ``E_grid`` / ``B_grid``: ``None`` (zero field), an ``(ng,3)`` array
        (``Grid.generate_field(3)``), or a 3-tuple of ``(ng,)`` arrays / ``None``
        per component.  ``formula`` selects the reference's rounding: ``"interp"``
        (``Interpolators.interpolate1D`` -> numpy.interp, the default),
        ``"create_interpolator"`` (``Grid.create_interpolator``) or
        ``"interp1d_nd"`` (scipy's N-D linear path).  ``grid``: the owner's
        ``Grid`` by default.  Particles whose coordinate is outside the grid are
        left unchanged and reported by :meth:`check_status` (the reference
        raises before pushing anyone; a kernel cannot).
        """
        grid = self.owner.grid if grid is None else grid
        k = _push_consts(charge, mass, self.owner.clock.dt, self.c2)
        code = _FORMULAS[formula]
        if axis not in (0, 1, 2):
            raise ValueError("axis must be 0, 1 or 2")
        lib = _lib.lib()
        p = _particles_struct(position, momentum)
        ng = int(grid.num_points) if hasattr(grid, "num_points") else int(len(grid.r))
        if _is_cuda(position):
            dev = position.device
            # a Simulation passes the same field tensors every step: the lowered tp_grid_fields is cached on
            # (grid, device, tensor pointers / shapes / strides); a struct only holds pointers and strides, so a
            # hit on recycled memory of the same shape is still the right struct
            key = (id(grid), dev, _field_key(E_grid), _field_key(B_grid))
            hit = self._grid_cache.get(key) if None not in key[2:] else None
            if hit is not None and hit[0] is grid:
                g = hit[1]
            else:
                comps = _grid_components(E_grid, "E_grid", ng) + _grid_components(B_grid, "B_grid", ng)
                g = _lib.GridFields()
                g.ng, g.dr = ng, float(grid.dr)
                gd = GridOnDevice.of(grid, dev)
                g.r, g.r_first, g.r_last = gd.r.data_ptr(), gd.r_first, gd.r_last
                for c, (t, st) in enumerate(comps):
                    if t is not None and not (_is_cuda(t) and t.device == dev):
                        raise TypeError("grid fields must be CUDA tensors on the particles' device")
                    g.comp[c] = t.data_ptr() if t is not None else None
                    g.comp_stride[c] = st
                if None not in key[2:]:
                    if len(self._grid_cache) > 16:
                        self._grid_cache.clear()
                    # gd rides along: the struct holds gd.r's device pointer, which must outlive the cached struct
                    # even if GridOnDevice's own cache lets go of it
                    self._grid_cache[key] = (grid, g, gd)
            status = self._status.get(dev)
            if status is None:
                status = self._status[dev] = GatherStatus(dev)
            if self.sort_every > 0:
                from . import particles
                ck = (position.data_ptr(), int(p.n), axis)
                calls = self._sort_calls.get(ck, 0)
                if calls % self.sort_every == 0:
                    # the first time (and every full_sort_every-th time) a global sort; in between the cheap
                    # block-local re-sort, which is all a slowly decaying order needs (SoA ensembles)
                    nth = calls // self.sort_every
                    soa = p.pos_stride_n == 1 and p.mom_stride_n == 1
                    if nth == 0 or not soa or (self.full_sort_every > 0 and nth % self.full_sort_every == 0):
                        particles.sort_by_cell(position, momentum, grid, axis=axis)
                    else:
                        particles.resort_local(position, momentum, grid, axis=axis, phase=nth)
                self._sort_calls[ck] = calls + 1
            # a KineticEnergyDiagnostic (fuse_with_push) waiting for the moments of exactly these momenta?
            diag = _diagnostics.take_moments_request(momentum) if _diagnostics._PENDING else None
            with _on_device(dev):
                if diag is not None and _steps is not None:      # multi-step pass: the standalone reduction, first
                    diag._flush()
                    diag = None
                if diag is not None:
                    # the step's push delivers the diagnostic's row: inside the kernel where a fused instance exists
                    # (formula B, SoA, TMA-pipelined), by the standalone reduction in front of the push otherwise
                    hit = self._tile_windows(p, g, position, momentum, axis, grid, dev, code)
                    win = hit[0] if hit is not None and hit[2] else None
                    stream = _stream_ptr(dev)
                    rec = self._record_scratch(ng, dev, stream) if hit is not None else None
                    mass_d, mass_sq, c2_d, row, scratch = diag.moments_request()
                    _lib.check(lib.tp_gather_push_moments_f64(
                        C.byref(p), C.byref(k), C.byref(g), axis, code,
                        C.c_void_p(win.data_ptr()) if win is not None else None, win.shape[0] if win is not None else 0,
                        C.c_void_p(rec.data_ptr()) if rec is not None else None, mass_d, mass_sq, c2_d,
                        C.c_void_p(row.data_ptr()), C.c_void_p(scratch.data_ptr()),
                        C.c_void_p(status.words.data_ptr()), stream))
                    diag.moments_delivered()
                elif _steps is None:
                    hit = self._tile_windows(p, g, position, momentum, axis, grid, dev, code)
                    if hit is not None:             # grid too large for shared memory: caller-owned record scratch
                        win = hit[0] if hit[2] else None        # (capturable on any stream); windows only if they pay
                        stream = _stream_ptr(dev)
                        rec = self._record_scratch(ng, dev, stream)
                        _lib.check(lib.tp_gather_push_windows_f64(C.byref(p), C.byref(k), C.byref(g), axis, code,
                                                                  C.c_void_p(win.data_ptr()) if win is not None else None,
                                                                  win.shape[0] if win is not None else 0,
                                                                  C.c_void_p(rec.data_ptr()),
                                                                  C.c_void_p(status.words.data_ptr()), _stream_ptr(dev)))
                    else:
                        _lib.check(lib.tp_gather_push_f64(C.byref(p), C.byref(k), C.byref(g), axis, code,
                                                          C.c_void_p(status.words.data_ptr()), _stream_ptr(dev)))
                else:
                    _lib.check(lib.tp_gather_push_steps_f64(C.byref(p), C.byref(k), C.byref(g), axis, code, _steps,
                                                            C.c_void_p(status.words.data_ptr()), _stream_ptr(dev)))
            return None
        if isinstance(position, np.ndarray):
            if not isinstance(momentum, np.ndarray):
                raise TypeError("position/momentum must both be CUDA float64 tensors or both numpy float64 arrays")
            _pin_host(position); _pin_host(momentum)
            comps = _grid_components(E_grid, "E_grid", ng) + _grid_components(B_grid, "B_grid", ng)
            g = _lib.GridFields()
            g.ng, g.dr = ng, float(grid.dr)
            r = np.ascontiguousarray(grid.r, dtype=np.float64)
            g.r, g.r_first, g.r_last = r.ctypes.data, float(r[0]), float(r[-1])
            _pin_host(r)                        # large grids: the per-call upload of r and the fields is DMA, not a
            for c, (t, st) in enumerate(comps):  # staged copy from pageable memory (56 MB per call at config 3's size)
                if t is not None and not isinstance(t, np.ndarray):
                    raise TypeError("host path: grid fields must be numpy arrays")
                if t is not None:
                    _pin_host(t)
                g.comp[c] = t.ctypes.data if t is not None else None
                g.comp_stride[c] = st
            words = (C.c_uint64 * 4)()
            _lib.check(lib.tp_gather_push_host_f64(C.byref(p), C.byref(k), C.byref(g), axis, code, words))
            if words[0]:
                exc = AssertionError if words[2] & _lib.TP_ST_BAD_WINDOW else ValueError
                raise exc(GatherStatus.message(int(words[0]), int(words[1]), int(words[2])))
            return None
        raise TypeError("position/momentum must both be CUDA float64 tensors or both numpy float64 arrays")

    def _record_scratch(self, ng, dev, stream):
        """Caller-owned record scratch of the large-grid calls, one per stream (calls on a stream are ordered)."""
        rkey = (ng, dev, stream.value)
        rec = self._records.get(rkey)
        if rec is None:
            if len(self._records) > 8:
                self._records.clear()
            rec = self._records[rkey] = torch.empty(18 * ng, dtype=torch.float64, device=dev)
        return rec

    def _tile_windows(self, p, g, position, momentum, axis, grid, dev, code=_lib.TP_INTERP_B):
        """``[tile-bounds tensor, order generation, windows pay]`` of this ensemble for tp_gather_push_windows_f64
        (None when the library would not use per-tile windows: grid staged whole in shared memory, small ensemble,
        odd layout).  Primed from the
        current coordinates the first time an ensemble is seen and whenever the particles were re-ordered
        (``sort_every``, or a ``tp.sort_by_cell`` call by anyone); afterwards each call leaves the next call's
        windows behind.  Stale bounds cost speed for one step, never correctness (include/turbopy_b200.h)."""
        from . import particles
        key = (position.data_ptr(), int(p.n), axis, id(grid), dev)
        hit = self._windows.get(key)
        if hit is None:
            need = int(_lib.lib().tp_gather_windows_len(C.byref(p), C.byref(g)))
            if len(self._windows) > 8:
                self._windows.clear()
            hit = self._windows[key] = [torch.zeros(need, dtype=torch.int32, device=dev) if need else None, -1, True]
        buf, generation, pays = hit
        if buf is None:
            return None                     # windows do not apply (grid staged whole, small ensemble, odd layout)
        if generation != particles.order_generation:
            _lib.check(_lib.lib().tp_gather_windows_init(C.byref(p), C.byref(g), axis, C.c_void_p(buf.data_ptr()),
                                                         buf.shape[0], _stream_ptr(dev)))
            hit[1] = particles.order_generation
            # Do windows pay for this ensemble?  In random order no tile's cells fit a window and every particle takes
            # the L2 path, for which round 1's packed 64-byte node records (tp_gather_push_f64) are the better table
            # (the 144-byte cell records do not fit L2 next to the particle stream: 31 vs 25 G pushes/s).  Decided
            # from the primed bounds, once per ensemble and re-ordering (one small D2H read; skipped under capture).
            hit[2] = True
            if not torch.cuda.is_current_stream_capturing():
                cap = int(_lib.lib().tp_gather_windows_capacity(C.byref(p), C.byref(g), code))
                spans = buf.view(-1, 2)[: int(p.n) // 960, 1]
                hit[2] = bool(((spans > 0) & (spans <= cap)).float().mean().item() >= 0.5) if spans.numel() else True
        return hit

    def check_status(self):
        """Poll the device error flags of :meth:`gather_push` (one small D2H
        copy + sync per device): raises ValueError / AssertionError if any
        particle left the grid since the last check."""
        for status in self._status.values():
            status.check()


# --------------------------------------------------------------------------- Interpolators
class _Interp1D:
    """The callable ``interpolate1D`` returns: ``f(x_new)`` evaluates the
    piecewise-linear interpolant on the GPU with scipy's rounding --
    numpy.interp's for 1-D ``y`` (formula B), ``interp1d._call_linear``'s for
    N-D ``y`` (formula C; interpolation along the last axis, interp1d's
    default axis=-1).  Out-of-range points raise ValueError like
    ``interp1d(bounds_error=True)``."""

    def __init__(self, x, y, bounds=(None, None)):
        self._numpy = not isinstance(x, torch.Tensor)
        if self._numpy:
            if not torch.cuda.is_available():
                _lib.check(_lib.TP_E_NODEVICE)
            dev = torch.device("cuda", torch.cuda.current_device())
            x = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(dev)
            y = torch.from_numpy(np.ascontiguousarray(np.asarray(y, dtype=np.float64))).to(dev)
        if not (_is_cuda(x) and _is_cuda(y)) or x.device != y.device:
            raise TypeError("x and y must be CUDA float64 tensors on one device (or numpy arrays)")
        _require_f64(x, "x"); _require_f64(y, "y")
        if x.ndim != 1:
            raise ValueError("the x array must have exactly one dimension.")
        if y.ndim == 0:
            raise ValueError("the y array must have at least one dimension.")
        if y.shape[-1] != x.shape[0]:
            raise ValueError("x and y arrays must be equal in length along interpolation axis.")
        if x.shape[0] < 2:
            raise ValueError("x and y arrays must have at least 2 entries")
        x_host = np.ascontiguousarray(x.cpu().numpy())
        if x_host.shape[0] > 1 and not bool(np.all(x_host[1:] >= x_host[:-1])):
            # interp1d(assume_sorted=False), scipy's default: ``ind = np.argsort(x, kind="mergesort")``,
            # ``x = x[ind]``, ``y = np.take(y, ind, axis=axis)`` -- the same stable sort, on the device
            ind = torch.sort(x, stable=True).indices
            x = x[ind]
            y = y.index_select(-1, ind)
            x_host = np.ascontiguousarray(x.cpu().numpy())
        self.x = x.contiguous()
        self.y = y
        self._y2 = y.reshape(-1, y.shape[-1]).contiguous()
        ends = self.x[[0, -1]].cpu()
        self.r_first, self.r_last = float(ends[0]), float(ends[1])
        self.dr = (self.r_last - self.r_first) / (x.shape[0] - 1)      # index guess only
        self.formula = _lib.TP_INTERP_B if y.ndim == 1 else _lib.TP_INTERP_C
        self._status = GatherStatus(x.device)
        # abscissas that are exactly turboPy's linspace grid (Grid.r): the kernel rebuilds them in registers
        self._lin = GridOnDevice.probe_linspace(x_host, bounds)

    #: interp1d raises its bounds error inside the call, which costs one device
    #: sync per evaluation; set ``f.deferred_check = True`` to keep evaluations
    #: asynchronous (out-of-range points come back as NaN) and call ``f.check()``
    #: when convenient.
    deferred_check = False

    def check(self):
        """Raise interp1d's bounds ValueError if any point evaluated since the
        last check was outside ``[x[0], x[-1]]``."""
        count, first, bits = self._status.read()
        if count:
            self._status.reset()
            bad = float(self._last_xq[first]) if getattr(self, "_last_xq", None) is not None else float("nan")
            side = "below" if bits & _lib.TP_ST_BELOW and bad < self.r_first else "above"
            bound = self.r_first if side == "below" else self.r_last
            raise ValueError(f"A value ({bad}) in x_new is {side} the interpolation range's "
                             f"{'minimum' if side == 'below' else 'maximum'} value ({bound}).")

    def __call__(self, x_new):
        dev = self.x.device
        as_numpy = not isinstance(x_new, torch.Tensor)
        xq = torch.as_tensor(np.asarray(x_new, dtype=np.float64)).to(dev) if as_numpy else x_new
        if not _is_cuda(xq):
            xq = xq.to(dev)
        _require_f64(xq, "x_new")
        shape = tuple(xq.shape)
        xq = xq.reshape(-1).contiguous()
        n = xq.shape[0]
        k = self._y2.shape[0]
        out = torch.empty((k, n), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            head = (C.c_void_p(self.x.data_ptr()), self.x.shape[0], self.dr, self.r_first, self.r_last,
                    C.c_void_p(self._y2.data_ptr()), k, self._y2.stride(0), self._y2.stride(1),
                    C.c_void_p(xq.data_ptr()), 1, n, C.c_void_p(out.data_ptr()), n, self.formula)
            tail = (C.c_void_p(self._status.words.data_ptr()), _stream_ptr(dev))
            if self._lin is not None:
                _lib.check(_lib.lib().tp_interp1d_lin_f64(*head, C.byref(self._lin), *tail))
            else:
                _lib.check(_lib.lib().tp_interp1d_f64(*head, *tail))
        self._last_xq = xq
        if not self.deferred_check:
            self.check()
        out = out.reshape(self.y.shape[:-1] + shape)
        return out.cpu().numpy() if (as_numpy and self._numpy) else out


class Interpolators(ComputeTool):
    """Interpolate a function y(x) given y at grid points in x -- replaces
    ``turbopy.computetools.Interpolators`` (computetools.py:444-481)."""

    def __init__(self, owner, input_data):
        super().__init__(owner, input_data)

    def interpolate1D(self, x, y, kind='linear'):
        """Return an interpolating function ``f(x_new)`` (computetools.py:459-481).
        Only ``kind='linear'`` is on the particle hot path and implemented here;
        spline kinds raise (they are scipy/LAPACK code, not a B200 kernel)."""
        if kind != 'linear':
            # Spline kinds are scipy / LAPACK set-up code on a handful of nodes, not part of the particle hot
            # path: host arrays go to the stock tool this class displaced in the registry, so a
            # Simulation that used them keeps working after ``import turbopy_b200``.  This is a hand-off of an
            # out-of-scope call, not a fallback of the linear path: that one has no CPU implementation here.
            stock = stock_tools.get("Interpolators")
            if stock is not None and not isinstance(x, torch.Tensor) and not isinstance(y, torch.Tensor):
                return stock.interpolate1D(self, x, y, kind)
            raise NotImplementedError(
                f"interpolate1D(kind={kind!r}): only 'linear' runs on the GPU (the field gather of the particle "
                "hot path); spline kinds need numpy arrays and the stock turboPy tool")
        grid = getattr(self.owner, "grid", None)
        return _Interp1D(x, y, (getattr(grid, "r_min", None), getattr(grid, "r_max", None)))


# --------------------------------------------------------------------------- FiniteDifference
class BandedOperator:
    """What the ``FiniteDifference`` matrix builders return in place of a
    ``scipy.sparse.dia_matrix``: supports ``D @ f`` / ``D.dot(f)`` on CUDA
    tensors (and numpy arrays, staged through the GPU), ``shape``, ``offsets``,
    ``data`` and ``toarray()`` -- the surface tests/test_computetools.py:141-291
    uses.  ``apply`` is the kernel launch; the entries are only materialised on
    request (by applying the operator to unit vectors)."""

    def __init__(self, n, offsets, apply_fn, device, stored=None):
        self.shape = (n, n)
        self.offsets = np.asarray(offsets, dtype=np.int32)
        self._apply = apply_fn
        self._device = device
        self._dense = None
        self._diag = None
        self._axpy = None                # fused `f + dtau * (self @ f)` kernel, if the operator has one
        self._stored = stored            # callable -> (ndiag, n) numpy array of the stored diagonals, if known

    # ---- constructors for the two stored forms --------------------------------
    @classmethod
    def from_constants(cls, n, offsets, coef, overrides, device):
        """Diagonal k is the constant ``coef[k]`` except ``data[k][j] = v`` for
        the ``(k, j, v)`` in ``overrides`` (j within 4 columns of a boundary).
        Applied by tp_fd_dia_const_apply_f64: only the operand is read."""
        offsets = [int(o) for o in offsets]
        nd = len(offsets)
        c_coef = (C.c_double * nd)(*[float(c) for c in coef])
        c_off = (C.c_int * nd)(*offsets)
        nov = len(overrides)
        c_ok = (C.c_int * max(nov, 1))(*[int(k) for k, _, _ in overrides])
        c_oj = (C.c_int64 * max(nov, 1))(*[int(j) % n for _, j, _ in overrides])
        c_ov = (C.c_double * max(nov, 1))(*[float(v) for _, _, v in overrides])

        def apply(f):
            out = torch.empty_like(f)
            with torch.cuda.device(f.device):
                _lib.check(_lib.lib().tp_fd_dia_const_apply_f64(
                    c_coef, c_off, nd, c_ok, c_oj, c_ov, nov,
                    C.c_void_p(f.data_ptr()), C.c_void_p(out.data_ptr()), n, _stream_ptr(f.device)))
            return out

        def stored():
            data = np.empty((nd, n))
            for k in range(nd):
                data[k, :] = coef[k]
            for k, j, v in overrides:
                data[k, j] = v
            return data

        return cls(n, offsets, apply, device, stored)

    @classmethod
    def from_diagonals(cls, data, offsets, device):
        """Stored diagonals ``data`` ((ndiag, n) CUDA tensor, scipy's dia
        layout), applied by tp_fd_dia_apply_f64 in scipy's summation order."""
        data = data.contiguous()
        nd, n = data.shape
        offsets = [int(o) for o in offsets]
        if nd > 5 or any(abs(o) > 2 for o in offsets):
            raise NotImplementedError("banded operators are applied up to bandwidth 5 (offsets -2 .. 2); "
                                      f"got offsets {offsets}")
        c_off = (C.c_int * nd)(*offsets)

        def apply(f):
            out = torch.empty_like(f)
            with torch.cuda.device(f.device):
                _lib.check(_lib.lib().tp_fd_dia_apply_f64(
                    C.c_void_p(data.data_ptr()), c_off, nd, C.c_void_p(f.data_ptr()),
                    C.c_void_p(out.data_ptr()), n, _stream_ptr(f.device)))
            return out

        op = cls(n, offsets, apply, device, lambda: data.cpu().numpy())
        if offsets == list(range(offsets[0], offsets[-1] + 1)):
            op._diag = (offsets, data)
        return op

    def _compose(self, other):
        """``self @ other``: apply ``other`` first (operator product, lazily)."""
        if other.shape != self.shape:
            raise ValueError("dimension mismatch")
        lo = int(self.offsets.min() + other.offsets.min())
        hi = int(self.offsets.max() + other.offsets.max())
        return BandedOperator(self.shape[0], list(range(lo, hi + 1)), lambda f: self._apply(other._apply(f)),
                              self._device)

    def dot(self, f):
        if isinstance(f, BandedOperator):
            return self._compose(f)
        if isinstance(f, torch.Tensor):
            if not _is_cuda(f):
                raise TypeError("operand must be a CUDA tensor (or a numpy array)")
            _require_f64(f, "operand")
            if f.ndim not in (1, 2) or f.shape[0] != self.shape[0]:
                raise ValueError("dimension mismatch")
            if f.ndim == 2:                         # a block of column vectors, like scipy's A @ X
                return torch.stack([self._apply(f[:, c].contiguous()) for c in range(f.shape[1])], dim=1)
            return self._apply(f.contiguous())
        if hasattr(f, "todia") and getattr(f, "shape", None) == self.shape:     # a scipy sparse matrix
            return self._compose(self._coerce(f))
        a = np.asarray(f, dtype=np.float64)
        if a.ndim not in (1, 2) or a.shape[0] != self.shape[0]:
            raise ValueError("dimension mismatch")
        out = self.dot(torch.from_numpy(np.ascontiguousarray(a)).to(self._device))
        return out.cpu().numpy()

    __matmul__ = dot

    def axpy(self, f, dtau):
        """``f + dtau * (self @ f)`` as a fresh array -- the explicit update step of a field module
        (``f += dtau * (D @ f)``), with numpy's rounding (matvec, scale, add).  ``del2_radial()`` does it in one
        kernel pass (tp_fd_del2_radial_axpy_f64: 16 B per point); other operators apply, scale and add."""
        if isinstance(f, torch.Tensor) and _is_cuda(f) and f.ndim == 1 and self._axpy is not None:
            _require_f64(f, "operand")
            if f.shape[0] != self.shape[0]:
                raise ValueError("dimension mismatch")
            return self._axpy(f.contiguous(), dtau)
        return f + dtau * self.dot(f)

    # ---- the stored diagonals of ANY banded operator, without forming it ------
    def diagonals(self):
        """``(offsets, data)``: every diagonal between the lowest and the highest offset, as an ``(nd, n)`` CUDA
        tensor in scipy's dia layout (``data[k, j] = A[j - offsets[k], j]``).  Extracted exactly by applying the
        operator to ``nd`` comb vectors (ones at the columns ``j = c mod nd``): a row of a matrix of bandwidth
        ``nd`` meets exactly one column of each comb, so each output element IS one matrix entry (times 1.0, plus
        zeros).  ``nd`` kernel launches, O(nd * n) work, no dense matrix."""
        if self._diag is None:
            n = self.shape[0]
            lo, hi = int(self.offsets.min()), int(self.offsets.max())
            w = hi - lo + 1
            dev = self._device
            i = torch.arange(n, device=dev)
            data = torch.zeros((w, n), dtype=torch.float64, device=dev)
            for c in range(w):
                col = self._apply((i % w == c).to(torch.float64))
                j = i + lo + ((c - (i + lo)) % w)          # the comb's only column inside row i's band
                ok = (j >= 0) & (j < n)
                data[(j - i - lo)[ok], j[ok]] = col[ok]
            self._diag = (list(range(lo, hi + 1)), data)
        return self._diag

    def toarray(self):
        if self._dense is None:
            offsets, data = self.diagonals()
            n = self.shape[0]
            dense = torch.zeros((n, n), dtype=torch.float64, device=self._device)
            j = torch.arange(n, device=self._device)
            for k, off in enumerate(offsets):
                ok = (j - off >= 0) & (j - off < n)
                dense[(j - off)[ok], j[ok]] = data[k][ok] + 0.0        # a stored -0.0 reads +0.0, as in scipy's toarray()
            self._dense = dense.cpu().numpy()
        return self._dense

    # ---- sparse-matrix arithmetic (what scipy's dia_matrix gives the reference's callers) ----
    # computetools.py:221 itself forms ``D1 + D2``; user modules scale and combine operators the same way
    # (``1 - dt * D``-style updates).  Results are new BandedOperators with merged diagonals, applied by
    # tp_fd_dia_apply_f64 in ascending-offset (= column) order, the order scipy's csr result sums in.
    def _coerce(self, other):
        if isinstance(other, BandedOperator):
            return other
        if hasattr(other, "todia"):
            m = other.todia()
            if m.shape != self.shape:
                raise ValueError("inconsistent shapes")
            order = np.argsort(m.offsets)
            data = torch.from_numpy(np.ascontiguousarray(m.data[order], dtype=np.float64)).to(self._device)
            if data.shape[1] < self.shape[0]:
                data = torch.nn.functional.pad(data, (0, self.shape[0] - data.shape[1]))
            return BandedOperator.from_diagonals(data, m.offsets[order], self._device)
        return None

    def _merged(self, other, sign):
        other = self._coerce(other)
        if other is None:
            return NotImplemented
        if other.shape != self.shape:
            raise ValueError("inconsistent shapes")
        (oa, da), (ob, db) = self.diagonals(), other.diagonals()
        lo, hi = min(oa[0], ob[0]), max(oa[-1], ob[-1])
        n = self.shape[0]
        data = torch.zeros((hi - lo + 1, n), dtype=torch.float64, device=self._device)
        data[oa[0] - lo:oa[-1] - lo + 1] = da
        rows = data[ob[0] - lo:ob[-1] - lo + 1]
        rows += db.to(self._device) if sign > 0 else -db.to(self._device)
        return BandedOperator.from_diagonals(data, list(range(lo, hi + 1)), self._device)

    def __add__(self, other):
        return self._merged(other, +1)

    __radd__ = __add__

    def __sub__(self, other):
        return self._merged(other, -1)

    def __rsub__(self, other):
        return (-self).__add__(other)

    def _scaled(self, c):
        if not np.isscalar(c):
            return NotImplemented
        offsets, data = self.diagonals()
        return BandedOperator.from_diagonals(data * float(c), offsets, self._device)

    def __mul__(self, c):
        return self._scaled(c)

    __rmul__ = __mul__

    def __truediv__(self, c):
        if not np.isscalar(c):
            return NotImplemented
        return self._scaled(1.0 / float(c))         # scipy's sparse true-divide is a multiplication by 1/c

    def __neg__(self):
        return self._scaled(-1.0)

    def transpose(self):
        offsets, data = self.diagonals()
        n = self.shape[0]
        out = torch.zeros_like(data)
        for k, off in enumerate(offsets):           # A^T[j, j - off] = A[j - off, j]: diagonal -off, re-indexed by column
            src = data[k]
            if off >= 0:
                out[k, :n - off] = src[off:]
            else:
                out[k, -off:] = src[:n + off]
        # the diagonals keep their stored order (offsets now descending), as scipy's dia transpose does: the
        # apply kernel sums in stored order, so (A.T @ f) rounds like scipy's
        return BandedOperator.from_diagonals(out, [-o for o in offsets], self._device)

    T = property(transpose)

    def todia(self):
        """A host ``scipy.sparse.dia_matrix`` with the same diagonals (for spsolve & co.)."""
        from scipy import sparse
        offsets, data = self.diagonals()
        return sparse.dia_matrix((data.cpu().numpy(), offsets), shape=self.shape)

    def tocsr(self):
        return self.todia().tocsr()

    def tocsc(self):
        return self.todia().tocsc()

    def tocoo(self):
        return self.todia().tocoo()

    @property
    def data(self):
        """scipy's dia layout: ``data[k, j] = A[j - offsets[k], j]`` (the stored
        diagonals themselves, unused corner entries included, when the operator
        was built from them)."""
        if self._stored is not None:
            return self._stored()
        offsets, data = self.diagonals()
        data = data.cpu().numpy()
        return np.stack([data[offsets.index(int(o))] for o in self.offsets])


class FiniteDifference(ComputeTool):
    """Finite-difference stencils on the 1-D grid -- replaces
    ``turbopy.computetools.FiniteDifference`` (computetools.py:62-381) for the
    operators on the hot path.  ``input_data["method"]`` selects
    ``"centered"`` / ``"upwind_left"`` for :meth:`setup_ddx`."""

    def __init__(self, owner, input_data):
        super().__init__(owner, input_data)
        self.dr = self.owner.grid.dr            # cached like the reference (:85)

    def setup_ddx(self):
        assert (self.input_data["method"] in ["centered", "upwind_left"])
        if self.input_data["method"] == "centered":
            return self.centered_difference
        if self.input_data["method"] == "upwind_left":
            return self.upwind_left

    def _device_in(self, y):
        """(device tensor, was_numpy)"""
        if isinstance(y, torch.Tensor):
            if not _is_cuda(y):
                raise TypeError("y must be a CUDA tensor (or a numpy array)")
            t = y
        else:
            if not torch.cuda.is_available():
                _lib.check(_lib.TP_E_NODEVICE)
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(y, dtype=np.float64))).cuda()
        _require_f64(t, "y")
        n = int(self.owner.grid.num_points)
        if t.ndim != 1 or t.shape[0] != n:
            raise ValueError(f"could not broadcast input array from shape {tuple(t.shape)} into shape ({n},)")
        return t.contiguous(), not isinstance(y, torch.Tensor)

    def centered_difference(self, y):
        """Centered estimate of dy/dx; a fresh array (computetools.py:104-120)."""
        t, was_numpy = self._device_in(y)
        d = torch.empty_like(t)
        with torch.cuda.device(t.device):
            _lib.check(_lib.lib().tp_fd_centered_f64(C.c_void_p(t.data_ptr()), C.c_void_p(d.data_ptr()),
                                                     t.shape[0], 2 * self.dr, _stream_ptr(t.device)))
        return d.cpu().numpy() if was_numpy else d

    def upwind_left(self, y):
        """Left upwind estimate of dy/dx; a fresh array (computetools.py:122-138)."""
        t, was_numpy = self._device_in(y)
        gd = GridOnDevice.of(self.owner.grid, t.device)
        d = torch.empty_like(t)
        with torch.cuda.device(t.device):
            if gd.linspace is not None and gd.widths_follow_r(self.owner.grid):
                _lib.check(_lib.lib().tp_fd_upwind_left_lin_f64(C.c_void_p(t.data_ptr()), C.byref(gd.linspace),
                                                                C.c_void_p(d.data_ptr()), t.shape[0],
                                                                _stream_ptr(t.device)))
            else:
                w = gd.cell_widths
                _lib.check(_lib.lib().tp_fd_upwind_left_f64(C.c_void_p(t.data_ptr()), C.c_void_p(w.data_ptr()),
                                                            C.c_void_p(d.data_ptr()), t.shape[0],
                                                            _stream_ptr(t.device)))
        return d.cpu().numpy() if was_numpy else d

    def del2_radial(self, device=None):
        """Operator for (1/r)(d/dr)(r df/dr) (computetools.py:189-223): apply it
        with ``D @ f``.  The tridiagonal rows are rebuilt inside the kernel
        from ``Grid.r`` and ``dr``; no diagonals are stored."""
        if not torch.cuda.is_available():
            _lib.check(_lib.TP_E_NODEVICE)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        gd = GridOnDevice.of(self.owner.grid, dev)
        g1 = 1 / (2.0 * self.dr)                # :197
        g2 = 1 / (self.dr ** 2)                 # :211
        n = gd.ng

        lin = gd.linspace

        def apply(f):
            out = torch.empty_like(f)
            with torch.cuda.device(f.device):
                if lin is not None:
                    _lib.check(_lib.lib().tp_fd_del2_radial_apply_lin_f64(
                        C.c_void_p(f.data_ptr()), C.byref(lin), C.c_void_p(out.data_ptr()),
                        n, g1, g2, _stream_ptr(f.device)))
                else:
                    _lib.check(_lib.lib().tp_fd_del2_radial_apply_f64(
                        C.c_void_p(f.data_ptr()), C.c_void_p(gd.r.data_ptr()), C.c_void_p(out.data_ptr()),
                        n, g1, g2, _stream_ptr(f.device)))
            return out

        def axpy(f, dtau):
            out = torch.empty_like(f)
            with torch.cuda.device(f.device):
                _lib.check(_lib.lib().tp_fd_del2_radial_axpy_f64(
                    C.c_void_p(f.data_ptr()), None if lin is not None else C.c_void_p(gd.r.data_ptr()),
                    C.byref(lin) if lin is not None else None, C.c_void_p(out.data_ptr()), n, g1, g2, float(dtau),
                    _stream_ptr(f.device)))
            return out

        op = BandedOperator(n, [-1, 0, 1], apply, dev)
        op._axpy = axpy
        return op

    # ---- the remaining matrix builders -------------------------
    # Operators whose diagonals are constant apart from a few boundary entries
    # carry their coefficients as kernel arguments (nothing but the operand is
    # read); radial_curl, whose entries depend on r, stores its diagonals.
    def _device(self, device):
        if not torch.cuda.is_available():
            _lib.check(_lib.TP_E_NODEVICE)
        return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def _n(self):
        return int(self.owner.grid.num_points)

    def ddx(self, device=None):
        """Centered df/dx, coefficients -+1/(2 dr) (computetools.py:140-155)."""
        g = 1 / (2.0 * self.dr)
        return BandedOperator.from_constants(self._n(), [-1, 1], [0.0 - g, 0.0 + g], [], self._device(device))

    def ddr(self, device=None):
        """df/dr with the r = 0 row zeroed (computetools.py:246-263)."""
        g1 = 1 / (2.0 * self.dr)
        return BandedOperator.from_constants(self._n(), [-1, 1], [-g1, g1], [(1, 1, 0.0)], self._device(device))

    def del2(self, device=None):
        """d2f/dx2 with the doubled first-row coefficient (computetools.py:225-244)."""
        g2 = 1 / (self.dr ** 2)
        return BandedOperator.from_constants(self._n(), [-1, 0, 1], [g2, -2 * g2, g2], [(2, 1, 2 * g2)],
                                             self._device(device))

    def radial_curl(self, device=None):
        """(1/r) d(r f)/dr (computetools.py:157-187); r-dependent entries, stored on the device."""
        dev = self._device(device)
        n = self._n()
        r = GridOnDevice.of(self.owner.grid, dev).r
        g = 1 / (2.0 * self.dr)
        data = torch.zeros((3, n), dtype=torch.float64, device=dev)
        data[0, :-1] = -g * (r[:-1] / r[1:])
        data[2, 1:] = g * (r[1:] / r[:-1])
        data[2, 1] = 2.0 / self.dr              # r = 0: B ~ linear
        data[1, -1] = 1.0 / self.dr             # r = R_w
        data[0, -2] = 2.0 * data[0, -1]
        return BandedOperator.from_diagonals(data, [-1, 0, 1], dev)

    def _bc_left(self, c1, c2, device):
        over = [(0, 0, 0.0), (1, 1, c1)] + ([(2, 2, c2)] if c2 is not None else [])
        offs, coef = ([0, 1, 2], [1.0, 0.0, 0.0]) if c2 is not None else ([0, 1], [1.0, 0.0])
        return BandedOperator.from_constants(self._n(), offs, coef, over, self._device(device))

    def BC_left_extrap(self, device=None):
        """Linear extrapolation to the left boundary point (computetools.py:265-287)."""
        return self._bc_left(2, -1, device)

    def BC_left_avg(self, device=None):
        """computetools.py:289-311."""
        return self._bc_left(1.5, -0.5, device)

    def BC_left_quad(self, device=None):
        """Quadratic extrapolation to the left boundary (computetools.py:313-335)."""
        r = self.owner.grid.r
        R2 = (r[1] ** 2 + r[2] ** 2) / (r[2] ** 2 - r[1] ** 2) / 2
        return self._bc_left(0.5 + R2, 0.5 - R2, device)

    def BC_left_flat(self, device=None):
        """Neumann condition at the left boundary (computetools.py:337-357)."""
        return self._bc_left(1, None, device)

    def BC_right_extrap(self, device=None):
        """Linear extrapolation to the right boundary point (computetools.py:359-381)."""
        n = self._n()
        return BandedOperator.from_constants(n, [-2, -1, 0], [0.0, 0.0, 1.0],
                                             [(2, n - 1, 0.0), (1, n - 2, 2.0), (0, n - 3, -1.0)],
                                             self._device(device))


# --------------------------------------------------------------------------- PoissonSolver1DRadial
class PoissonSolver1DRadial(ComputeTool):
    """Solve the 1-D radial Poisson equation by two running sums -- replaces
    ``turbopy.computetools.PoissonSolver1DRadial`` (computetools.py:20-59).
    The running sums are parallel fp64 scans on the GPU, so the result agrees
    with numpy's sequential ``cumsum`` to rounding, not bit for bit."""

    def __init__(self, owner, input_data):
        super().__init__(owner, input_data)
        self._scratch = {}

    def solve(self, sources):
        """``sources``: (N,) CUDA float64 tensor (or numpy array, staged through
        the GPU).  Returns a fresh array of the same kind (computetools.py:35-59)."""
        grid = self.owner.grid
        was_numpy = not isinstance(sources, torch.Tensor)
        if was_numpy:
            if not torch.cuda.is_available():
                _lib.check(_lib.TP_E_NODEVICE)
            s = torch.from_numpy(np.ascontiguousarray(np.asarray(sources, dtype=np.float64))).cuda()
        else:
            if not _is_cuda(sources):
                raise TypeError("sources must be a CUDA tensor (or a numpy array)")
            s = sources.contiguous()
        _require_f64(s, "sources")
        n = int(grid.num_points)
        if s.ndim != 1 or s.shape[0] != n:
            raise ValueError(f"operands could not be broadcast together with shapes ({n},) {tuple(s.shape)}")
        dev = s.device
        gd = GridOnDevice.of(grid, dev)
        dr = gd.mean_cell_width                             # :49, cached per grid
        lib = _lib.lib()
        need = int(lib.tp_poisson_scratch_doubles(n))
        scratch = self._scratch.get(dev)
        if scratch is None or scratch.shape[0] < need:
            scratch = self._scratch[dev] = torch.empty(need, dtype=torch.float64, device=dev)
        out = torch.empty_like(s)
        with torch.cuda.device(dev):
            if gd.linspace is not None:
                _lib.check(lib.tp_poisson_radial_solve_lin_f64(C.c_void_p(s.data_ptr()), C.byref(gd.linspace), dr,
                                                               C.c_void_p(out.data_ptr()), n,
                                                               C.c_void_p(scratch.data_ptr()), _stream_ptr(dev)))
            else:
                _lib.check(lib.tp_poisson_radial_solve_f64(C.c_void_p(s.data_ptr()), C.c_void_p(gd.r.data_ptr()), dr,
                                                           C.c_void_p(out.data_ptr()), n,
                                                           C.c_void_p(scratch.data_ptr()), _stream_ptr(dev)))
        return out.cpu().numpy() if was_numpy else out

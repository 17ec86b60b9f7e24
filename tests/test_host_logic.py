"""CPU: host-side logic of the drop-in tools (no kernel launches)."""
import ast
import os

import numpy as np
import pytest

from helpers import make_grid, make_owner

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registry_semantics():
    """turbopy/core.py:302-327: duplicate -> ValueError unless override,
    non-subclass -> TypeError, unknown lookup -> KeyError."""
    import turbopy_b200 as tp
    if not tp.HAVE_TURBOPY:
        pytest.skip("no turboPy on sys.path (neither /root/reference nor oracle/_ref): there is no registry to test")
    assert tp.ComputeTool.lookup("BorisPush") is tp.BorisPush
    assert tp.ComputeTool.lookup("Interpolators") is tp.Interpolators
    assert tp.ComputeTool.lookup("FiniteDifference") is tp.FiniteDifference
    assert tp.Diagnostic.lookup("kinetic_energy") is tp.KineticEnergyDiagnostic
    with pytest.raises(ValueError, match="already registered"):
        tp.ComputeTool.register("BorisPush", tp.BorisPush)
    tp.ComputeTool.register("BorisPush", tp.BorisPush, override=True)
    with pytest.raises(TypeError):
        tp.ComputeTool.register("NotATool", dict)
    with pytest.raises(KeyError):
        tp.ComputeTool.lookup("nope")
    assert tp.ComputeTool.is_valid_name("FiniteDifference") and not tp.ComputeTool.is_valid_name("nope")


def test_constructors_touch_nothing():
    """Reference tests build tools from a bare Simulation with no grid/clock
    (tests/test_computetools.py:9-10); the clock does not exist yet when tools
    are constructed (turbopy/core.py:166-178)."""
    import turbopy_b200 as tp
    from types import SimpleNamespace
    owner = SimpleNamespace(clock=None, grid=None)
    t = tp.BorisPush(owner, {"type": "BorisPush"})
    assert t.c2 == 2.9979e8 ** 2 and t.name == "BorisPush" and t.owner is owner
    assert callable(t.push) and callable(t.gather_push)          # bound methods exist before initialize()
    i = tp.Interpolators(owner, {"type": "Interpolators"})
    assert callable(i.interpolate1D)
    f = tp.FiniteDifference(make_owner(grid=make_grid(10, 0.0, 10.0)), {"type": "FiniteDifference", "method": "centered"})
    assert f.dr == make_grid(10, 0.0, 10.0).dr
    assert f.setup_ddx() == f.centered_difference
    t.initialize(); i.initialize(); f.initialize()


def test_field_argument_lowering():
    from turbopy_b200 import _lib
    from turbopy_b200.tools import _FieldArg
    e = _FieldArg(np.array([1.0, 2.0, 3.0]), "E", 5, None)
    assert e.mode == _lib.TP_FIELD_UNIFORM_HOST and list(e.keep) == [1.0, 2.0, 3.0]
    e = _FieldArg(np.array([[1.0, 2.0, 3.0]]), "E", 5, None)
    assert e.mode == _lib.TP_FIELD_UNIFORM_HOST
    e = _FieldArg(0, "E", 5, None)                      # scalar E broadcasts (dt*E*charge/2)
    assert list(e.keep) == [0.0, 0.0, 0.0]
    with pytest.raises(ValueError, match="cross product"):
        _FieldArg(0, "B", 5, None)                      # np.cross with a scalar
    pp = _FieldArg(np.ones((5, 3)), "B", 5, None)
    assert pp.mode == _lib.TP_FIELD_PER_PARTICLE and (pp.sn, pp.sc) == (3, 1)
    with pytest.raises(ValueError):
        _FieldArg(np.ones((4, 3)), "B", 5, None)


def test_particle_layout_checks():
    from turbopy_b200.tools import _n3_strides
    a = np.zeros((7, 3))
    assert _n3_strides(a, "position") == (3, 1)                       # numpy C order = AoS
    assert _n3_strides(np.asfortranarray(a), "position") == (1, 7)    # Fortran order = SoA
    with pytest.raises(IndexError):
        _n3_strides(np.zeros(3), "position")                          # reference: IndexError on (3,) particles
    with pytest.raises(ValueError):
        _n3_strides(np.zeros((3, 7)), "position")
    with pytest.raises(TypeError):
        _n3_strides(np.zeros((7, 3), dtype=np.float32), "position")


def test_grid_component_lowering():
    from turbopy_b200.tools import _grid_components
    ng = 12
    E = np.zeros((ng, 3))
    comps = _grid_components(E, "E_grid", ng)
    assert [st for _, st in comps] == [3, 3, 3]
    assert comps[1][0].ctypes.data == E.ctypes.data + 8
    comps = _grid_components((None, np.zeros(ng), None), "E_grid", ng)
    assert comps[0][0] is None and comps[1][1] == 1
    assert _grid_components(None, "B_grid", ng) == [(None, 1)] * 3
    with pytest.raises(ValueError, match="ambiguous"):
        _grid_components(np.zeros(ng), "E_grid", ng)
    with pytest.raises(ValueError):
        _grid_components((None, np.zeros(ng + 1), None), "E_grid", ng)


def test_soa_particles_layout():
    import turbopy_b200 as tp
    v = tp.soa_particles(7, device="cpu", fill=0.0)
    assert tuple(v.shape) == (7, 3) and v.stride() == (1, 8)
    v[3, 1] = 5.0
    assert v[:, 1][3] == 5.0 and v[:, 1].is_contiguous()
    assert v[0, :].shape == (3,)                                     # diagnostics index data[0, :]


def test_shard_ranges_cover_everything():
    from turbopy_b200.sharding import shard_range
    for n in (0, 1, 7, 1000, 10 ** 9):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under turbopy_b200/ may import it."""
    pkg = os.path.join(ROOT, "turbopy_b200")
    for name in os.listdir(pkg):
        if not name.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, name)).read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom) and node.module:
                mods = [node.module]
            assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), (name, mods)


def test_interpolate1d_hands_splines_to_the_stock_tool():
    """Spline kinds are outside the GPU path: host arrays go to the stock tool the import displaced
    (the reference's own tests/test_computetools.py:82-86 call pattern), device tensors raise."""
    import turbopy_b200 as tp
    tool = tp.Interpolators(make_owner(), {"type": "Interpolators"})
    x, y = np.arange(6.0), np.exp(-np.arange(6.0) / 3.0)
    if tp.HAVE_TURBOPY:
        from scipy import interpolate
        assert tp.stock_tools["Interpolators"].__module__ == "turbopy.computetools"
        xq = np.arange(0, 5, 0.1)
        assert np.allclose(tool.interpolate1D(x, y, "cubic")(xq), interpolate.interp1d(x, y, kind="cubic")(xq))
    else:
        with pytest.raises(NotImplementedError):
            tool.interpolate1D(x, y, "cubic")


def test_additive_apis_validate_before_touching_the_device():
    """push_steps / gather_push_steps / sort_by_cell are CUDA-tensor-only additions: host arrays, bad step counts and
    bad axes are rejected by the Python layer (no device, no library call needed), and the C ABI rejects bad
    arguments of the new entry points without a device."""
    import ctypes as C

    import torch

    import turbopy_b200 as tp
    from turbopy_b200 import _lib
    tool = tp.BorisPush(make_owner(), {"type": "BorisPush"})
    pos, mom = np.zeros((4, 3)), np.zeros((4, 3))
    with pytest.raises(TypeError):
        tool.push_steps(pos, mom, 1.0, 1.0, np.zeros(3), np.zeros(3), 3)
    with pytest.raises(ValueError):
        tool.push_steps(pos, mom, 1.0, 1.0, np.zeros(3), np.zeros(3), -1)
    with pytest.raises(TypeError):
        tool.gather_push_steps(pos, mom, 1.0, 1.0, 2)
    with pytest.raises(TypeError):
        tp.sort_by_cell(torch.zeros((4, 3), dtype=torch.float64), torch.zeros((4, 3), dtype=torch.float64),
                        make_owner().grid)
    L = _lib.lib()
    assert L.tp_sort_scratch_bytes(-1, 10, 0) == -1 and L.tp_sort_scratch_bytes(10, 0, 0) == -1
    assert L.tp_sort_scratch_bytes(1 << 33, 10, 0) == -1 and L.tp_sort_scratch_bytes(10, 10, 41) == -1
    assert L.tp_sort_scratch_bytes(1000, 1 << 20, 0) > 2 * 8 * 1000        # two permutations + one scratch plane + counts
    assert L.tp_poisson_scratch_doubles(2048 * 3 + 1) == 4 * 4 + 8
    if not torch.cuda.is_available():
        p = _lib.Particles(0, 1, 1, 0, 1, 1, 0)
        k = _lib.PushConsts(1.0, 1.0, 1.0, 1.0)
        assert L.tp_boris_push_steps_f64(C.byref(p), C.byref(k), None, 0, 0, 0, None, 0, 0, 0, 3, None) == _lib.TP_E_NODEVICE
        assert L.tp_permute_f64(None, 1, 0, None, None, None) == _lib.TP_E_NODEVICE


def test_formula_a_window_margin_is_sufficient():
    """csrc/gather_math.cuh: window_margin().  Formula A's fast path accepts a particle when it is at least `margin`
    away from both nodes of its cell and claims that r[j-1] and r[j+2] are then outside (x-dr, x+dr) as the
    reference evaluates that window (core.py:728, `r0 - dr < r` / `r < r0 + dr` with the sums rounded once).  Here
    the kernel's margin arithmetic is replayed in numpy on linspace, offset and jittered grids, and the claim is
    checked for points placed right at the margin (the adversarial spot) and at random."""
    rng = np.random.default_rng(11)

    def margin_of(r, dr):
        w = r[1:] - r[:-1]
        if not (w > 0).all():
            return 0.0
        wmin = w.min()
        wlb = wmin - wmin * 2.0 ** -52
        gap = max(dr - wlb, 0.0)
        eps = (max(abs(r[0]), abs(r[-1])) + dr) * 2.0 ** -51
        m = 2.0 * (gap + eps)
        return m if (m < dr * 0.015625 and m > 0.0) else 0.0

    checked = 0
    for trial in range(60):
        ng = int(rng.integers(5, 3000))
        lo = float(rng.choice([0.0, -3.7, 1e-3, 12345.678, -1e6]))
        span = float(rng.choice([1.0, 15.0, 1e-6, 3e4]))
        r = lo + span * np.linspace(0.0, 1.0, ng)
        dr = span / (ng - 1)
        if trial % 3 == 1:
            r[1:-1] += 1e-3 * dr * rng.uniform(-1, 1, ng - 2)
        if trial % 3 == 2:
            r[1:-1] += 7e-3 * dr * rng.uniform(-1, 1, ng - 2)
        qa = margin_of(r, dr)
        resolved = dr > 2.0 ** 20 * np.spacing(max(abs(r[0]), abs(r[-1])))      # cells far wider than an ulp of r
        if trial % 3 != 2 and resolved:
            assert qa > 0.0                       # turboPy's own grids (and slightly perturbed ones) get a margin
        if qa == 0.0:
            continue
        j = rng.integers(0, ng - 1, 20000)
        x0, x1 = r[j], r[j + 1]
        # at the margin from the left node, from the right node (a few ulps either side), and uniformly inside
        t = rng.integers(0, 3, len(j))
        k = rng.integers(-3, 4, len(j))
        x = np.where(t == 0, x0 + qa, np.where(t == 1, x1 - qa, x0 + (x1 - x0) * rng.uniform(0, 1, len(j))))
        x = x + k * np.spacing(np.abs(x)) * (t != 2)
        accepted = (x0 < x) & (x < x1) & ((x - x0) >= qa) & ((x1 - x) >= qa)
        lo_edge, hi_edge = x - dr, x + dr
        xm = np.where(j > 0, r[np.maximum(j - 1, 0)], -np.inf)
        xp = np.where(j < ng - 2, r[np.minimum(j + 2, ng - 1)], np.inf)
        inside = (lo_edge < xm) | (xp < hi_edge)
        assert not (accepted & inside).any()
        assert accepted.mean() > 0.3               # and the margin is not so wide that nothing passes
        checked += int(accepted.sum())
    assert checked > 100_000


def test_deferred_diagnostic_protocol(monkeypatch):
    """Host side of ``KineticEnergyDiagnostic(fuse_with_push=True)`` (no device: the reduction is replaced by a numpy
    stand-in and CPU tensors are declared deliverable): diagnose() only posts a request keyed by the momentum tensor;
    a push takes it exactly once; values() / the next diagnose() / finalize() / flush_moments_request() deliver an
    untaken request from the momenta as they are at that moment; rows stay in call order."""
    import torch
    import turbopy_b200 as tp
    from turbopy_b200 import diagnostics as D

    calls = []

    def fake_reduce(momentum, mass, c2, out=None, scratch=None):
        calls.append(float(momentum.sum()))
        out[:] = 0.0
        out[0], out[5] = float((momentum * momentum).sum()), float(momentum.shape[0])
        return out

    monkeypatch.setattr(D, "reduce_moments", fake_reduce)
    monkeypatch.setattr(D, "_can_ride_on_push", lambda m: True)
    monkeypatch.setattr(D._lib, "lib", lambda: type("L", (), {"tp_reduce_scratch_doubles": staticmethod(lambda: 8)})())
    mom = torch.ones((4, 3), dtype=torch.float64)
    owner = make_owner(num_steps=6)
    d = tp.KineticEnergyDiagnostic(owner, {"momentum": "p", "mass": 1.0, "fuse_with_push": True, "interval": 2,
                                           "allreduce": False})
    d.inspect_resource({"p": mom})
    d.initialize()
    assert not D._PENDING
    d.diagnose()                                            # call 1: a diagnostic step -> posted, nothing reduced
    assert calls == [] and D._PENDING and d._pending_row is not None
    got = D.take_moments_request(mom)                       # the push takes it ...
    assert got is d and not D._PENDING and D.take_moments_request(mom) is None       # ... exactly once
    mass, mass_sq, c2, row, scratch = d.moments_request()
    assert (mass, mass_sq) == (1.0, 1.0) and row.shape == (8,) and scratch.numel() == 8
    row[:] = 0.0
    row[0] = 12.0                                           # what the fused kernel would write
    d.moments_delivered()
    assert d.fused_rows == 1 and d._pending_row is None
    d.diagnose()                                            # call 2: not a diagnostic step (interval 2)
    assert not D._PENDING
    d.diagnose()                                            # call 3: posted again
    mom.mul_(2.0)                                           # nobody pushed; values() must flush NOW (momenta as they are)
    rows = d.values()
    assert calls == [24.0] and not D._PENDING and rows.shape == (2, 6)
    assert rows[0, 0] == 12.0 and rows[1, 0] == 48.0 and d.fused_rows == 1
    d.diagnose()                                            # call 4: skipped
    d.diagnose()                                            # call 5: posted
    D.flush_moments_request(mom)                            # a path that cannot deliver (push_steps) flushes first
    assert calls == [24.0, 24.0] and not D._PENDING
    D.flush_moments_request(mom)                            # nothing pending: no-op
    assert calls == [24.0, 24.0]
    d.finalize()                                            # call 6: skipped by the interval; nothing left to flush
    assert d.values().shape == (3, 6) and calls == [24.0, 24.0]
    # a diagnostic that goes away un-finalized takes its request with it (weak registry)
    d2 = tp.KineticEnergyDiagnostic(owner, {"momentum": "p", "mass": 1.0, "fuse_with_push": True, "allreduce": False})
    d2.inspect_resource({"p": mom})
    d2.initialize()
    d2.diagnose()
    assert len(D._PENDING) == 1
    del d2
    import gc
    gc.collect()
    assert not D._PENDING

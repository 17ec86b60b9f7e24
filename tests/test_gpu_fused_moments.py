"""The energy / moment diagnostic riding on the push (tp_gather_push_moments_f64,
``KineticEnergyDiagnostic`` with ``fuse_with_push``): the sums of the momenta BEFORE the push, accumulated inside the
windowed gather+push kernel from the momenta it has just loaded.  Checked against the standalone reduction
(tp_reduce_moments_f64, 1e-12: the sum order differs), against the oracle (oracle/moments.py, exactly rounded sums),
and the particles against the un-fused run bit for bit; every path that cannot fuse must deliver the standalone
reduction's row of the PRE-push momenta."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import M_E, Q_E, assert_bits_equal, make_grid, make_owner
from oracle import moments as omoments
from test_gpu_windows import NAME, TILE, ensemble, fields_on

pytestmark = pytest.mark.gpu
C2 = 2.9979e8 ** 2


@pytest.fixture(scope="module")
def tp():
    import turbopy_b200
    return turbopy_b200


def close(a, b, tol=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), 1e-300)
    return bool(np.all(np.abs(a - b) <= tol * scale))


def oracle_row(mom_host):
    return omoments.moments(mom_host[:, 0], mom_host[:, 1], mom_host[:, 2], M_E, C2)[:6]


def diag_for(tp, owner, mom, fuse, interval=1):
    d = tp.KineticEnergyDiagnostic(owner, {"momentum": "p", "mass": M_E, "interval": interval, "fuse_with_push": fuse})
    d.inspect_resource({"p": mom})
    d.initialize()
    return d


def launches_of(tp):
    from turbopy_b200 import _lib
    return _lib.launch_count()


@pytest.mark.parametrize("shape", ["windows", "staged6", "staged_ey"])
@pytest.mark.parametrize("n", [148 * TILE * 2 + 333, 148 * TILE * 3, 148 * TILE + 1])
def test_fused_rows_match_standalone_and_oracle(tp, n, shape):
    """The three fused instances (formula B, SoA): per-tile windows on a grid too large for shared memory with
    cell-ordered particles; the grid staged whole in shared memory with all six components; the pif shape (E_y only,
    the mask-specialised kernel).  Ragged tail, no tail, one tile per SM and a one-particle tail.  A fused step
    launches ONE kernel less than reduction + push (moments_partial_kernel is gone)."""
    ng = 5_001 if shape == "windows" else 1025
    grid = make_grid(ng)
    dt = 0.5 * grid.dr / 2.9979e8
    pos0, mom0 = ensemble(n, ng, 23, shape == "windows")
    comps, E, B = fields_on(grid)
    if shape == "staged_ey":
        E, B = (None, E[:, 1].contiguous(), None), None
    steps = 5
    runs = {}
    for fuse in (False, True):
        owner = make_owner(dt=dt, grid=grid, num_steps=steps)
        tool = tp.BorisPush(owner, {"type": "BorisPush"})
        pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
        diag = diag_for(tp, owner, mom, fuse)
        pre = []
        launched = 0
        for _ in range(steps):
            pre.append(mom.cpu().numpy().copy())
            l0 = launches_of(tp)
            diag.diagnose()
            tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)
            launched += launches_of(tp) - l0
        tool.check_status()
        runs[fuse] = (pos.cpu().numpy(), mom.cpu().numpy(), diag.values(), diag.fused_rows, pre, launched)
    assert runs[True][3] == steps and runs[False][3] == 0, "the push did not deliver the rows"
    assert runs[True][5] == runs[False][5] - steps, "no fused instance ran (same number of launches as reduction + push)"
    assert_bits_equal(runs[True][0], runs[False][0], "position")
    assert_bits_equal(runs[True][1], runs[False][1], "momentum")
    assert runs[True][2].shape == (steps, 6)
    assert close(runs[True][2], runs[False][2]), (runs[True][2], runs[False][2])
    assert np.array_equal(runs[True][2][:, 5], np.full(steps, float(n)))
    for s in range(steps):                                          # row s = the momenta BEFORE push s
        want = oracle_row(runs[True][4][s])
        assert close(runs[True][2][s, [0, 4, 5]], want[[0, 4, 5]]), (s, runs[True][2][s], want)
        psum_scale = np.abs(runs[True][4][s]).sum(axis=0)           # sum(p) cancels: compare against sum |p|
        assert np.all(np.abs(runs[True][2][s, 1:4] - want[1:4]) <= 1e-12 * psum_scale)


def test_fused_rows_are_deterministic(tp):
    ng, n = 5_001, 148 * TILE * 2 + 77
    grid = make_grid(ng)
    dt = 0.5 * grid.dr / 2.9979e8
    pos0, mom0 = ensemble(n, ng, 29, True)
    comps, E, B = fields_on(grid)
    rows = []
    for _ in range(2):
        owner = make_owner(dt=dt, grid=grid, num_steps=4)
        tool = tp.BorisPush(owner, {"type": "BorisPush"})
        pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
        diag = diag_for(tp, owner, mom, True)
        for _ in range(3):
            diag.diagnose()
            tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)
        rows.append(diag.values())
        assert diag.fused_rows == 3
    assert_bits_equal(rows[0], rows[1], "fused rows of two identical runs")


def test_interval_and_extreme_momenta(tp):
    """interval = 3 over 7 steps (rows at steps 0, 3, 6 + finalize); a few momenta outside the fast arithmetic's range
    (zero, 1e-200, 1e+150) send their thread's pair through the IEEE path of the accumulation."""
    ng, n = 5_001, 148 * TILE * 2 + 5
    grid = make_grid(ng)
    dt = 0.5 * grid.dr / 2.9979e8
    pos0, mom0 = ensemble(n, ng, 31, True)
    mom0[5] = 0.0
    mom0[TILE + 9] = [1e-200, -1e-200, 3e-201]
    mom0[n - 2] = [0.0, 1e-160, 0.0]                                 # in the ragged tail
    comps, E, B = fields_on(grid)
    owner = make_owner(dt=dt, grid=grid, num_steps=7)
    tool = tp.BorisPush(owner, {"type": "BorisPush"})
    pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
    diag = diag_for(tp, owner, mom, True, interval=3)
    want = []
    for s in range(7):
        if s % 3 == 0:
            want.append(oracle_row(mom.cpu().numpy()))
        diag.diagnose()
        tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)
    tool.check_status()
    got = diag.values()
    assert got.shape == (3, 6) and diag.fused_rows == 3
    for g, w in zip(got, want):
        assert close(g[[0, 4, 5]], w[[0, 4, 5]]), (g, w)


@pytest.mark.parametrize("case", ["formula_A", "formula_C_staged", "aos", "random_order", "small_ensemble", "push_aos",
                                  "push_fields", "push_steps", "gather_push_steps", "values_first"])
def test_paths_that_cannot_fuse_deliver_the_pre_push_row(tp, case):
    """Other formulas / layouts / orders / entry points: the row is the standalone reduction's, bit for bit, of the
    momenta as they were when diagnose() was called."""
    ng = 1025 if case == "formula_C_staged" else 5_001
    n = 5 * TILE + 1 if case == "small_ensemble" else 148 * TILE * 2 + 11      # fewer tiles than SMs: no ring kernel
    grid = make_grid(ng)
    dt = 0.5 * grid.dr / 2.9979e8
    pos0, mom0 = ensemble(n, ng, 37, case != "random_order")
    comps, E, B = fields_on(grid)
    owner = make_owner(dt=dt, grid=grid, num_steps=4)
    tool = tp.BorisPush(owner, {"type": "BorisPush"})
    if case in ("aos", "push_aos"):
        pos, mom = torch.from_numpy(pos0).cuda(), torch.from_numpy(mom0).cuda()
    else:
        pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
    want = tp.reduce_moments(mom, M_E, C2).cpu().numpy()[:6]
    diag = diag_for(tp, owner, mom, True)
    diag.diagnose()
    assert diag._pending_row is not None
    Eu, Bu = np.array([0.0, 1e3, 0.0]), np.array([0.0, 0.0, 0.1])
    if case == "push_aos":
        tool.push(pos, mom, Q_E, M_E, Eu, Bu)
    elif case == "push_fields":                                     # per-particle (n,3) fields: the field ring, not fused
        En = torch.from_numpy(np.tile(Eu, (n, 1))).cuda()
        Bn = torch.from_numpy(np.tile(Bu, (n, 1))).cuda()
        tool.push(pos, mom, Q_E, M_E, En, Bn)
    elif case == "push_steps":
        tool.push_steps(pos, mom, Q_E, M_E, Eu, Bu, 3)
    elif case == "gather_push_steps":
        tool.gather_push_steps(pos, mom, Q_E, M_E, 2, E_grid=E, B_grid=B)
    elif case == "values_first":
        pass
    else:
        tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B,
                         formula=NAME["A"] if case == "formula_A" else NAME["C"] if case == "formula_C_staged" else "interp")
    got = diag.values()
    assert got.shape == (1, 6) and diag._pending_row is None
    from turbopy_b200 import diagnostics
    assert not diagnostics._PENDING
    assert_bits_equal(got[0], want, f"{case}: row of the pre-push momenta")
    if case != "values_first":
        assert not np.array_equal(mom.cpu().numpy(), mom0), "the push did not run"


@pytest.mark.parametrize("n", [148 * 480 * 3 + 77, 148 * 480, 1000])
def test_push_with_uniform_fields_delivers_the_row(tp, n):
    """BorisPush.push (uniform E, B; SoA): the ring kernel's MOM instance (tp_boris_push_moments_f64) -- one launch less
    than reduction + push, particles bit-identical to the plain push, rows within 1e-12 of the oracle.  n = 1000: too
    few particles for the ring, the library launches the standalone reduction in front of the register-path push."""
    grid = make_grid(1025)
    rng = np.random.default_rng(47)
    pos0 = rng.uniform(0.0, 1.0, (n, 3))
    mom0 = M_E * 2.9979e8 * rng.normal(0.0, 0.5, (n, 3))
    Eu, Bu = np.array([0.0, 1e5, 0.0]), np.array([0.0, 0.0, 1.0])
    steps = 4
    runs = {}
    for fuse in (False, True):
        owner = make_owner(dt=1e-12, grid=grid, num_steps=steps)
        tool = tp.BorisPush(owner, {"type": "BorisPush"})
        pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
        diag = diag_for(tp, owner, mom, fuse)
        pre, launched = [], 0
        for _ in range(steps):
            pre.append(mom.cpu().numpy().copy())
            l0 = launches_of(tp)
            diag.diagnose()
            tool.push(pos, mom, Q_E, M_E, Eu, Bu)
            launched += launches_of(tp) - l0
        runs[fuse] = (pos.cpu().numpy(), mom.cpu().numpy(), diag.values(), diag.fused_rows, pre, launched)
    assert runs[True][3] == steps and runs[False][3] == 0
    ring = n >= 148 * 480
    assert runs[True][5] == runs[False][5] - (steps if ring else 0)
    assert_bits_equal(runs[True][0], runs[False][0], "position")
    assert_bits_equal(runs[True][1], runs[False][1], "momentum")
    assert close(runs[True][2], runs[False][2])
    if not ring:
        assert_bits_equal(runs[True][2], runs[False][2], "rows (standalone reduction on both sides)")
    for s in range(steps):
        want = oracle_row(runs[True][4][s])
        assert close(runs[True][2][s, [0, 4, 5]], want[[0, 4, 5]]), (s, runs[True][2][s], want)
        psum_scale = np.abs(runs[True][4][s]).sum(axis=0)
        assert np.all(np.abs(runs[True][2][s, 1:4] - want[1:4]) <= 1e-12 * psum_scale)


def test_capi_fallback_inside_the_library(tp):
    """tp_gather_push_moments_f64 with NULL tile bounds (no windows): the library itself launches the standalone
    reduction in front of the push -- out8 bit-identical to tp_reduce_moments_f64, particles to tp_gather_push_f64."""
    from turbopy_b200 import _lib
    lib = _lib.lib()
    ng, n = 5_001, 148 * TILE + 3
    grid = make_grid(ng)
    dt = 0.5 * grid.dr / 2.9979e8
    pos0, mom0 = ensemble(n, ng, 41, True)
    comps, E, B = fields_on(grid)
    tool = tp.BorisPush(make_owner(dt=dt, grid=grid), {"type": "BorisPush"})
    pos_a, mom_a = tp.to_soa(pos0), tp.to_soa(mom0)
    tool.gather_push(pos_a, mom_a, Q_E, M_E, E_grid=E, B_grid=B)
    pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
    want = tp.reduce_moments(mom, M_E, C2).cpu().numpy()
    ps = tp.tools._particles_struct(pos, mom)
    k = tp.tools._push_consts(Q_E, M_E, dt, C2)
    gf = _lib.GridFields()
    gf.ng, gf.dr, gf.r = ng, float(grid.dr), tp.GridOnDevice.of(grid, pos.device).r.data_ptr()
    gf.r_first, gf.r_last = float(grid.r[0]), float(grid.r[-1])
    for c in range(3):
        gf.comp[c], gf.comp_stride[c] = E[:, c].data_ptr(), 3
        gf.comp[3 + c], gf.comp_stride[3 + c] = B[:, c].data_ptr(), 3
    out8 = torch.full((8,), -1.0, dtype=torch.float64, device="cuda")
    scratch = torch.empty(int(lib.tp_reduce_scratch_doubles()), dtype=torch.float64, device="cuda")
    rec = torch.empty(18 * ng, dtype=torch.float64, device="cuda")
    status = torch.zeros(4, dtype=torch.int64, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.tp_gather_push_moments_f64(C.byref(ps), C.byref(k), C.byref(gf), 0, _lib.TP_INTERP_B, None, 0,
                                                      C.c_void_p(rec.data_ptr()), M_E, M_E ** 2, C2,
                                                      C.c_void_p(out8.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                                      C.c_void_p(status.data_ptr()), stream))
    torch.cuda.synchronize()
    assert_bits_equal(out8.cpu().numpy(), want, "out8")
    assert_bits_equal(pos.cpu().numpy(), pos_a.cpu().numpy(), "position")
    assert_bits_equal(mom.cpu().numpy(), mom_a.cpu().numpy(), "momentum")
    assert int(status[0]) == 0
    # argument validation
    assert lib.tp_gather_push_moments_f64(C.byref(ps), C.byref(k), C.byref(gf), 0, _lib.TP_INTERP_B, None, 0,
                                                  C.c_void_p(rec.data_ptr()), M_E, M_E ** 2, C2, None,
                                                  C.c_void_p(scratch.data_ptr()), C.c_void_p(status.data_ptr()),
                                                  stream) == -1            # TP_E_ARG


def test_fused_diagnostic_in_a_step_graph(tp):
    """[diagnose, gather_push] x K captured once and replayed: the posted request is host-side state, the kernels that
    fill the row are in the graph."""
    ng, n = 5_001, 148 * TILE * 2
    grid = make_grid(ng)
    dt = 0.5 * grid.dr / 2.9979e8
    pos0, mom0 = ensemble(n, ng, 43, True)
    comps, E, B = fields_on(grid)
    owner = make_owner(dt=dt, grid=grid, num_steps=16)
    tool = tp.BorisPush(owner, {"type": "BorisPush"})
    pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
    diag = diag_for(tp, owner, mom, True)
    diag.diagnose()
    tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)        # warm: windows primed, buffers allocated
    with tp.StepGraph(pos.device) as g:
        for _ in range(3):
            diag.diagnose()
            tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=B)
    before = mom.cpu().numpy().copy()
    g.replay()
    torch.cuda.synchronize()
    got = diag.values()
    assert got.shape == (4, 6)
    assert close(got[1, [0, 4, 5]], oracle_row(before)[[0, 4, 5]])
    # the same three steps eagerly from the same state
    pos2, mom2 = tp.to_soa(pos0), tp.to_soa(mom0)
    tool2 = tp.BorisPush(owner, {"type": "BorisPush"})
    for _ in range(4):
        tool2.gather_push(pos2, mom2, Q_E, M_E, E_grid=E, B_grid=B)
    assert_bits_equal(mom.cpu().numpy(), mom2.cpu().numpy(), "momentum after replay")

"""GPU parity of the particle reordering (additive helper; the spec is the stable
grouping stated in include/turbopy_b200.h, restated by oracle/sort.py).  Integer
work: the permutation must be identical, the arrays must be exact gathers."""
import numpy as np
import pytest
import torch

from helpers import make_grid, make_owner
from oracle import sort as osort

pytestmark = pytest.mark.gpu

Q_E, M_E = -1.6022e-19, 9.1094e-31


def layout(tp, a, kind):
    return tp.to_soa(a) if kind == "soa" else torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 1000, 100_001, 3_000_017])
@pytest.mark.parametrize("kind", ["soa", "aos"])
def test_permutation_matches_stable_argsort(n, kind):
    import turbopy_b200 as tp
    ng = (1 << 20) + 1
    grid = make_grid(ng, 0.0, 1.0)
    rng = np.random.default_rng(n + 5)
    pos0 = rng.uniform(0.0, 1.0, (n, 3)); mom0 = rng.normal(size=(n, 3))
    if n > 40:
        pos0[3, 0] = np.nan; pos0[7, 0] = -0.5; pos0[11, 0] = 1.0; pos0[13, 0] = 7.0; pos0[17, 0] = 0.0
    pos, mom = layout(tp, pos0, kind), layout(tp, mom0, kind)
    ptrs = (pos.data_ptr(), mom.data_ptr())
    perm = tp.sort_by_cell(pos, mom, grid, axis=0, return_perm=True).cpu().numpy()
    ref = osort.permutation(pos0[:, 0], grid.r[0], grid.dr, ng - 1)
    assert np.array_equal(perm, ref)
    assert (pos.data_ptr(), mom.data_ptr()) == ptrs                       # same storage
    assert np.array_equal(pos.cpu().numpy().view(np.uint64), pos0[ref].view(np.uint64))
    assert np.array_equal(mom.cpu().numpy().view(np.uint64), mom0[ref].view(np.uint64))


@pytest.mark.parametrize("ng,coarsen,axis", [(30, 0, 0), (4097, 6, 1), (4098, 0, 2), (100_000, 3, 0), (2, 0, 0), (20_000_001, 0, 0)])
def test_bins_axes_and_extras(ng, coarsen, axis):
    import turbopy_b200 as tp
    n = 200_003
    grid = make_grid(ng, -1.0, 3.0)
    rng = np.random.default_rng(ng)
    pos0 = rng.uniform(-1.0, 3.0, (n, 3)); mom0 = rng.normal(size=(n, 3))
    w0 = rng.normal(size=n); e0 = rng.normal(size=(n, 3))
    pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
    w = torch.from_numpy(w0).cuda(); e = torch.from_numpy(e0).cuda(); es = tp.to_soa(e0)
    perm = tp.sort_by_cell(pos, mom, grid, axis=axis, cells_per_bin_log2=coarsen, extra=(w, e, es), return_perm=True).cpu().numpy()
    ref = osort.permutation(pos0[:, axis], grid.r[0], grid.dr, ng - 1, coarsen)
    assert np.array_equal(perm, ref)
    b = osort.keys(pos0[ref, axis], grid.r[0], grid.dr, ng - 1, coarsen)
    assert np.all(np.diff(b) >= 0)                                        # grouped
    assert np.array_equal(pos.cpu().numpy(), pos0[ref]) and np.array_equal(mom.cpu().numpy(), mom0[ref])
    assert np.array_equal(w.cpu().numpy(), w0[ref])
    assert np.array_equal(e.cpu().numpy(), e0[ref]) and np.array_equal(es.cpu().numpy(), e0[ref])
    assert tp.sort_by_cell(pos, mom, grid, axis=axis, cells_per_bin_log2=coarsen) is None
    assert np.array_equal(pos.cpu().numpy(), pos0[ref])                   # idempotent: already grouped, stable


def test_sorted_gather_push_is_the_same_particles():
    """Reordering does not change any particle's update: gather+push on the sorted ensemble equals the
    unsorted result permuted (bit for bit), on a grid that takes the L2 record path."""
    import turbopy_b200 as tp
    ng = (1 << 18) + 1
    n = 148 * 960 * 2 + 77
    grid = make_grid(ng, 0.0, 1.0)
    tool = tp.BorisPush(make_owner(grid=grid, dt=1e-13), {"type": "BorisPush"})
    rng = np.random.default_rng(8)
    pos0 = rng.uniform(0.01, 0.99, (n, 3)); mom0 = rng.normal(0, 0.1 * M_E * 2.9979e8, (n, 3))
    r = torch.from_numpy(grid.r).cuda()
    E = torch.stack([1e4 * torch.cos(6.28 * (k + 1) * r) for k in range(3)], dim=1).contiguous()
    B = torch.stack([torch.sin(6.28 * (k + 1) * r) for k in range(3)], dim=1).contiguous()
    pa, ma = tp.to_soa(pos0), tp.to_soa(mom0)
    pb, mb = tp.to_soa(pos0), tp.to_soa(mom0)
    perm = tp.sort_by_cell(pb, mb, grid, return_perm=True).cpu().numpy()
    for _ in range(3):
        tool.gather_push(pa, ma, Q_E, M_E, E_grid=E, B_grid=B)
        tool.gather_push(pb, mb, Q_E, M_E, E_grid=E, B_grid=B)
    tool.check_status()
    assert np.array_equal(pb.cpu().numpy().view(np.uint64), pa.cpu().numpy()[perm].view(np.uint64))
    assert np.array_equal(mb.cpu().numpy().view(np.uint64), ma.cpu().numpy()[perm].view(np.uint64))


def test_sort_errors():
    import turbopy_b200 as tp
    grid = make_grid(100, 0.0, 1.0)
    pos, mom = tp.soa_particles(10, fill=0.5), tp.soa_particles(10, fill=0.0)
    with pytest.raises(ValueError):
        tp.sort_by_cell(pos, mom, grid, axis=3)
    with pytest.raises(TypeError):
        tp.sort_by_cell(pos.cpu(), mom.cpu(), grid)
    with pytest.raises(TypeError):
        tp.sort_by_cell(pos, mom, grid, extra=(torch.zeros(10, device="cuda", dtype=torch.float32),))


@pytest.mark.parametrize("n", [1, 2, 959, 960, 1919, 1920, 1921, 3841, 500_003, 3_000_017])
@pytest.mark.parametrize("phase", [0, 1])
@pytest.mark.parametrize("axis", [0, 2])
def test_resort_local_matches_blockwise_stable_argsort(n, phase, axis):
    """tp_resort_blocks_f64: inside every block of 1920 consecutive particles the stable order by cell; exact data
    movement (bit-identical gathers), in place, block boundaries shifted by 960 on odd phases."""
    import turbopy_b200 as tp
    ng = (1 << 20) + 1
    grid = make_grid(ng, 0.0, 1.0)
    rng = np.random.default_rng(n + phase)
    pos0 = rng.uniform(0.0, 1.0, (n, 3)); mom0 = rng.normal(size=(n, 3))
    # a nearly ordered ensemble (what the helper is for), with a few wild values
    pos0[:, axis] = np.sort(pos0[:, axis]) + rng.normal(0, 20.0 / ng, n)
    if n > 40:
        pos0[3, axis] = np.nan; pos0[7, axis] = -0.5; pos0[11, axis] = 7.0
    pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
    ptrs = (pos.data_ptr(), mom.data_ptr())
    tp.resort_local(pos, mom, grid, axis=axis, phase=phase)
    ref = osort.resort_blocks_permutation(pos0[:, axis], grid.r[0], grid.dr, ng - 1, offset=960 * (phase & 1))
    assert (pos.data_ptr(), mom.data_ptr()) == ptrs
    assert np.array_equal(np.sort(ref), np.arange(n))
    assert np.array_equal(pos.cpu().numpy().view(np.uint64), pos0[ref].view(np.uint64))
    assert np.array_equal(mom.cpu().numpy().view(np.uint64), mom0[ref].view(np.uint64))


def test_resort_local_keeps_a_drifting_ensemble_ordered():
    """Global sort once, then 200 steps with a local re-sort every 25: the tiles stay compact (the windows of the
    large-grid gather keep fitting) and the physics is the unsorted run's, particle for particle."""
    import turbopy_b200 as tp
    ng = (1 << 17) + 1
    n = 148 * 960 * 2 + 77
    grid = make_grid(ng, 0.0, 1.0)
    dt = 0.5 * grid.dr / 2.9979e8
    rng = np.random.default_rng(8)
    pos0 = rng.uniform(0.05, 0.95, (n, 3)); mom0 = rng.normal(0, 0.1 * M_E * 2.9979e8, (n, 3))
    tag = np.arange(n, dtype=np.float64)
    pos0[:, 2] = tag                                               # z rides along as a particle id (B = 0: z only drifts by pz)
    mom0[:, 2] = 0.0
    r = torch.from_numpy(grid.r).cuda()
    E = torch.stack([1e4 * torch.cos(6.28 * (k + 1) * r) for k in range(2)] + [torch.zeros_like(r)], dim=1).contiguous()
    plain = tp.BorisPush(make_owner(grid=grid, dt=dt), {"type": "BorisPush"})
    sorting = tp.BorisPush(make_owner(grid=grid, dt=dt), {"type": "BorisPush", "sort_every": 25})
    pa, ma = tp.to_soa(pos0), tp.to_soa(mom0)
    pb, mb = tp.to_soa(pos0), tp.to_soa(mom0)
    for _ in range(200):
        plain.gather_push(pa, ma, Q_E, M_E, E_grid=E, B_grid=None)
        sorting.gather_push(pb, mb, Q_E, M_E, E_grid=E, B_grid=None)
    plain.check_status(); sorting.check_status()
    ga, gb = pa.cpu().numpy(), pb.cpu().numpy()
    ida, idb = np.rint(ga[:, 2]).astype(np.int64), np.rint(gb[:, 2]).astype(np.int64)
    assert np.array_equal(np.sort(idb), np.arange(n))
    assert np.array_equal(ga[np.argsort(ida)].view(np.uint64), gb[np.argsort(idb)].view(np.uint64))
    assert np.array_equal(ma.cpu().numpy()[np.argsort(ida)].view(np.uint64), mb.cpu().numpy()[np.argsort(idb)].view(np.uint64))
    bounds = [w[0] for w in sorting._windows.values() if w[0] is not None][0].cpu().numpy().reshape(-1, 2)[: n // 960]
    spans_sorted = bounds[:, 1]
    cells = osort.keys(ga[:, 0], grid.r[0], grid.dr, ng - 1)[: (n // 960) * 960].reshape(-1, 960)
    spans_plain = cells.max(1) - cells.min(1) + 1
    assert spans_sorted.mean() < 0.05 * spans_plain.mean()         # random order spans the grid; re-sorted tiles stay compact
    assert spans_sorted.max() < 3 * 960 * (ng - 1) / n + 40

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

"""CUDA-graph replay of time steps equals the same steps launched eagerly."""
import numpy as np
import pytest
import torch

from helpers import M_E, Q_E, assert_bits_equal, make_grid, make_owner, thermal_ensemble

pytestmark = pytest.mark.gpu


def test_step_graph_replay_matches_eager():
    import turbopy_b200 as tp
    from turbopy_b200 import _lib
    grid = make_grid(1025)
    tool = tp.BorisPush(make_owner(dt=1e-12, grid=grid), {"type": "BorisPush"})
    pos0, mom0 = thermal_ensemble(100_000, seed=3)
    pos0[:, 0] = 0.1 + 0.8 * pos0[:, 0]                                 # stay inside the grid for 8 steps
    E = torch.from_numpy(np.stack([1e5 * np.cos(2 * np.pi * (k + 1) * grid.r) for k in range(3)], axis=1)).cuda()
    B = torch.from_numpy(np.stack([np.sin(2 * np.pi * (k + 1) * grid.r) for k in range(3)], axis=1)).cuda()
    pos_e, mom_e = tp.to_soa(pos0), tp.to_soa(mom0)
    for _ in range(8):
        tool.gather_push(pos_e, mom_e, Q_E, M_E, E_grid=E, B_grid=B)
    pos_g, mom_g = tp.to_soa(pos0), tp.to_soa(mom0)
    tool.gather_push(pos_g, mom_g, Q_E, M_E, E_grid=E, B_grid=B)      # warm-up outside the capture (step 1)
    before = _lib.launch_count()
    with tp.StepGraph() as g:
        for _ in range(7):
            tool.gather_push(pos_g, mom_g, Q_E, M_E, E_grid=E, B_grid=B)
    assert _lib.launch_count() == before                              # captured, not run
    g.replay()
    assert _lib.launch_count() == before + 7
    torch.cuda.synchronize()
    tool.check_status()
    assert_bits_equal(pos_g.cpu().numpy(), pos_e.cpu().numpy())
    assert_bits_equal(mom_g.cpu().numpy(), mom_e.cpu().numpy())
    g.close()


def test_concurrent_streams_keep_their_own_scratch():
    """Two CUDA streams driving independent ensembles at the same time (launches on torch's current stream):
    the large-grid path packs its node records into per-stream scratch, so neither stream may see the other's
    grid; results equal the serial ones bit for bit."""
    import numpy as np

    import turbopy_b200 as tp
    from helpers import M_E, Q_E, make_grid, make_owner, thermal_ensemble
    n = 148 * 960 + 123
    grids = [make_grid((1 << 16) + 1), make_grid((1 << 16) + 1, 0.0, 2.0)]
    tools = [tp.BorisPush(make_owner(dt=1e-12, grid=g), {"type": "BorisPush"}) for g in grids]
    fields, starts = [], []
    for k, g in enumerate(grids):
        comps = [1e5 * (k + 1) * np.cos((c + 1) * g.r) for c in range(3)] + [np.sin((c + 2 + k) * g.r) for c in range(3)]
        fields.append((torch.from_numpy(np.stack(comps[:3], axis=1)).cuda(), torch.from_numpy(np.stack(comps[3:], axis=1)).cuda()))
        pos0, mom0 = thermal_ensemble(n, seed=70 + k)
        pos0[:, 0] = np.random.default_rng(k).uniform(0.05, 0.95, n) * g.r[-1]
        starts.append((pos0, mom0))

    def run(concurrent):
        state = [(tp.to_soa(p), tp.to_soa(m)) for p, m in starts]
        streams = [torch.cuda.Stream(), torch.cuda.Stream()] if concurrent else [torch.cuda.current_stream()] * 2
        torch.cuda.synchronize()
        for _ in range(4):
            for k in range(2):
                with torch.cuda.stream(streams[k]):
                    tools[k].gather_push(*state[k], Q_E, M_E, E_grid=fields[k][0], B_grid=fields[k][1])
        torch.cuda.synchronize()
        for t in tools:
            t.check_status()
        return [(p.cpu().numpy(), m.cpu().numpy()) for p, m in state]

    serial, overlapped = run(False), run(True)
    for (ps, ms), (po, mo) in zip(serial, overlapped):
        assert np.array_equal(ps.view(np.uint64), po.view(np.uint64)) and np.array_equal(ms.view(np.uint64), mo.view(np.uint64))

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

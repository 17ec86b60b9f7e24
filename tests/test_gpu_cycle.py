"""the whole particle-in-field cycle on the device -- plane-wave
field update, fused gather+push, clock advance -- eagerly and as a replayed
CUDA graph."""
import numpy as np
import pytest
import torch

from helpers import M_E, Q_E, assert_bits_equal, make_grid, make_owner
from oracle import boris as oboris, gather as ogather

pytestmark = pytest.mark.gpu

OMEGA, C_WAVE, DT = 2e8, 2.998e8, 5e-10


def test_plane_wave_matches_numpy_to_ulps():
    """EMWave.update: same operation order; cos() differs from numpy's by a few ulp
    at most (stated tolerance: 4e-16 * |E0| absolute)."""
    import turbopy_b200 as tp
    grid = make_grid(4097)
    clock = tp.DeviceClock(0.0, DT)
    wave = tp.PlaneWave(grid, 1.0, OMEGA)
    k = OMEGA / C_WAVE
    for step in range(5):
        wave.update(clock)
        t = 0.0 + DT * step
        ref = 1.0 * np.cos(2 * np.pi * (-OMEGA * t + k * (grid.r - 0.5)))
        assert clock.this_step == step and clock.time == t
        assert np.abs(wave.E.cpu().numpy() - ref).max() <= 4e-16
        clock.advance()


def test_cycle_graph_equals_eager_and_oracle():
    import turbopy_b200 as tp
    grid = make_grid(1025)
    owner = make_owner(dt=DT, grid=grid)
    n, K = 200_003, 6
    rng = np.random.default_rng(5)
    pos0 = np.zeros((n, 3)); pos0[:, 0] = rng.uniform(0.05, 0.95, n)
    mom0 = np.zeros((n, 3))

    def run(graph):
        pusher = tp.BorisPush(owner, {"type": "BorisPush"})
        clock = tp.DeviceClock(0.0, DT)
        wave = tp.PlaneWave(grid, 1.0, OMEGA)
        pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
        fields = []

        def cycle():
            wave.update(clock)
            pusher.gather_push(pos, mom, Q_E, M_E, E_grid=(None, wave.E, None))
            clock.advance()

        if graph:
            cycle()                                   # step 0 eagerly (also warms everything up)
            fields.append(wave.E.cpu().numpy().copy())
            with tp.StepGraph() as g:
                cycle()
            for _ in range(K - 1):
                g.replay()
                fields.append(wave.E.cpu().numpy().copy())
            g.close()
        else:
            for _ in range(K):
                cycle()
                fields.append(wave.E.cpu().numpy().copy())
        pusher.check_status()
        assert clock.this_step == K
        return pos.cpu().numpy(), mom.cpu().numpy(), fields

    pe, me, fe = run(False)
    pg, mg, fg = run(True)
    assert_bits_equal(pe, pg, "graph vs eager position"); assert_bits_equal(me, mg, "graph vs eager momentum")
    # oracle with the device-produced fields (cos differs from numpy by ulps; everything else is exact)
    rp, rm = pos0.copy(), mom0.copy()
    for s in range(K):
        assert_bits_equal(fe[s], fg[s])
        ey, st = ogather.interp_b(rp[:, 0], grid.r, fe[s])
        assert (st == 0).all()
        E = np.zeros_like(rp); E[:, 1] = ey
        oboris.push_aos(rp, rm, Q_E, M_E, E, np.zeros(3), DT)
    assert_bits_equal(pe, rp); assert_bits_equal(me, rm)

"""Device-resident pieces of a particle-in-field cycle: the
plane-wave field update of the reference's example module and a device clock,
so that K steps of [field update -> fused gather+push -> clock.advance]
(turbopy/core.py:145-158) can be recorded once in a CUDA graph and replayed.

    clock = DeviceClock(start_time, dt, device)
    wave = PlaneWave(grid, amplitude, omega, device)          # EMWave of particle_in_field.py
    with StepGraph(device) as g:
        for _ in range(K):
            wave.update(clock)                                 # E(r, t) on the grid, t read on the device
            pusher.gather_push(pos, mom, q, m, E_grid=(None, wave.E, None))
            clock.advance()
    g.replay()                                                 # K more steps, time keeps advancing
"""
import ctypes as C

import torch

from . import _lib
from .tools import GridOnDevice


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class DeviceClock:
    """``SimulationClock`` state on the device: ``time = start_time + dt*this_step``
    (turbopy/core.py:533-538), advanced by a kernel so graphs can replay it."""

    def __init__(self, start_time, dt, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.start_time, self.dt = float(start_time), float(dt)
        self.state = torch.tensor([self.start_time, 0.0], dtype=torch.float64, device=self.device)

    def advance(self):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().tp_clock_advance_f64(C.c_void_p(self.state.data_ptr()), self.start_time, self.dt,
                                                       _stream(self.device)))

    @property
    def time(self):
        return float(self.state[0])

    @property
    def this_step(self):
        return int(self.state[1])


class PlaneWave:
    """The EMWave module's field (tests/fixtures/particle_in_field/particle_in_field.py:12-37):
    ``E(r, t) = E0*cos(2*pi*(-omega*t + k*(r - 0.5)))``, ``k = omega/c`` with the
    fixture's ``c = 2.998e8``, evaluated on the device into ``self.E``."""

    def __init__(self, grid, amplitude, omega, device=None, c=2.998e8, x0=0.5):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.E0, self.omega, self.k, self.x0 = float(amplitude), float(omega), float(omega) / c, float(x0)
        self._grid = GridOnDevice.of(grid, self.device)
        self.E = torch.zeros(self._grid.ng, dtype=torch.float64, device=self.device)

    def update(self, clock):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().tp_plane_wave_f64(
                C.c_void_p(self._grid.r.data_ptr()), self._grid.ng, self.E0, self.omega, self.k, self.x0,
                C.c_void_p(clock.state.data_ptr()), C.c_void_p(self.E.data_ptr()), 1, _stream(self.device)))

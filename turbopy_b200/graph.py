"""CUDA-graph capture of time steps through the C ABI
(tp_graph_begin / tp_graph_end / tp_graph_launch).

turboPy's ``Simulation`` loop calls ``update()`` once per step
(turbopy/core.py:145-158); at small particle counts the per-call launch overhead
dominates, so a module can record K steps once and replay them:

    with StepGraph(device) as g:          # everything launched in the block is recorded
        for _ in range(K):
            pusher.gather_push(pos, mom, q, m, E_grid=E, B_grid=B)
    g.replay()                            # K steps, one launch
"""
import ctypes as C

import torch

from . import _lib


class StepGraph:
    def __init__(self, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._exec = C.c_void_p()
        self._side = None
        self._ctx = None

    def __enter__(self):
        # capture on a side stream: the legacy default stream cannot be captured
        self._side = torch.cuda.Stream(self.device)
        self._side.wait_stream(torch.cuda.current_stream(self.device))
        self._ctx = torch.cuda.stream(self._side)
        self._ctx.__enter__()
        _lib.check(_lib.lib().tp_graph_begin(C.c_void_p(self._side.cuda_stream)))
        return self

    def __exit__(self, exc_type, exc, tb):
        rc = _lib.lib().tp_graph_end(C.c_void_p(self._side.cuda_stream), C.byref(self._exec))
        self._ctx.__exit__(exc_type, exc, tb)
        if exc_type is None:
            _lib.check(rc)
        return False

    def replay(self, stream=None):
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        _lib.check(_lib.lib().tp_graph_launch(self._exec, C.c_void_p(s.cuda_stream)))

    def close(self):
        if self._exec:
            _lib.lib().tp_graph_destroy(self._exec)
            self._exec = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

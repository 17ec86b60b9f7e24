import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from types import SimpleNamespace
import turbopy_b200 as tp
Q_E, M_E = -1.6022e-19, 9.1094e-31
tool = tp.BorisPush(SimpleNamespace(clock=SimpleNamespace(dt=1e-12), grid=None), {"type": "BorisPush"})
E, B = np.array([0.0, 1e5, 0.0]), np.array([0.0, 0.0, 1.0])
for n in (1_000_000, 16 << 20):
    pos, mom = tp.soa_particles(n), tp.soa_particles(n)
    pos.copy_(torch.rand((n, 3), dtype=torch.float64, device="cuda"))
    mom.copy_(torch.randn((n, 3), dtype=torch.float64, device="cuda") * (M_E * 2.9979e8 * 0.1))
    for K in (1, 2, 4, 8, 16, 50):
        for _ in range(3): tool.push_steps(pos, mom, Q_E, M_E, E, B, K)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 20
        e0.record()
        for _ in range(it): tool.push_steps(pos, mom, Q_E, M_E, E, B, K)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        print(f"n={n} K={K}: {ms:.4f} ms per launch, {n*K/ms/1e6:.1f} G pushes/s")

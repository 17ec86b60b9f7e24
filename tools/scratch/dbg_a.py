import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import M_E, Q_E, make_grid, make_owner, thermal_ensemble
from oracle import gather as og
import turbopy_b200 as tp
ng, n, gridkind = 1025, 7001, "jitter_large"
rng = np.random.default_rng(n + len(gridkind))
grid = make_grid(ng)
r = grid.r.copy(); r[1:-1] += 0.2 * grid.dr * rng.uniform(-1, 1, ng - 2); grid.r = r
span = r[-1] - r[0]
pos0, mom0 = thermal_ensemble(n, seed=n)
pos0[:, 0] = r[0] + span * rng.uniform(0.01, 0.99, n)
y = [1e5 * np.cos(2 * np.pi * (k + 1) * r) for k in range(3)]
_, st = og.interp_a(pos0[:, 0], r, grid.dr, y[0], r[0], r[-1])
bad = st != 0
tool = tp.BorisPush(make_owner(dt=1e-13, grid=grid), {"type": "BorisPush"})
E = torch.from_numpy(np.stack(y, axis=1)).cuda()
pos, mom = tp.to_soa(pos0), tp.to_soa(mom0)
tool.gather_push(pos, mom, Q_E, M_E, E_grid=E, B_grid=None, axis=0, formula="A")
try:
    tool.check_status()
except Exception as e:
    print("status:", str(e)[:120])
gp = pos.cpu().numpy(); gm = mom.cpu().numpy()
untouched = (gp == pos0).all(axis=1) & (gm == mom0).all(axis=1)
print("oracle bad", bad.sum(), "gpu untouched", untouched.sum(), "bad&~untouched", (bad & ~untouched).sum(), "untouched&~bad", (untouched & ~bad).sum())
idx = np.flatnonzero(bad & ~untouched)[:5]
for i in idx:
    x = pos0[i, 0]; j = np.searchsorted(r, x, side="right") - 1
    print(i, repr(x), "cell", j, [(k, (r[k] - x) / grid.dr) for k in range(j - 2, j + 4)], "guess", int((x - r[0]) / grid.dr))

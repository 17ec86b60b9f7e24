import sys
sys.path.insert(0, "/root/repo")
import torch
from types import SimpleNamespace
import turbopy_b200 as tp
from oracle import grid as ogrid
n, ng = 16 << 20, (1 << 20) + 1
r = ogrid.grid_points(0.0, 1.0, ng)
grid = SimpleNamespace(r=r, dr=ogrid.grid_dr(0.0, 1.0, ng), num_points=ng)
pos, mom = tp.soa_particles(n), tp.soa_particles(n, fill=0.0)
pos.copy_(torch.rand((n, 3), dtype=torch.float64, device="cuda"))
keep = pos.clone()
for _ in range(3):
    pos.copy_(keep)
    tp.sort_by_cell(pos, mom, grid)
torch.cuda.synchronize()
print("ok")

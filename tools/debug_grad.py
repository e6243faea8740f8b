import numpy as np, torch, sys
sys.path.insert(0, '.')
import dcol_trajectory_optimization_b200 as d
from dcol_trajectory_optimization_b200 import workloads as W
shapes, i1, i2, p1, p2 = W.config4_batch(40, seed=77)
eng = d.ProximityEngine(shapes)
h = eng.solve_host(i1, i2, p1, p2)
print("host  rowmax", np.abs(h.grad).max(axis=1)[:10].round(4), "status", h.status[:10])
plan = eng.plan(i1, i2)
d1, d2 = torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda()
for kw in [{}, {"one_pair_per_thread": True}, {"lane_refill": True}]:
    r = eng.solve(plan, d1, d2, **kw)
    torch.cuda.synchronize()
    g = r.grad.cpu().numpy()
    print(kw, "rowmax", np.abs(g).max(axis=1)[:10].round(4), "maxdiff vs host", np.abs(g - h.grad).max(), "alpha diff", np.abs(r.alpha.cpu().numpy() - h.alpha).max())

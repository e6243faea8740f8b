"""Small end-to-end run for compute-sanitizer (memcheck): every specialisation incl. runtime face counts,
record mode, host entry point, unsupported pairs."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcol_trajectory_optimization_b200 as d
from dcol_trajectory_optimization_b200 import workloads as W
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden

g = load_golden("edge_cases")                       # 16 shapes: 14-face polytope, irregular polygons, offsets, NaN/Inf poses
eng = d.ProximityEngine((g["shape_records"], g["A"], g["b"]))
r = eng.solve_host(g["idx1"], g["idx2"], g["pose1"], g["pose2"])
print("edge cases", np.bincount(r.status, minlength=5))
eng.close()
shapes, i1, i2, p1, p2 = W.config4_batch(3001, seed=2)
eng = d.ProximityEngine(shapes)
plan = eng.plan(i1, i2)
d1, d2 = torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda()
out = eng.solve(plan, d1, d2)
rec = torch.zeros((3001, 14), dtype=torch.float64, device="cuda")
eng.solve_records(plan, d1, d2, [rec.data_ptr()])
torch.cuda.synchronize()
print("config4", int((out.status != 0).sum()), float(rec[:, 0].sum()) == float(out.alpha[plan.perm().long()].sum()))
print("SANITIZE_RUN_OK")

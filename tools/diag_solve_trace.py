"""Diagnostic (GPU box): device start / end of every group kernel of one device-resident solve (DCOL_SOLVE_TRACE=1)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcol_trajectory_optimization_b200 as d
from dcol_trajectory_optimization_b200 import workloads as W
for lg in [int(a) for a in sys.argv[1:]] or [18, 20]:
    B = 1 << lg
    shapes, i1, i2, p1, p2 = W.config4_batch(B, seed=1)
    eng = d.ProximityEngine(shapes)
    dev = torch.device("cuda:0")
    plan = eng.plan(torch.as_tensor(i1, device=dev), torch.as_tensor(i2, device=dev))
    P1, P2 = torch.as_tensor(p1, device=dev), torch.as_tensor(p2, device=dev)
    out = eng.solve(plan, P1, P2, want_contact=False)
    for _ in range(3):
        eng.solve(plan, P1, P2, want_contact=False, out=out)
    torch.cuda.synchronize()
    os.environ["DCOL_SOLVE_TRACE"] = "1"
    print(f"=== B = 2^{lg}", file=sys.stderr, flush=True)
    eng.solve(plan, P1, P2, want_contact=False, out=out)
    torch.cuda.synchronize()
    del os.environ["DCOL_SOLVE_TRACE"]
    plan.close(); eng.close()

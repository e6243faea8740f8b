"""Diagnostic (GPU box): device-resident solve time against batch size and side streams (DCOL_SIDE_STREAMS is read once
per process).  usage: python tools/diag_small_solve.py [log2 sizes...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcol_trajectory_optimization_b200 as d
from dcol_trajectory_optimization_b200 import workloads as W

sizes = [int(a) for a in sys.argv[1:]] or [16, 18, 19, 20, 23]
for lg in sizes:
    B = 1 << lg
    shapes, i1, i2, p1, p2 = W.config4_batch(B, seed=1)
    eng = d.ProximityEngine(shapes)
    dev = torch.device("cuda:0")
    plan = eng.plan(torch.as_tensor(i1, device=dev), torch.as_tensor(i2, device=dev))
    P1, P2 = torch.as_tensor(p1, device=dev), torch.as_tensor(p2, device=dev)
    out = eng.solve(plan, P1, P2, want_contact=False)
    torch.cuda.synchronize()
    n = 20 if lg < 22 else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        eng.solve(plan, P1, P2, want_contact=False, out=out)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        eng.solve(plan, P1, P2, want_contact=False, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    # one solve at a time (host synchronises in between): no launch pipelining across solves
    import time
    t = time.perf_counter()
    for _ in range(n):
        eng.solve(plan, P1, P2, want_contact=False, out=out); torch.cuda.synchronize()
    ms1 = (time.perf_counter() - t) / n * 1e3
    print(f"side {os.environ.get('DCOL_SIDE_STREAMS','default')} conn {os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS','default')} "
          f"B 2^{lg}: back-to-back {ms:.3f} ms/solve ({B/ms/1e3:.0f} M pairs/s), synchronised {ms1:.3f} ms, max iters {int(out.iters.max())}", flush=True)
    plan.close(); eng.close()

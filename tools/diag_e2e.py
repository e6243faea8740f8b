"""Diagnostic (GPU box): PCIe copy rates and where the host entry point spends its time."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcol_trajectory_optimization_b200 as d
from dcol_trajectory_optimization_b200 import workloads as W
from dcol_trajectory_optimization_b200.shapes import flatten_shapes

n = 1 << 27
h = torch.empty(n, dtype=torch.uint8).pin_memory(); g = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("h2d", lambda: g.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(g, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); print(name, "torch pinned GB/s", 5 * n / (time.perf_counter() - t) / 1e9, flush=True)
hp = d.pinned_empty(n, np.uint8); ht = torch.from_numpy(hp)
g.copy_(ht, non_blocking=True); torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5): g.copy_(ht, non_blocking=True)
torch.cuda.synchronize(); print("h2d dcol_host_alloc GB/s", 5 * n / (time.perf_counter() - t) / 1e9, flush=True)

B = 1 << 23
shapes, i1, i2, p1, p2 = W.config4_batch(B, seed=1)
eng = d.ProximityEngine(shapes)
hi1, hi2 = d.pinned_empty(B, np.int32), d.pinned_empty(B, np.int32)
hp1, hp2 = d.pinned_empty((B, 6)), d.pinned_empty((B, 6))
hi1[:], hi2[:], hp1[:], hp2[:] = i1, i2, p1, p2
out = d.BatchResult(alpha=d.pinned_empty(B), contact=None, grad=d.pinned_empty((B, 12)),
                    iters=d.pinned_empty(B, np.int32), status=d.pinned_empty(B, np.int32))   # what bench.py's e2e leg moves
for noprio in ((True, False) if '--prio-ab' in sys.argv else (False,)):
    if noprio:
        os.environ["DCOL_HOST_NOPRIO"] = "1"
    else:
        os.environ.pop("DCOL_HOST_NOPRIO", None)
    eng = d.ProximityEngine(shapes)   # the host pipeline's streams are created per table
    for slots in ((2, 4) if '--slots-ab' in sys.argv else (int(os.environ.get('DCOL_HOST_SLOTS', 4)),)):
        os.environ["DCOL_HOST_SLOTS"] = str(slots)
        for chunk in ([int(sys.argv[sys.argv.index('--only-chunk') + 1])] if '--only-chunk' in sys.argv else (1 << 17, 1 << 18, 1 << 19, 1 << 20)):
            os.environ["DCOL_HOST_CHUNK"] = str(chunk)
            eng.solve_host(hi1, hi2, hp1, hp2, out=out)
            eng.solve_host(hi1, hi2, hp1, hp2, out=out)
            t = time.perf_counter()
            for _ in range(5):
                eng.solve_host(hi1, hi2, hp1, hp2, out=out)
            dt = (time.perf_counter() - t) / 5
            print(f"side streams {os.environ.get('DCOL_SIDE_STREAMS', 'default')} plan-stream priority {not noprio} slots {slots} chunk {chunk} pairs {B}: {dt*1e3:.2f} ms  {B/dt/1e6:.1f} Mpairs/s  {(B*216)/dt/1e9:.1f} GB/s both ways", flush=True)
del os.environ["DCOL_HOST_SLOTS"], os.environ["DCOL_HOST_CHUNK"]
# pageable
t = time.perf_counter(); r = eng.solve_host(i1, i2, p1, p2); dt = time.perf_counter() - t
print(f"pageable numpy buffers: {dt*1e3:.1f} ms", flush=True)

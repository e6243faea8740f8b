"""Diagnostic (GPU box): per-chunk device timeline of dcol_proximity_batch_host (DCOL_HOST_TRACE=1, printed by the library)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcol_trajectory_optimization_b200 as d
from dcol_trajectory_optimization_b200 import workloads as W

B = 1 << 23
shapes, i1, i2, p1, p2 = W.config4_batch(B, seed=1)
eng = d.ProximityEngine(shapes)
hi1, hi2 = d.pinned_empty(B, np.int32), d.pinned_empty(B, np.int32)
hp1, hp2 = d.pinned_empty((B, 6)), d.pinned_empty((B, 6))
hi1[:], hi2[:], hp1[:], hp2[:] = i1, i2, p1, p2
out = d.BatchResult(alpha=d.pinned_empty(B), contact=None, grad=d.pinned_empty((B, 12)),
                    iters=d.pinned_empty(B, np.int32), status=d.pinned_empty(B, np.int32))
for chunk in [int(a) for a in sys.argv[1:]] or [1 << 18, 1 << 20]:
    os.environ["DCOL_HOST_CHUNK"] = str(chunk)
    os.environ.pop("DCOL_HOST_TRACE", None)
    eng.solve_host(hi1, hi2, hp1, hp2, out=out)
    eng.solve_host(hi1, hi2, hp1, hp2, out=out)
    os.environ["DCOL_HOST_TRACE"] = "1"
    print(f"=== chunk {chunk}", file=sys.stderr, flush=True)
    eng.solve_host(hi1, hi2, hp1, hp2, out=out)

"""Large parity sweep on the GPU box: the CUDA path against the CPU oracle (pinned to the reference) on millions of
seeded pairs per workload; prints one JSON line per workload with mismatch counts and error quantiles."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import oracle as O  # noqa: E402
import dcol_trajectory_optimization_b200 as d  # noqa: E402
from dcol_trajectory_optimization_b200 import workloads as W  # noqa: E402
from dcol_trajectory_optimization_b200.shapes import flatten_shapes  # noqa: E402


def sweep(name, n):
    if name == "config4":
        shapes, i1, i2, p1, p2 = W.config4_batch(n, seed=777)
    else:
        shapes, i1, i2, p1, p2 = W.config5_batch(n_obs=1024, n_knots=100, n_cand=max(1, n // 102400), seed=778)
    rec, A, b = flatten_shapes(shapes)
    t = time.time()
    ref = O.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=O.GRAD_EXACT)
    alt = O.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=O.GRAD_EXACT, fma=True)
    t_cpu = time.time() - t
    eng = d.ProximityEngine((rec, A, b))
    res = eng.solve_host(i1, i2, p1, p2)
    eng.close()
    ok = (ref["status"] == 0) & (res.status == 0)
    aerr = np.abs(res.alpha - ref["alpha"])[ok] / np.maximum(np.abs(ref["alpha"][ok]), 1.0)
    gscale = np.abs(ref["grad"][ok]).max(axis=1)
    gerr = np.abs(res.grad[ok] - ref["grad"][ok]).max(axis=1) / gscale
    gsens = np.abs(alt["grad"][ok] - ref["grad"][ok]).max(axis=1) / gscale      # the reference's own rounding sensitivity
    flips = res.iters != ref["iters"]
    oflips = alt["iters"] != ref["iters"]
    return {"workload": name, "pairs": int(len(i1)), "status_mismatches": int((res.status != ref["status"]).sum()),
            "failed_pairs_reference": int((ref["status"] != 0).sum()),
            "iteration_count_mismatches": int(flips.sum()),
            "iteration_count_mismatches_oracle_fma_vs_oracle": int(oflips.sum()),
            "alpha_rel_err_max": float(aerr.max()), "alpha_rel_err_p999999": float(np.quantile(aerr, 0.999999)),
            "pairs_alpha_err_gt_1e-8": int((aerr > 1e-8).sum()),
            "grad_err_max": float(gerr.max()), "grad_err_median": float(np.median(gerr)),
            "pairs_grad_err_gt_1e-6": int((gerr > 1e-6).sum()),
            "pairs_where_reference_itself_moves_gt_1e-6_under_fma": int((gsens > 1e-6).sum()),
            "pairs_grad_err_gt_1e-6_not_explained_by_reference_sensitivity": int(((gerr > 1e-6) & (gsens < 1e-7)).sum()),
            "mean_iters": float(ref["iters"].mean()), "max_iters": int(ref["iters"].max()), "oracle_seconds": t_cpu}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    for name in ("config4", "config5"):
        print(json.dumps(sweep(name, n)), flush=True)

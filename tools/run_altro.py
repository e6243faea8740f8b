"""Run the three reference scenarios through the batched AL-iLQR caller on the GPU engine and print one JSON
line per scenario (wall time, passes, proximity problems solved, deviation from the reference trajectory)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcol_trajectory_optimization_b200.altro import PROBLEMS, altro_solve  # noqa: E402

REF_WALL = {"piano_mover": 104.0, "coneThroughWall": 300.0, "quadrotor": 2878.0}     # Report.pdf p.43 / .prof files
JULIA_WALL = {"piano_mover": 0.445, "coneThroughWall": 0.866, "quadrotor": 6.98}


def run(names=("piano_mover", "coneThroughWall", "quadrotor"), repeats=2):
    out = []
    for name in names:
        best = None
        for _ in range(repeats):                      # first run includes CUDA context / plan creation
            res = altro_solve(PROBLEMS[name]())
            if best is None or res.wall_s < best.wall_s:
                best = res
        line = {"scenario": name, "wall_s": best.wall_s, "passes": best.passes, "converged": best.converged,
                "pair_solves": best.pair_solves, "batched_calls": best.batched_calls,
                "host": "native core (libdcol_altro.so)" if "native" in PROBLEMS[name]().extra else "numpy",
                "timing_s": {k: round(v, 4) for k, v in best.timing.items()},
                "reference_python_wall_s": REF_WALL[name], "reference_julia_wall_s": JULIA_WALL[name]}
        gpath = os.path.join(ROOT, "tests", "golden", f"altro_{name}.npz")
        if os.path.exists(gpath):
            with np.load(gpath) as g:
                line["max_abs_dX_vs_reference"] = float(np.abs(best.X - g["X"]).max())
                line["max_abs_dU_vs_reference"] = float(np.abs(best.U - g["U"]).max())
                line["reference_passes"] = int(g["n_passes"]) - 1
        out.append(line)
    return out


if __name__ == "__main__":
    for line in run():
        print(json.dumps(line), flush=True)

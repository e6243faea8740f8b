"""Diagnostic (GPU box): repeat the three batched AL-iLQR solves and print wall time and phase timings of every repeat."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcol_trajectory_optimization_b200.altro import PROBLEMS, altro_solve  # noqa: E402
for name in ("piano_mover", "coneThroughWall", "quadrotor"):
    for rep in range(4):
        res = altro_solve(PROBLEMS[name]())
        print(name, rep, f"wall {res.wall_s:.4f} s passes {res.passes}", {k: round(v, 4) for k, v in res.timing.items()}, flush=True)

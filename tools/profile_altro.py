"""cProfile of the batched AL-iLQR caller on the GPU engine (where the host time of a pass goes)."""
import cProfile
import io
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcol_trajectory_optimization_b200.altro import PROBLEMS, altro_solve  # noqa: E402

for name in sys.argv[1:] or ("piano_mover", "quadrotor"):
    altro_solve(PROBLEMS[name]())            # warm-up: CUDA context, plans
    pr = cProfile.Profile()
    pr.enable()
    res = altro_solve(PROBLEMS[name]())
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print(f"==== {name}: wall {res.wall_s:.3f} s, {res.passes} passes")
    print(s.getvalue())

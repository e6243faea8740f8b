"""Diagnostic (GPU box): cost of creating / closing an engine around one scene call (what every altro_solve pays)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcol_trajectory_optimization_b200.altro import PROBLEMS
from dcol_trajectory_optimization_b200.altro.solver import EngineEvaluator
for name in ("piano_mover", "coneThroughWall", "quadrotor"):
    prob = PROBLEMS[name]()
    ts = []
    for rep in range(12):
        t0 = time.perf_counter(); ev = EngineEvaluator(prob); t1 = time.perf_counter()
        poses = np.zeros((80, 6)); poses[:, 0] = np.linspace(-20.0, -10.0, 80); poses[:, 2] = 30.0   # far from every obstacle
        t2 = time.perf_counter()
        ev.evaluate(poses, True)
        t3 = time.perf_counter(); ev.close(); t4 = time.perf_counter()
        ts.append((t1 - t0, t3 - t2, t4 - t3))
    print(name, "create / first call / close (ms):", " | ".join(f"{a*1e3:.1f} {b*1e3:.1f} {c*1e3:.1f}" for a, b, c in ts), flush=True)

"""Where the teardown time of an ALTRO solve goes: plan / engine destruction timed separately."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dcol_trajectory_optimization_b200.altro import PROBLEMS  # noqa: E402
from dcol_trajectory_optimization_b200.altro.solver import EngineEvaluator  # noqa: E402

for name in ("piano_mover", "quadrotor", "quadrotor"):
    prob = PROBLEMS[name]()
    t = time.perf_counter()
    ev = EngineEvaluator(prob)
    t_new = time.perf_counter() - t
    for M in (prob.N, 20 * prob.N):
        poses = prob.pose_of_state(np.repeat(prob.X0[:1], M, axis=0))
        t = time.perf_counter()
        ev(poses, True)
        t1 = time.perf_counter() - t
        t = time.perf_counter()
        ev(poses, True)
        print(f"{name} M={M}: first call {t1 * 1e3:.2f} ms, second {1e3 * (time.perf_counter() - t):.2f} ms")
    torch.cuda.synchronize()
    for M, (plan, _) in ev._plans.items():
        t = time.perf_counter()
        plan.close()
        print(f"{name}: plan({M}).close {1e3 * (time.perf_counter() - t):.2f} ms")
    t = time.perf_counter()
    ev.engine.close()
    print(f"{name}: engine new {t_new * 1e3:.2f} ms, engine.close {1e3 * (time.perf_counter() - t):.2f} ms")

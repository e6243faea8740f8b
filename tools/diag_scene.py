"""Diagnostic (GPU box): the scene entry point on the config-5 workload against the chunk size (DCOL_HOST_CHUNK)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcol_trajectory_optimization_b200 as d
from dcol_trajectory_optimization_b200 import workloads as Wl
from dcol_trajectory_optimization_b200.shapes import flatten_shapes

n_obs, n_knots, n_cand = 1024, 100, 81
rng = np.random.default_rng(2)
shapes = [Wl.SphereMRP(0.25)] + Wl.quadrotor_obstacle_shapes()
obs_pose = np.concatenate([rng.uniform([-8.0, -2.5, 1.0], [8.0, 2.5, 6.0], size=(n_obs, 3)), rng.normal(size=(n_obs, 3)) * 0.5], axis=1)
obs_shape = (1 + (np.arange(n_obs) % 11)).astype(np.int32)
knots = np.linspace([-8.0, 0.0, 4.0], [8.0, 0.0, 4.0], n_knots)
Ms = n_cand * n_knots
hv = d.pinned_empty((Ms, 6)); hv[:] = 0.0
hv[:, :3] = (knots[None] + rng.normal(size=(n_cand, n_knots, 3)) * 0.3).reshape(Ms, 3)
sout = d.SceneResult(alpha=d.pinned_empty((Ms, n_obs)), grad1=d.pinned_empty((Ms, n_obs, 6)),
                     iters=d.pinned_empty((Ms, n_obs), np.int32), status=d.pinned_empty((Ms, n_obs), np.int32))
for rep in range(2):
    for chunk in (1 << 18, 1 << 19, 1 << 20, 1 << 21):
        os.environ["DCOL_HOST_CHUNK"] = str(chunk)
        eng = d.ProximityEngine(flatten_shapes(shapes))
        for _ in range(2):
            eng.solve_scene_host(0, hv, obs_shape, obs_pose, out=sout)
        t = time.perf_counter()
        for _ in range(5):
            eng.solve_scene_host(0, hv, obs_shape, obs_pose, out=sout)
        dt = (time.perf_counter() - t) / 5
        print(f"scene chunk {chunk}: {dt*1e3:.2f} ms  {Ms*n_obs/dt/1e6:.1f} M pairs/s", flush=True)
        eng.close()

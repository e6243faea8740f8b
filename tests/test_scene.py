"""The scene form of the host entry point (``dcol_proximity_scene_host``): M victim poses x n_obs posed obstacles in
one call — the reference's per-knot, per-obstacle loops (systems/cluttered_hallway_quadrotor.py:127-133, 155-163) —
must return exactly what the pair-list entry point returns for the same pairs, through every chunking of the call."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(seed=3, M=257, n_obs=33):
    from dcol_trajectory_optimization_b200 import workloads as W
    from dcol_trajectory_optimization_b200.primitives import SphereMRP
    rng = np.random.default_rng(seed)
    shapes = [SphereMRP(0.25)] + W.quadrotor_obstacle_shapes()
    obs_shape = (1 + rng.integers(0, 11, size=n_obs)).astype(np.int32)
    obs_pose = np.concatenate([rng.uniform([-8, -2.5, 1], [8, 2.5, 6], size=(n_obs, 3)), 0.5 * rng.normal(size=(n_obs, 3))], axis=1)
    vic_pose = np.concatenate([rng.uniform([-8, -1, 3], [8, 1, 5], size=(M, 3)), 0.2 * rng.normal(size=(M, 3))], axis=1)
    return shapes, obs_shape, obs_pose, vic_pose


@pytest.mark.parametrize("chunk", [None, "2048"])
def test_scene_equals_pair_list(chunk):
    import dcol_trajectory_optimization_b200 as d
    shapes, obs_shape, obs_pose, vic_pose = _scene()
    M, n_obs = len(vic_pose), len(obs_shape)
    eng = d.ProximityEngine(shapes)
    i1 = np.zeros(M * n_obs, np.int32)
    i2 = np.tile(obs_shape, M)
    ref = eng.solve_host(i1, i2, np.repeat(vic_pose, n_obs, axis=0), np.tile(obs_pose, (M, 1)), want_contact=False)
    old = os.environ.get("DCOL_HOST_CHUNK")
    try:
        if chunk:
            os.environ["DCOL_HOST_CHUNK"] = chunk       # 62 victim poses per chunk: four full chunks and a shorter last one
        got = eng.solve_scene_host(0, vic_pose, obs_shape, obs_pose)
        light = eng.solve_scene_host(0, vic_pose, obs_shape, obs_pose, want_grad=False, want_iters=False)
    finally:
        if chunk:
            if old is None:
                del os.environ["DCOL_HOST_CHUNK"]
            else:
                os.environ["DCOL_HOST_CHUNK"] = old
    assert int((ref.status != 0).sum()) == 0
    assert np.array_equal(got.status.ravel(), ref.status) and np.array_equal(got.iters.ravel(), ref.iters)
    assert np.array_equal(got.alpha.ravel(), ref.alpha)                       # same kernels, same inputs: same bits
    assert np.array_equal(got.grad1.reshape(-1, 6), ref.grad[:, :6])
    assert light.grad1 is None and light.iters is None and np.array_equal(light.alpha, got.alpha)
    eng.close()


def test_scene_argument_errors_and_unsupported_pairs():
    import dcol_trajectory_optimization_b200 as d
    from dcol_trajectory_optimization_b200._lib import DcolError
    eng = d.ProximityEngine([d.CapsuleMRP(0.3, 1.2), d.CylinderMRP(0.4, 1.5), d.SphereMRP(0.5)])
    vic = np.zeros((3, 6))
    obs = np.array([[3.0, 0, 0, 0, 0, 0], [0, 4.0, 0, 0, 0, 0]])
    with pytest.raises(DcolError) as e:
        eng.solve_scene_host(7, vic, [1, 2], obs)
    assert e.value.code == -3
    with pytest.raises(DcolError):
        eng.solve_scene_host(0, vic, [1, 9], obs)
    r = eng.solve_scene_host(0, vic, [1, 2], obs)           # capsule x cylinder: the reference cannot assemble it
    assert np.all(r.status[:, 0] == 4) and np.all(np.isnan(r.alpha[:, 0])) and np.all(r.status[:, 1] == 0)
    f = eng.solve_scene_host(0, vic, [1, 2], obs, fix_case4=True)
    assert np.all(f.status == 0) and np.allclose(f.alpha[:, 1], r.alpha[:, 1], rtol=0, atol=0)
    empty = eng.solve_scene_host(0, np.zeros((0, 6)), [1, 2], obs)
    assert empty.alpha.shape == (0, 2)
    eng.close()

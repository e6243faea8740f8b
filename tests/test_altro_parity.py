"""Trajectory parity of the batched AL-iLQR caller (altro/solver.py) with the UNMODIFIED reference ALTRO.

Goldens: tests/golden/altro_<system>.npz, produced by oracle/gen_altro_golden.py running the reference's
main.py scenarios in the build container (call counts equal the authors' cProfile dumps: piano_mover 36,483
solves / 319,642 NT scalings, coneThroughWall 40,324 / 416,509).  Bars: identical number of iLQR passes,
states within 1e-6, controls within 1e-5 (the reference's gradient is a finite difference with ~1e-7 noise;
SURVEY.md section 7 measured 2.1e-6 / 3.5e-7 on U for an exact gradient).

CPU variant: the collision evaluator is injected (the oracle, as the checker) — the product's own evaluator
is the CUDA engine and is exercised by the gpu-marked variant.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN

EXPECTED_PASSES = {"piano_mover": 35, "coneThroughWall": 37, "quadrotor": 60}   # Report.pdf pp.34,38,41; .prof files


def _golden(name):
    path = os.path.join(GOLDEN, f"altro_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    with np.load(path) as d:
        return {k: d[k] for k in d.files}


def _oracle_evaluator(problem, oracle):
    from dcol_trajectory_optimization_b200.shapes import flatten_shapes, pose_of
    rec, A, b = flatten_shapes([problem.victim] + list(problem.obstacles))
    obs = np.stack([pose_of(o) for o in problem.obstacles])
    n = len(obs)

    def ev(poses, want_grad):
        M = poses.shape[0]
        r = oracle.solve_batch(rec, A, b, np.zeros(M * n, np.int32), np.tile(np.arange(1, n + 1, dtype=np.int32), M),
                               np.repeat(poses, n, axis=0), np.tile(obs, (M, 1)),
                               grad_mode=oracle.GRAD_EXACT if want_grad else oracle.GRAD_NONE)
        assert not r["status"].any()
        return r["alpha"].reshape(M, n), (r["grad"][:, :6].reshape(M, n, 6) if want_grad else None)
    return ev


def _check(name, res, g):
    assert res.converged
    assert res.passes == EXPECTED_PASSES[name] == int(g["n_passes"]) - 1
    assert np.abs(res.X - g["X"]).max() < 1e-6
    assert np.abs(res.U - g["U"]).max() < 1e-5
    # the batched caller solves fewer problems than the reference's scalar loops make calls... per pass it
    # evaluates 1 + max_linesearch_iters trajectories, all in two batched calls
    assert res.batched_calls == 2 * res.passes


@pytest.mark.parametrize("name", ["piano_mover", "coneThroughWall", "quadrotor"])
def test_altro_trajectory_parity_cpu(oracle, name):
    from dcol_trajectory_optimization_b200.altro import PROBLEMS, altro_solve
    g = _golden(name)
    problem = PROBLEMS[name]()
    res = altro_solve(problem, evaluator=_oracle_evaluator(problem, oracle))
    _check(name, res, g)


def test_numpy_host_core_still_matches_the_reference(oracle):
    """altro_solve(native=False): the NumPy implementation of the per-pass host work (user-defined dynamics)."""
    from dcol_trajectory_optimization_b200.altro import PROBLEMS, altro_solve
    g = _golden("piano_mover")
    problem = PROBLEMS["piano_mover"]()
    res = altro_solve(problem, evaluator=_oracle_evaluator(problem, oracle), native=False)
    _check("piano_mover", res, g)


def test_speculative_line_search_equals_sequential(oracle):
    from dcol_trajectory_optimization_b200.altro import PROBLEMS, altro_solve
    problem = PROBLEMS["piano_mover"]()
    a = altro_solve(problem, evaluator=_oracle_evaluator(problem, oracle), speculative=True)
    b = altro_solve(problem, evaluator=_oracle_evaluator(problem, oracle), speculative=False)
    assert a.passes == b.passes and [r[4] for r in a.log] == [r[4] for r in b.log]      # same step sizes accepted
    np.testing.assert_allclose(a.X, b.X, rtol=0, atol=1e-7)   # BLAS row-count dependent rounding, amplified by the iLQR passes
    assert b.batched_calls > a.batched_calls


def test_scenario_data_matches_reference_dumps():
    """data/scenes.npz (extracted from the reference's initialize_* functions) vs the goldens' X0/U0."""
    from dcol_trajectory_optimization_b200.altro import PROBLEMS
    for name in ("piano_mover", "coneThroughWall", "quadrotor"):
        g = _golden(name)
        p = PROBLEMS[name]()
        assert np.array_equal(p.X0, g["X0"]) and np.array_equal(p.U0, g["U0"])
        assert p.X0.shape == (p.N, p.nx) and p.U0.shape == (p.N - 1, p.nu)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["piano_mover", "coneThroughWall", "quadrotor"])
def test_altro_trajectory_parity_gpu(name):
    """The same check with the product's evaluator: every constraint through the CUDA engine."""
    from dcol_trajectory_optimization_b200.altro import PROBLEMS, altro_solve
    g = _golden(name)
    res = altro_solve(PROBLEMS[name]())
    _check(name, res, g)
    assert res.pair_solves == res.passes * 21 * PROBLEMS[name]().N * PROBLEMS[name]().n_obs

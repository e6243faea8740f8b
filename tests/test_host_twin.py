"""The product's device solver header (csrc/dcol_solver.cuh), compiled as plain C++ (tests/host_twin),
against the reference goldens and the oracle: proves on the CPU that the algorithm the CUDA threads run
— body-frame rows, scaled-space Newton step, closed-form cone scalings — follows the reference's
iterate path (identical iteration counts and status words, alpha to 1e-11)."""
import os
import sys

import numpy as np
import pytest

from conftest import load_golden

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_twin"))


@pytest.fixture(scope="module")
def twin():
    import twin as T
    T.build()
    return T


def _run(T, g):
    B = len(g["idx1"])
    out = dict(alpha=np.empty(B), iters=np.empty(B, np.int32), status=np.empty(B, np.int32),
               grad=np.empty((B, 12)), contact=np.empty((B, 3)))
    for tol in np.unique(g["tol"]):
        sel = np.where(g["tol"] == tol)[0]
        r = T.solve_batch(g["shape_records"], g["A"], g["b"], g["idx1"][sel], g["idx2"][sel], g["pose1"][sel],
                          g["pose2"][sel], tol=float(tol))
        for k in out:
            out[k][sel] = r[k]
    return out


@pytest.mark.parametrize("name", ["scenarios", "config4_sample", "config5_sample"])
def test_twin_matches_reference_golden(twin, name):
    g = load_golden(name)
    out = _run(twin, g)
    assert np.array_equal(out["status"], g["status"])
    assert np.array_equal(out["iters"], g["iters"])
    rel = np.abs(out["alpha"] - g["alpha"]) / np.maximum(np.abs(g["alpha"]), 1.0)
    assert rel.max() < 1e-11
    gerr = np.abs(out["grad"] - g["grad"]).max(axis=1) / np.abs(g["grad"]).max(axis=1)
    assert gerr.max() < 1e-6          # analytic gradient vs the reference's finite differences
    scale = np.maximum(np.abs(g["x"][:, :3]).max(axis=1), 1.0)
    assert (np.abs(out["contact"] - g["x"][:, :3]).max(axis=1) / scale).max() < 1e-9


def test_twin_edge_cases(twin):
    g = load_golden("edge_cases")
    out = _run(twin, g)
    tag = np.array([str(t) for t in g["tag"]])
    with np.errstate(invalid="ignore"):
        ties = (np.abs(g["mu"] - g["tol"][:, None]) <= 1e-9 * g["tol"][:, None]).any(axis=1)
    stable = ~np.isin(tag, ["tol0", "sep1e+09"]) & ~ties
    assert np.array_equal(out["status"][stable], g["status"][stable])
    assert np.array_equal(out["iters"][stable], g["iters"][stable])
    assert np.all(out["status"][tag == "case4"] == 4)
    assert np.all(out["status"][np.isin(tag, ["nan_r", "nan_p", "inf_r"])] == 2)
    assert np.all(out["status"][tag == "tol0"] != 0)
    ok = stable & (g["status"] == 0)
    rel = np.abs(out["alpha"][ok] - g["alpha"][ok]) / np.maximum(np.abs(g["alpha"][ok]), 1.0)
    assert rel.max() < 1e-11


def test_twin_analytic_gradient_matches_oracle_exact_gradient(twin, oracle):
    """Two independent routes to the same derivative (closed forms in the kernel header, differences of
    the affine block builder in the oracle) on every supported type pair."""
    from dcol_trajectory_optimization_b200 import workloads as W
    from dcol_trajectory_optimization_b200.shapes import flatten_shapes
    shapes, i1, i2, p1, p2 = W.config4_batch(20_000, seed=77)
    rec, A, b = flatten_shapes(shapes)
    ref = oracle.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=oracle.GRAD_EXACT)
    out = twin.solve_batch(rec, A, b, i1, i2, p1, p2)
    assert np.array_equal(out["status"], ref["status"]) and np.array_equal(out["iters"], ref["iters"])
    gerr = np.abs(out["grad"] - ref["grad"]).max(axis=1) / np.abs(ref["grad"]).max(axis=1)
    assert gerr.max() < 1e-7 and np.median(gerr) < 1e-10


def test_twin_gradient_error_tail_on_the_scene_workload(twin, oracle):
    """Guards the ROUNDING-level closeness to the reference's iterate path (profiles/r02_arithmetic_parity.md): identities
    that are exact in exact arithmetic can still fatten the gradient-error tail.  The rejected `s^-1` form of the cone
    centring term puts ~40 pairs per million of this workload above 1e-7 (largest 1.3e-6); the kernel's arithmetic has none
    in 8 million (largest 4e-8)."""
    from dcol_trajectory_optimization_b200 import workloads as W
    from dcol_trajectory_optimization_b200.shapes import flatten_shapes
    shapes, i1, i2, p1, p2 = W.config5_batch(n_obs=1024, n_knots=100, n_cand=2, seed=778)       # 204,800 pairs
    rec, A, b = flatten_shapes(shapes)
    ref = oracle.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=oracle.GRAD_EXACT)
    out = twin.solve_batch(rec, A, b, i1, i2, p1, p2)
    assert np.array_equal(out["status"], ref["status"]) and np.array_equal(out["iters"], ref["iters"])
    gerr = np.abs(out["grad"] - ref["grad"]).max(axis=1) / np.abs(ref["grad"]).max(axis=1)
    aerr = np.abs(out["alpha"] - ref["alpha"]) / np.maximum(np.abs(ref["alpha"]), 1.0)
    assert gerr.max() < 1e-7 and int((gerr > 3e-8).sum()) <= 2, (gerr.max(), int((gerr > 3e-8).sum()))
    assert aerr.max() < 1e-10


def test_twin_trace_world_frame_sz(twin):
    """(x, s, z) exported in the reference's world-frame row order, and the mu trace."""
    g = load_golden("scenarios")
    for k in range(len(g["idx1"])):
        r = twin.solve_pair(g["shape_records"], g["A"], g["b"], g["idx1"][k], g["idx2"][k], g["pose1"][k],
                            g["pose2"][k])
        n, m, it = int(g["n"][k]), int(g["m"][k]), int(g["iters"][k])
        assert (r["n"], r["m"], r["iters"], r["status"]) == (n, m, it, 0)
        np.testing.assert_allclose(r["x"], g["x"][k, :n], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(r["s"], g["s"][k, :m], rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(r["z"], g["z"][k, :m], rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(r["mu"][:it + 1], g["mu"][k, :it + 1], rtol=1e-5)

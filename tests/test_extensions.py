"""EXTENSIONS beyond the reference's code — parity UNPINNED: the reference
raises ValueError for pairs in which both primitives carry extra variables (combine_problem_matrices.py:58-67)
and has no ellipsoid.  Checked against the oracle's equally extended restatement (same algorithm, independent
code), closed forms, and symmetry."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_twin"))


def _all_pairs_batch(n=16_200, seed=9):
    import dcol_trajectory_optimization_b200 as d
    from dcol_trajectory_optimization_b200 import workloads as W
    from dcol_trajectory_optimization_b200.shapes import flatten_shapes
    shapes = W.config4_shapes() + [d.EllipsoidMRP(0.7, 0.4, 0.25), d.EllipsoidMRP(0.5, 0.5, 0.5)]
    ns = len(shapes)
    pairs = np.array([(i, j) for i in range(ns) for j in range(ns)], dtype=np.int32)      # all 81 ordered pairs
    sel = np.arange(n) % len(pairs)
    p1, p2 = W.config4_poses(n, seed=seed)
    return flatten_shapes(shapes), pairs[sel, 0].copy(), pairs[sel, 1].copy(), p1, p2


def _compare(res, ref, kinds, i1, i2):
    assert np.array_equal(res["status"], ref["status"]) and int(ref["status"].sum()) == 0
    assert np.array_equal(res["iters"], ref["iters"])
    assert (np.abs(res["alpha"] - ref["alpha"]) / np.maximum(np.abs(ref["alpha"]), 1.0)).max() < 1e-8
    gerr = np.abs(res["grad"] - ref["grad"]).max(axis=1) / np.abs(ref["grad"]).max(axis=1)
    assert np.quantile(gerr, 0.999) < 1e-6 and gerr.max() < 1e-3
    case4 = np.isin(kinds[i1], [1, 2, 5]) & np.isin(kinds[i2], [1, 2, 5])
    assert case4.sum() > 1000


def test_twin_all_81_type_pairs_vs_extended_oracle(oracle):
    import twin as T
    T.build()
    (rec, A, b), i1, i2, p1, p2 = _all_pairs_batch()
    ref = oracle.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=oracle.GRAD_EXACT, fix_case4=True)
    res = T.solve_batch(rec, A, b, i1, i2, p1, p2, fix_case4=True)
    _compare(res, ref, rec["type"], i1, i2)
    # parity mode (flag off): the reference's ValueError -> status 4, exactly on the both-extras pairs
    off = T.solve_batch(rec, A, b, i1, i2, p1, p2)
    case4 = np.isin(rec["type"][i1], [1, 2, 5]) & np.isin(rec["type"][i2], [1, 2, 5])
    assert np.array_equal(off["status"] == 4, case4)


@pytest.mark.gpu
def test_cuda_all_81_type_pairs_vs_extended_oracle(oracle):
    import dcol_trajectory_optimization_b200 as d
    (rec, A, b), i1, i2, p1, p2 = _all_pairs_batch()
    ref = oracle.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=oracle.GRAD_EXACT, fix_case4=True)
    eng = d.ProximityEngine((rec, A, b))
    r = eng.solve_host(i1, i2, p1, p2, fix_case4=True)
    _compare(dict(status=r.status, iters=r.iters, alpha=r.alpha, grad=r.grad), ref, rec["type"], i1, i2)
    off = eng.solve_host(i1, i2, p1, p2)
    case4 = np.isin(rec["type"][i1], [1, 2, 5]) & np.isin(rec["type"][i2], [1, 2, 5])
    assert np.array_equal(off.status == 4, case4) and np.all(np.isnan(off.alpha[case4]))
    eng.close()


@pytest.mark.gpu
def test_extension_closed_forms():
    import dcol_trajectory_optimization_b200 as d
    sph, ell = d.SphereMRP(0.5), d.EllipsoidMRP(0.7, 0.4, 0.25)
    c1, c2 = d.CapsuleMRP(0.3, 1.2), d.CapsuleMRP(0.2, 2.0)
    eng = d.ProximityEngine([sph, ell, c1, c2])
    z = np.zeros(3)
    # sphere vs axis-aligned ellipsoid on its x / y / z axis: alpha (R + axis) = distance
    for axis, semi in enumerate((0.7, 0.4, 0.25)):
        r2 = np.zeros(3)
        r2[axis] = 3.0
        r = eng.solve_host([0], [1], np.concatenate([z, z])[None], np.concatenate([r2, z])[None])
        assert r.status[0] == 0 and abs(r.alpha[0] - 3.0 / (0.5 + semi)) / r.alpha[0] < 3e-5
    # two parallel capsules (axes along x) side by side at distance d: alpha (R1 + R2) = d; symmetric in the order
    p = np.concatenate([np.array([0.0, 2.5, 0.0]), z])[None]
    o = np.concatenate([z, z])[None]
    a = eng.solve_host([2], [3], o, p, fix_case4=True)
    bb = eng.solve_host([3], [2], p, o, fix_case4=True)
    assert a.status[0] == 0 and bb.status[0] == 0
    assert abs(a.alpha[0] - 2.5 / 0.5) / a.alpha[0] < 3e-5 and abs(a.alpha[0] - bb.alpha[0]) / a.alpha[0] < 3e-5
    # the gradient of the symmetric problem: d alpha / d r1 = -d alpha / d r2 = -(unit y) / (R1 + R2)
    assert np.abs(a.grad[0, 0:3] - np.array([0, -2.0, 0])).max() < 1e-3
    assert np.abs(a.grad[0, 0:3] + a.grad[0, 6:9]).max() < 1e-6
    eng.close()


# ---- pinned: goldens produced by the REFERENCE'S OWN solve_lp_pdip / finite-difference gradient, fed by its own
# per-primitive problem_matrices through an assembly patched in exactly two places (oracle/gen_golden_extensions.py):
# case 4 with the first primitive's blocks padded (combine_problem_matrices.py:58-67 as intended) and a hand-assembled
# ellipsoid SOC block (Report.pdf section 3.1.5 eq. 27).  41 type pairs + offsets, 2,400 pairs.
def _check_against_reference_goldens(res):
    from conftest import load_golden
    g = load_golden("extensions")
    assert int(g["status"].sum()) == 0 and int(g["case4"].sum()) >= 360
    assert np.array_equal(res["status"], g["status"])
    assert np.array_equal(res["iters"], g["iters"])                                  # identical iteration counts
    a_err = np.abs(res["alpha"] - g["alpha"]) / np.maximum(np.abs(g["alpha"]), 1.0)
    assert a_err.max() < 1e-8, a_err.max()
    g_err = np.abs(res["grad"] - g["grad"]).max(axis=1) / np.abs(g["grad"]).max(axis=1)
    assert g_err.max() < 1e-6, g_err.max()                                           # the reference's FD values
    return g


def _solve_extension_goldens(solve):
    from conftest import load_golden
    g = load_golden("extensions")
    return solve((g["shape_records"], g["A"], g["b"]), g["idx1"], g["idx2"], g["pose1"], g["pose2"])


def test_oracle_extensions_vs_reference_solver(oracle):
    r = _solve_extension_goldens(lambda t, i1, i2, p1, p2: oracle.solve_batch(*t, i1, i2, p1, p2, grad_mode=oracle.GRAD_FD,
                                                                              fix_case4=True))
    _check_against_reference_goldens(r)


def test_twin_extensions_vs_reference_solver():
    import twin as T
    T.build()
    r = _solve_extension_goldens(lambda t, i1, i2, p1, p2: T.solve_batch(*t, i1, i2, p1, p2, fix_case4=True))
    _check_against_reference_goldens(r)


@pytest.mark.gpu
def test_cuda_extensions_vs_reference_solver():
    import dcol_trajectory_optimization_b200 as d

    def solve(t, i1, i2, p1, p2):
        eng = d.ProximityEngine(t)
        r = eng.solve_host(i1, i2, p1, p2, fix_case4=True)
        eng.close()
        return dict(status=r.status, iters=r.iters, alpha=r.alpha, grad=r.grad)
    _check_against_reference_goldens(_solve_extension_goldens(solve))

"""Solution Jacobian d(contact point, alpha) / d[r1 p1 r2 p2].

The reference has no such function (it only differentiates the frozen-(x, z) Lagrangian,
proximity/proximity_gradient.py:8-88), so parity here is UNPINNED; the checks are
  * a dense NumPy restatement of the same KKT differentiation, built on the oracle's assembly of
    (G, h) (the reference's problem_matrices / combine_problem_matrices) and central differences of it,
  * the ground truth: central differences of contact points / alpha of 1e-12 solves,
  * closed forms (sphere-sphere) and the envelope theorem (alpha row = the reference's gradient),
on the CPU through the host twin of the device header, and on the GPU through the C ABI
(``dcol_proximity_batch_jacobian``) against the twin and the same closed forms.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "host_twin"))

import oracle as O  # noqa: E402
import twin as T  # noqa: E402
from dcol_trajectory_optimization_b200 import workloads as W  # noqa: E402
from dcol_trajectory_optimization_b200.shapes import flatten_shapes  # noqa: E402


def _soc_w2_inv(s, z):
    """W^-2 of one second-order cone with the NT direction wbar (NT_scaling.py:340-405) and the scale s_0/z_0
    (the symmetric linearisation of s o z = const that is exact on the cone's tangent plane, dcol_solver.cuh)."""
    J = lambda u: u[0] ** 2 - u[1:] @ u[1:]  # noqa: E731
    sb, zb = s / np.sqrt(J(s)), z / np.sqrt(J(z))
    gamma = np.sqrt((1.0 + sb @ zb) / 2.0)
    Jm = np.diag([1.0] + [-1.0] * (len(s) - 1))
    w = (sb + Jm @ zb) / (2.0 * gamma)
    wh = Jm @ w
    return (z[0] / s[0]) * (2.0 * np.outer(wh, wh) - Jm)


def _arw(u):
    M = u[0] * np.eye(len(u))
    M[0, 1:] = u[1:]
    M[1:, 0] = u[1:]
    return M


def dense_jacobian(rec, A, b, i1, i2, pose1, pose2, x, s, z, exact_arw=False, h=1e-6):
    """dx = -(G^T H G)^-1 (dG^T z + G^T H (dG x - dh)), H = W^-2, with dG, dh by central differences of the
    oracle's assembly.  ``exact_arw``: H = Arw(s)^-1 Arw(z) on the cones (unsymmetric exact linearisation)."""
    _, _, G, _, dims = O.assemble(rec, A, b, i1, i2, pose1, pose2)
    n, m, no, q1, q2 = dims
    H = np.zeros((m, m))
    H[:no, :no] = np.diag(z[:no] / s[:no])
    o = no
    for q in (q1, q2):
        if q > 0:
            sq, zq = s[o:o + q], z[o:o + q]
            H[o:o + q, o:o + q] = np.linalg.inv(_arw(sq)) @ _arw(zq) if exact_arw else _soc_w2_inv(sq, zq)
            o += q
    M = G.T @ H @ G
    th = np.concatenate([pose1, pose2])
    Jm = np.zeros((n, 12))
    for j in range(12):
        tp, tm = th.copy(), th.copy()
        tp[j] += h
        tm[j] -= h
        _, _, Gp, hp, _ = O.assemble(rec, A, b, i1, i2, tp[:6], tp[6:])
        _, _, Gm, hm, _ = O.assemble(rec, A, b, i1, i2, tm[:6], tm[6:])
        dG, dh = (Gp - Gm) / (2 * h), (hp - hm) / (2 * h)
        Jm[:, j] = np.linalg.solve(M, -(dG.T @ z + G.T @ H @ (dG @ x - dh)))
    return Jm[:4]


def fd_truth(rec, A, b, i1, i2, p1, p2, h=1e-6, tol=1e-12):
    """Central differences of (contact, alpha) of tightly converged solves, [B, 4, 12]; NaN where a solve failed."""
    B = len(i1)
    out = np.zeros((B, 4, 12))
    th = np.concatenate([p1, p2], axis=1)
    for j in range(12):
        tp, tm = th.copy(), th.copy()
        tp[:, j] += h
        tm[:, j] -= h
        rp = T.solve_batch(rec, A, b, i1, i2, tp[:, :6].copy(), tp[:, 6:].copy(), tol=tol, want_grad=False)
        rm = T.solve_batch(rec, A, b, i1, i2, tm[:, :6].copy(), tm[:, 6:].copy(), tol=tol, want_grad=False)
        out[:, :3, j] = (rp["contact"] - rm["contact"]) / (2 * h)
        out[:, 3, j] = (rp["alpha"] - rm["alpha"]) / (2 * h)
    return out


def _rel(J, Jref):
    J, Jref = J.reshape(len(J), -1), Jref.reshape(len(Jref), -1)
    return np.abs(J - Jref).max(axis=1) / np.maximum(1.0, np.abs(Jref).max(axis=1))


@pytest.fixture(scope="module")
def batch():
    shapes, i1, i2, p1, p2 = W.config4_batch(240, seed=5)     # 6 pairs of each of the 40 type pairs
    rec, A, b = flatten_shapes(shapes)
    return rec, A, b, i1, i2, p1, p2


def test_twin_matches_dense_kkt_differentiation(batch):
    rec, A, b, i1, i2, p1, p2 = batch
    out = T.solve_batch(rec, A, b, i1, i2, p1, p2, want_jac=True)
    assert (out["status"] == 0).all() and np.isfinite(out["jac"]).all()
    errs = []
    for k in range(len(i1)):
        r = T.solve_pair(rec, A, b, i1[k], i2[k], p1[k], p2[k])
        Jd = dense_jacobian(rec, A, b, i1[k], i2[k], p1[k], p2[k], r["x"], r["s"], r["z"])
        errs.append(np.abs(out["jac"][k] - Jd).max() / max(1.0, np.abs(Jd).max()))
    errs = np.array(errs)
    # the residue is the dense check's own finite differences and the conditioning of M (~1/mu)
    assert np.median(errs) < 1e-8 and errs.max() < 5e-6, (np.median(errs), errs.max())


def test_alpha_row_is_the_reference_gradient(batch):
    """Envelope theorem: d alpha/d theta through the KKT system = gradient of the frozen Lagrangian + O(mu)."""
    rec, A, b, i1, i2, p1, p2 = batch
    out = T.solve_batch(rec, A, b, i1, i2, p1, p2, want_jac=True)
    err = np.abs(out["jac"][:, 3, :] - out["grad"]).max(axis=1) / np.abs(out["grad"]).max(axis=1)
    # O(mu) plus the iterate's primal / dual residuals, which the reference never tests (pdip.py:418-422)
    assert np.median(err) < 1e-4 and err.max() < 5e-2, (np.median(err), err.max())


def test_sphere_sphere_closed_form():
    """Contact of two scaled spheres: x = c1 + R1/(R1+R2) (c2 - c1), alpha = |c2 - c1| / (R1 + R2)."""
    from dcol_trajectory_optimization_b200.primitives import SphereMRP
    R1, R2 = 0.5, 1.25
    rec, A, b = flatten_shapes([SphereMRP(R1), SphereMRP(R2)])
    rng = np.random.default_rng(3)
    B = 64
    p1 = np.concatenate([rng.normal(size=(B, 3)), rng.normal(size=(B, 3)) * 0.5], axis=1)
    p2 = np.concatenate([rng.normal(size=(B, 3)) + 4.0, rng.normal(size=(B, 3)) * 0.5], axis=1)
    i1, i2 = np.zeros(B, np.int32), np.ones(B, np.int32)
    out = T.solve_batch(rec, A, b, i1, i2, p1, p2, want_jac=True)
    J = out["jac"]
    d = p2[:, :3] - p1[:, :3]
    n = d / np.linalg.norm(d, axis=1, keepdims=True)
    eye = np.broadcast_to(np.eye(3), (B, 3, 3))
    assert np.abs(J[:, :3, 0:3] - R2 / (R1 + R2) * eye).max() < 1e-5
    assert np.abs(J[:, :3, 6:9] - R1 / (R1 + R2) * eye).max() < 1e-5
    assert np.abs(J[:, 3, 0:3] + n / (R1 + R2)).max() < 1e-5 and np.abs(J[:, 3, 6:9] - n / (R1 + R2)).max() < 1e-5
    assert np.abs(J[:, :, 3:6]).max() < 1e-9 and np.abs(J[:, :, 9:12]).max() < 1e-9     # spheres do not see rotations


def test_against_finite_differences_of_tight_solves():
    """Ground truth: the derivative of the solution map itself.  The relaxed (mu ~ 1e-6) Jacobian smooths
    weakly active constraints, so a few near-degenerate contacts differ; the bulk must agree."""
    shapes, i1, i2, p1, p2 = W.config4_batch(120, seed=9)
    rec, A, b = flatten_shapes(shapes)
    Jfd = fd_truth(rec, A, b, i1, i2, p1, p2)
    ok = np.isfinite(Jfd).reshape(len(i1), -1).all(axis=1)
    assert ok.sum() >= 115
    for tol, med, p90 in ((1e-6, 5e-4, 1e-2), (1e-9, 5e-5, 2e-3)):
        out = T.solve_batch(rec, A, b, i1, i2, p1, p2, tol=tol, want_jac=True)
        err = _rel(out["jac"][ok], Jfd[ok])
        assert np.median(err) < med and np.quantile(err, 0.9) < p90, (tol, np.median(err), np.quantile(err, 0.9))
    # the symmetric scale s_0/z_0 is as good as the unsymmetric exact linearisation of s o z = const
    e_sym, e_arw = [], []
    for k in np.flatnonzero(ok)[:60]:
        r = T.solve_pair(rec, A, b, i1[k], i2[k], p1[k], p2[k])
        for exact, acc in ((False, e_sym), (True, e_arw)):
            Jd = dense_jacobian(rec, A, b, i1[k], i2[k], p1[k], p2[k], r["x"], r["s"], r["z"], exact_arw=exact)
            acc.append(np.abs(Jd - Jfd[k]).max() / max(1.0, np.abs(Jfd[k]).max()))
    assert np.median(e_sym) < 2.0 * np.median(e_arw) + 1e-6, (np.median(e_sym), np.median(e_arw))


def test_extension_pairs_have_jacobians():
    """Both-extras pairs (DCOL_FIX_CASE4) and the ellipsoid go through the same code."""
    from dcol_trajectory_optimization_b200.primitives import CapsuleMRP, CylinderMRP, EllipsoidMRP
    rec, A, b = flatten_shapes([CapsuleMRP(0.3, 1.2), CylinderMRP(0.4, 1.5), EllipsoidMRP(0.5, 0.8, 1.1)])
    rng = np.random.default_rng(11)
    B = 36
    i1 = np.repeat(np.arange(3), 12).astype(np.int32)
    i2 = np.tile(np.arange(3), 12).astype(np.int32)
    p1 = np.concatenate([rng.normal(size=(B, 3)) * 0.3, rng.normal(size=(B, 3)) * 0.5], axis=1)
    p2 = np.concatenate([rng.normal(size=(B, 3)) * 0.3 + 3.0, rng.normal(size=(B, 3)) * 0.5], axis=1)
    out = T.solve_batch(rec, A, b, i1, i2, p1, p2, want_jac=True, fix_case4=True)
    assert (out["status"] == 0).all() and np.isfinite(out["jac"]).all()
    err = np.abs(out["jac"][:, 3, :] - out["grad"]).max(axis=1) / np.abs(out["grad"]).max(axis=1)
    assert err.max() < 5e-2
    # translating both primitives together translates the contact point and leaves alpha alone
    s = out["jac"][:, :, 0:3] + out["jac"][:, :, 6:9]
    assert np.abs(s[:, :3] - np.eye(3)).max() < 1e-6 and np.abs(s[:, 3]).max() < 1e-6


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_jacobian_matches_twin_and_plain_solve():
    import torch
    import dcol_trajectory_optimization_b200 as d
    shapes, i1, i2, p1, p2 = W.config4_batch(8000, seed=21)
    rec, A, b = flatten_shapes(shapes)
    eng = d.ProximityEngine((rec, A, b), device=0)
    plan = eng.plan(i1, i2)
    P1 = torch.as_tensor(p1, device="cuda")
    P2 = torch.as_tensor(p2, device="cuda")
    plain = eng.solve(plan, P1, P2)
    res = eng.solve(plan, P1, P2, want_jac=True)
    torch.cuda.synchronize()
    ref = T.solve_batch(rec, A, b, i1, i2, p1, p2, want_jac=True)
    st, it = res.status.cpu().numpy(), res.iters.cpu().numpy()
    assert np.array_equal(st, ref["status"]) and np.array_equal(it, ref["iters"])
    assert np.array_equal(st, plain.status.cpu().numpy()) and np.array_equal(it, plain.iters.cpu().numpy())
    assert np.abs(res.alpha.cpu().numpy() - plain.alpha.cpu().numpy()).max() < 1e-12
    assert np.abs(res.grad.cpu().numpy() - plain.grad.cpu().numpy()).max() < 1e-9
    J = res.jac.cpu().numpy()
    assert np.isfinite(J).all()
    err = _rel(J, ref["jac"])
    # same algorithm, different rounding (FMA contraction, MUFU-seeded reciprocals), amplified by cond(M) ~ 1/mu
    assert np.median(err) < 1e-9 and np.quantile(err, 0.99) < 1e-6 and err.max() < 1e-3, \
        (np.median(err), np.quantile(err, 0.99), err.max())
    # envelope theorem on the device results
    g = res.grad.cpu().numpy()
    e = np.abs(J[:, 3, :] - g).max(axis=1) / np.abs(g).max(axis=1)
    assert np.median(e) < 1e-4
    eng.close()


@pytest.mark.gpu
def test_gpu_jacobian_unsupported_and_failed_pairs_are_nan():
    import torch
    import dcol_trajectory_optimization_b200 as d
    from dcol_trajectory_optimization_b200.primitives import CapsuleMRP, SphereMRP
    rec, A, b = flatten_shapes([CapsuleMRP(0.3, 1.2), SphereMRP(0.5)])
    eng = d.ProximityEngine((rec, A, b), device=0)
    i1 = np.array([0, 0, 1], np.int32)
    i2 = np.array([0, 1, 1], np.int32)
    p1 = np.zeros((3, 6))
    p2 = np.zeros((3, 6))
    p2[:, 0] = 3.0
    plan = eng.plan(i1, i2)
    res = eng.solve(plan, torch.as_tensor(p1, device="cuda"), torch.as_tensor(p2, device="cuda"), want_jac=True)
    torch.cuda.synchronize()
    st = res.status.cpu().numpy()
    J = res.jac.cpu().numpy()
    assert st[0] == 4 and np.isnan(J[0]).all()          # capsule x capsule without DCOL_FIX_CASE4
    assert st[1] == 0 and np.isfinite(J[1]).all()
    assert st[2] == 0 and np.abs(J[2, :3, 0:3] - 0.5 * np.eye(3)).max() < 1e-5
    res = eng.solve(plan, torch.as_tensor(p1, device="cuda"), torch.as_tensor(p2, device="cuda"), want_jac=True, max_iter=2)
    torch.cuda.synchronize()
    assert (res.status.cpu().numpy()[1:] == 1).all() and np.isnan(res.jac.cpu().numpy()[1:]).all()
    eng.close()


@pytest.mark.gpu
def test_gpu_jacobian_vs_dense_kkt_differentiation_and_finite_differences():
    """The CUDA Jacobian against checks that share no code with the kernel (nor with its host twin):
    (1) the dense NumPy KKT differentiation above, evaluated at the iterate (x, s, z) the GPU itself returned
        (``dcol_debug_trace_pair``) on the oracle's assembly of (G, h) — all 40 type pairs;
    (2) central differences of contact points / alpha of tol = 1e-12 solves run on the GPU."""
    import torch
    import dcol_trajectory_optimization_b200 as d
    shapes, i1, i2, p1, p2 = W.config4_batch(240, seed=5)
    rec, A, b = flatten_shapes(shapes)
    eng = d.ProximityEngine((rec, A, b), device=0)
    plan = eng.plan(i1, i2)
    res = eng.solve(plan, torch.as_tensor(p1, device="cuda"), torch.as_tensor(p2, device="cuda"), want_jac=True)
    torch.cuda.synchronize()
    J = res.jac.cpu().numpy()
    assert (res.status.cpu().numpy() == 0).all() and np.isfinite(J).all()
    errs = []
    for k in range(len(i1)):
        t = eng.trace_pair(i1[k], i2[k], p1[k], p2[k])
        assert t["status"] == 0
        Jd = dense_jacobian(rec, A, b, i1[k], i2[k], p1[k], p2[k], t["x"], t["s"], t["z"])
        errs.append(np.abs(J[k] - Jd).max() / max(1.0, np.abs(Jd).max()))
    errs = np.array(errs)
    assert np.median(errs) < 1e-8 and errs.max() < 5e-6, (np.median(errs), errs.max())

    # (2) finite differences of tight GPU solves
    n, h = 120, 1e-6
    th = np.concatenate([p1[:n], p2[:n]], axis=1)
    Jfd = np.zeros((n, 4, 12))
    for j in range(12):
        tp, tm = th.copy(), th.copy()
        tp[:, j] += h
        tm[:, j] -= h
        rp = eng.solve_host(i1[:n], i2[:n], tp[:, :6].copy(), tp[:, 6:].copy(), tol=1e-12, want_grad=False)
        rm = eng.solve_host(i1[:n], i2[:n], tm[:, :6].copy(), tm[:, 6:].copy(), tol=1e-12, want_grad=False)
        Jfd[:, :3, j] = (rp.contact - rm.contact) / (2 * h)
        Jfd[:, 3, j] = (rp.alpha - rm.alpha) / (2 * h)
    ok = np.isfinite(Jfd).reshape(n, -1).all(axis=1)
    assert ok.sum() >= 115
    for tol, med, p90 in ((1e-6, 5e-4, 1e-2), (1e-9, 5e-5, 2e-3)):
        out = eng.solve(plan, torch.as_tensor(p1, device="cuda"), torch.as_tensor(p2, device="cuda"), tol=tol, want_jac=True)
        torch.cuda.synchronize()
        err = _rel(out.jac.cpu().numpy()[:n][ok], Jfd[ok])
        assert np.median(err) < med and np.quantile(err, 0.9) < p90, (tol, np.median(err), np.quantile(err, 0.9))
    plan.close()
    eng.close()

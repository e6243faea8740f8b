"""The native host core of the AL-iLQR caller (csrc/altro_core.cpp, include/dcol_altro.h) against the NumPy
implementation of the same functions in altro/solver.py (which has trajectory parity with the unmodified
reference ALTRO, tests/test_altro_parity.py): dynamics, RK4, rollouts, forward-difference Jacobians, Riccati
sweep, augmented-Lagrangian cost — on random trajectories of the reference's three systems."""
import os

import numpy as np
import pytest

from dcol_trajectory_optimization_b200.altro import PROBLEMS
from dcol_trajectory_optimization_b200.altro import native as NV
from dcol_trajectory_optimization_b200.altro import solver as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["piano_mover", "coneThroughWall", "quadrotor"]


def test_library_exports_every_declared_symbol():
    import re
    NV.build()
    header = open(os.path.join(ROOT, "include", "dcol_altro.h")).read()
    declared = set(re.findall(r"\b(dcol_altro_[a-z0-9_]+)\s*\(", header))
    assert declared == set(NV.SYMBOLS), declared ^ set(NV.SYMBOLS)
    L = NV.lib()
    for name in NV.SYMBOLS:
        getattr(L, name)
    assert L.dcol_altro_version().startswith(b"dcol-altro-core")


def _random_state(p, rng, n):
    X = p.X0[0] + rng.normal(size=(n, p.nx)) * 0.3
    U = p.U0[0] + rng.normal(size=(n, p.nu)) * (2.0 if p.name != "quadrotor" else 0.5)
    return X, U


@pytest.mark.parametrize("name", NAMES)
def test_dynamics_and_rk4(name):
    p = PROBLEMS[name]()
    core = NV.NativeCore(p)
    X, U = _random_state(p, np.random.default_rng(1), 200)
    np.testing.assert_allclose(core.dynamics(X, U), p.dynamics(X, U), rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(core.rk4(X, U), S._rk4(p, X, U), rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("name", NAMES)
def test_pass_functions(name):
    p = PROBLEMS[name]()
    core = NV.NativeCore(p)
    rng = np.random.default_rng(2)
    N, nx, nu, no = p.N, p.nx, p.nu, p.n_obs
    X = p.X0 + np.cumsum(rng.normal(size=(N, nx)) * 0.02, axis=0)
    U = p.U0 + rng.normal(size=(N - 1, nu)) * (0.3 if name != "quadrotor" else 1e-3)   # rotor noise spins the MRP up
    hx = rng.normal(size=(N, no)) * 0.5
    ghx = rng.normal(size=(N, no, nx))
    mu = np.maximum(0.0, rng.normal(size=(N - 1, 2 * nu)))
    mux = np.maximum(0.0, rng.normal(size=(N, no)))
    lambd = rng.normal(size=nx)
    rho, reg = 10.0, 1e-4
    # forward-difference Jacobians: identical perturbations, so only rounding separates the two
    A, B = core.jacobians(X, U)
    A2, B2 = S._fd_jacobians(p, X, U)
    np.testing.assert_allclose(A, A2, rtol=0, atol=1e-8)
    np.testing.assert_allclose(B, B2, rtol=0, atol=1e-8)
    K, k, dJ = core.backward_pass(X, U, hx, ghx, mu, mux, lambd, rho, reg)
    K2, k2, dJ2 = S._backward_pass(p, X, U, hx, ghx, mu, mux, lambd, rho, reg)
    scale = max(1.0, np.abs(K2).max())
    assert np.abs(K - K2).max() / scale < 1e-6 and np.abs(k - k2).max() / max(1.0, np.abs(k2).max()) < 1e-6
    assert abs(dJ - dJ2) / max(1.0, abs(dJ2)) < 1e-6
    alphas = [0.5 ** i for i in range(p.max_linesearch_iters)]
    g = 1.0 if name != "quadrotor" else 1e-2
    Kr, kr = rng.normal(size=K2.shape) * 0.02 * g, rng.normal(size=k2.shape) * 0.05 * g   # mild gains: bounded rollouts
    Xn, Un = core.rollouts(X, U, Kr, kr, alphas)
    Xn2, Un2 = S._rollouts(p, X, U, Kr, kr, alphas)
    ok = np.isfinite(Xn2).all(axis=(1, 2)) & (np.abs(Xn2).max(axis=(1, 2)) < 1e3)
    assert ok.sum() >= 10
    np.testing.assert_allclose(Xn[ok], Xn2[ok], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(Un[ok], Un2[ok], rtol=1e-7, atol=1e-7)
    hxn = rng.normal(size=(len(alphas), N, no)) * 0.5
    c = core.total_cost(Xn2[ok], Un2[ok], hxn[ok], mu, mux, lambd, rho)
    c2 = S._total_cost(p, Xn2[ok], Un2[ok], hxn[ok], mu, mux, lambd, rho)
    np.testing.assert_allclose(c, c2, rtol=1e-12)
    assert abs(core.total_cost(X, U, hx, mu, mux, lambd, rho) - float(S._total_cost(p, X, U, hx, mu, mux, lambd, rho))) \
        < 1e-9 * abs(float(S._total_cost(p, X, U, hx, mu, mux, lambd, rho)))


def test_not_positive_definite_raises_like_scipy():
    p = PROBLEMS["piano_mover"]()
    core = NV.NativeCore(p)
    N, nx, nu, no = p.N, p.nx, p.nu, p.n_obs
    X, U = p.X0.copy(), p.U0.copy()
    X[5, 0] = np.nan
    with pytest.raises(np.linalg.LinAlgError):
        core.backward_pass(X, U, np.zeros((N, no)), np.zeros((N, no, nx)), np.zeros((N - 1, 2 * nu)), np.zeros((N, no)),
                           np.zeros(nx), 1.0, 1e-6)

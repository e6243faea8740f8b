"""AL-iLQR ("ALTRO") caller with BATCHED collision constraints (SURVEY.md section 8(f), rows N1 + N2).

This is the caller side of the drop-in boundary, restated so that it feeds the batched proximity engine:
the optimiser of the reference (``ALTRO.py:365-488``: backward Riccati pass with augmented-Lagrangian
terms ``:242-338``, forward rollout with a halving line search ``:183-239``, regularisation update ``:48-75``,
dual / penalty updates ``:444-483``) with the same formulas, tolerances and update rules, but

* every collision constraint of a pass — all knots x all obstacles, value AND gradient — comes from ONE
  batched solve (the reference makes ``2 N n_obs`` scalar calls per backward pass and re-solves each problem
  for its gradient, ``ALTRO.py:276-278,310-312``);
* the line search is speculative: all ``max_linesearch_iters`` step sizes 2^-k are rolled out together
  (vectorised RK4) and their ``K N n_obs`` constraints evaluated in one batched solve; the first step size
  that lowers the cost is taken, which is exactly what the sequential loop ``ALTRO.py:212-234`` returns;
* the cost of the current trajectory is computed once per pass, not once per trial (``ALTRO.py:215``), and the
  constraint values of the accepted trajectory are re-used for the dual update (``ALTRO.py:457-461``).

The Riccati recursion itself is sequential over knots on 12x12 matrices and stays on the host: for the three
built-in systems in the native core ``libdcol_altro.so`` (``csrc/altro_core.cpp``, ``include/dcol_altro.h``: RK4
rollouts, forward-difference Jacobians + Riccati sweep, cost), otherwise — user-defined dynamics — in the NumPy
functions below, which are also the cross-check of the native core (``tests/test_altro_native.py``).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np
from scipy.linalg import cho_factor, cho_solve

from .problems import Problem


class EngineEvaluator:
    """Collision constraints of many victim poses against the problem's obstacles through the CUDA engine.

    One plan per batch size (pairs = poses x obstacles, victim = shape 0, obstacle j = shape 1 + j); the
    obstacle poses never change, so a solve moves only the victim poses in and (alpha, grad) out."""

    def __init__(self, problem: Problem, device: int = 0):
        from ..engine import ProximityEngine, raise_for_status
        from ..shapes import pose_of
        self._raise = raise_for_status
        self.n_obs = problem.n_obs
        self.engine = ProximityEngine([problem.victim] + list(problem.obstacles), device=device)
        self.obs_shape = np.arange(1, self.n_obs + 1, dtype=np.int32)
        self.obs_pose = np.ascontiguousarray(np.stack([pose_of(o) for o in problem.obstacles]))
        self.pair_solves = 0
        self.calls = 0

    def evaluate(self, victim_poses: np.ndarray, want_grad: bool):
        """``victim_poses [M, 6]`` -> ``alpha [M, n_obs]``, ``grad1 [M, n_obs, 6]`` (or None), ``status [M, n_obs]``.
        ONE call of the scene entry point of the C ABI (``dcol_proximity_scene_host``): the victim poses go in, alpha,
        the victim half of the gradient and the status words come out; nothing is raised."""
        res = self.engine.solve_scene_host(0, victim_poses, self.obs_shape, self.obs_pose, want_grad=want_grad,
                                           want_iters=False)
        self.pair_solves += victim_poses.shape[0] * self.n_obs
        self.calls += 1
        return res.alpha, res.grad1, res.status

    def __call__(self, victim_poses: np.ndarray, want_grad: bool):
        """As :meth:`evaluate`, raising what the reference raises for the first failed pair."""
        alpha, grad, status = self.evaluate(victim_poses, want_grad)
        if status.any():
            self._raise(int(status[status != 0][0]))
        return alpha, grad

    def close(self):
        self.engine.close()


@dataclass
class AltroResult:
    X: np.ndarray
    U: np.ndarray
    passes: int                 # backward/forward passes performed
    converged: bool
    cost: float
    rho: float
    penalty_updates: int
    pair_solves: int            # proximity problems solved
    batched_calls: int
    wall_s: float
    X_hist: list = field(default_factory=list)
    log: list = field(default_factory=list)
    timing: dict = field(default_factory=dict)   # seconds: setup, constraints (engine calls), backward, rollouts, cost, teardown


def _rk4(problem: Problem, X, U):
    """discrete_dynamics of every system script (e.g. piano_mover.py:25-43), vectorised over leading axes."""
    dt, f = problem.dt, problem.dynamics
    lead = np.broadcast_shapes(X.shape[:-1], U.shape[:-1])
    X = np.broadcast_to(X, lead + X.shape[-1:])
    U = np.broadcast_to(U, lead + U.shape[-1:])
    k1 = dt * f(X, U)
    k2 = dt * f(X + 0.5 * k1, U)
    k3 = dt * f(X + 0.5 * k2, U)
    k4 = dt * f(X + k3, U)
    return X + (1.0 / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)


def _fd_jacobians(problem: Problem, X, U, delta=1e-6):
    """Forward-difference dynamics Jacobians of all knots at once (compute_jacobian, ALTRO.py:77-100)."""
    N1, nx, nu = U.shape[0], problem.nx, problem.nu
    Xs, Us = X[:N1], U
    base = _rk4(problem, Xs, Us)                                         # [N-1, nx]
    Xp = Xs[:, None, :] + delta * np.eye(nx)[None, :, :]                 # [N-1, nx, nx]  row i: x + delta e_i
    A = (_rk4(problem, Xp, Us[:, None, :]) - base[:, None, :]) / delta   # [N-1, i, out]
    Up = Us[:, None, :] + delta * np.eye(nu)[None, :, :]
    B = (_rk4(problem, Xs[:, None, :], Up) - base[:, None, :]) / delta
    return np.swapaxes(A, 1, 2), np.swapaxes(B, 1, 2)                    # [N-1, nx, nx], [N-1, nx, nu]


def _al_terms(h, mult, rho):
    """mult . h + rho/2 h' I_mask h with mask = (mult > 0 or h > 0)  (eval_mask, ALTRO.py:16-30; :127-133)."""
    mask = (mult > 0) | (h > 0)
    return np.sum(mult * h, axis=-1) + 0.5 * rho * np.sum(np.where(mask, h * h, 0.0), axis=-1)


def _total_cost(problem: Problem, X, U, hx, mu, mux, lambd, rho):
    """compute_total_cost (ALTRO.py:103-143) for a stack of trajectories X[..., N, nx], U[..., N-1, nu],
    hx[..., N, n_obs]; terms are accumulated knot by knot in the reference's order."""
    N = problem.N
    dX = X - problem.Xref
    dU = U - problem.Uref[:N - 1]
    run = 0.5 * np.einsum("...ti,ij,...tj->...t", dX[..., :N - 1, :], problem.Q, dX[..., :N - 1, :]) \
        + 0.5 * np.einsum("...ti,ij,...tj->...t", dU, problem.R, dU)
    hu = np.concatenate([U - problem.u_max, -U + problem.u_min], axis=-1)
    cu = _al_terms(hu, mu, rho)                                          # [..., N-1]
    cx = _al_terms(hx, mux, rho)                                         # [..., N]
    per_knot = np.stack([run, cu, cx[..., :N - 1]], axis=-1).reshape(X.shape[:-2] + (3 * (N - 1),))
    cost = np.cumsum(per_knot, axis=-1)[..., -1]
    cost = cost + 0.5 * np.einsum("...i,ij,...j->...", dX[..., -1, :], problem.Qf, dX[..., -1, :])
    cost = cost + cx[..., -1]
    goal = X[..., -1, :] - problem.Xref[-1]
    return cost + np.sum(lambd * goal, axis=-1) + 0.5 * rho * np.sum(goal * goal, axis=-1)


def _backward_pass(problem: Problem, X, U, hx, ghx, mu, mux, lambd, rho, reg):
    """backward_pass (ALTRO.py:242-338) given the batched constraint values hx[N, n_obs] and gradients
    ghx[N, n_obs, nx] of the whole trajectory."""
    N, nx, nu = problem.N, problem.nx, problem.nu
    A, B = _fd_jacobians(problem, X, U)
    Iu = np.vstack([np.eye(nu), -np.eye(nu)])
    K = np.zeros((N - 1, nu, nx))
    k = np.zeros((N - 1, nu))
    delta_J = 0.0
    # terminal knot
    mask = ((mux[-1] > 0) | (hx[-1] > 0)).astype(float)
    Vx = problem.Qf @ (X[-1] - problem.Xref[-1]) + ghx[-1].T @ (mux[-1] + rho * (mask * hx[-1]))
    Vxx = problem.Qf + rho * ghx[-1].T @ (mask[:, None] * ghx[-1])
    goal = X[-1] - problem.Xref[-1]
    Vx = Vx + (lambd + rho * goal)
    Vxx = Vxx + rho * np.eye(nx)
    for t in range(N - 2, -1, -1):
        lx = problem.Q @ (X[t] - problem.Xref[t])
        lu = problem.R @ (U[t] - problem.Uref[t])
        lxx, luu = problem.Q.copy(), problem.R.copy()
        hu = np.concatenate([U[t] - problem.u_max, -U[t] + problem.u_min])
        mask_u = ((mu[t] > 0) | (hu > 0)).astype(float)
        lu = lu + Iu.T @ (mu[t] + rho * (mask_u * hu))
        luu = luu + rho * Iu.T @ (mask_u[:, None] * Iu)
        mask = ((mux[t] > 0) | (hx[t] > 0)).astype(float)
        lx = lx + ghx[t].T @ (mux[t] + rho * (mask * hx[t]))
        lxx = lxx + rho * ghx[t].T @ (mask[:, None] * ghx[t])
        At, Bt = A[t], B[t]
        Vreg = Vxx + reg * np.eye(nx)
        Qx = lx + At.T @ Vx
        Qu = lu + Bt.T @ Vx
        Quu = luu + Bt.T @ Vreg @ Bt
        Qux = Bt.T @ Vreg @ At
        cf = cho_factor(Quu)
        kt = cho_solve(cf, Qu)
        Kt = cho_solve(cf, Qux)
        Acl = At - Bt @ Kt
        Vx_new = lx - Kt.T @ lu + Kt.T @ luu @ kt + Acl.T @ (Vx - Vxx @ Bt @ kt)
        Vxx = lxx + Kt.T @ luu @ Kt + Acl.T @ Vxx @ Acl
        Vx = Vx_new
        delta_J += float(Qu @ kt)
        K[t], k[t] = Kt, kt
    return K, k, delta_J


def _rollouts(problem: Problem, X, U, K, k, alphas):
    """Closed-loop rollouts for every step size at once (forward_pass, ALTRO.py:219-221)."""
    C, N = len(alphas), problem.N
    Xn = np.empty((C, N, problem.nx))
    Un = np.empty((C, N - 1, problem.nu))
    Xn[:, 0] = X[0]
    a = np.asarray(alphas)[:, None]
    for t in range(N - 1):
        Un[:, t] = U[t] - (Xn[:, t] - X[t]) @ K[t].T - a * k[t]
        Xn[:, t + 1] = _rk4(problem, Xn[:, t], Un[:, t])
    return Xn, Un


class NumpyCore:
    """The per-pass host work in NumPy, for user-defined dynamics (``Problem.dynamics`` is any vectorised callable);
    same interface as :class:`native.NativeCore`, which runs the reference's three systems in C++."""

    def __init__(self, problem: Problem):
        self.problem = problem

    def rk4(self, X, U):
        return _rk4(self.problem, np.atleast_2d(X), np.atleast_2d(U))

    def rollouts(self, X, U, K, k, alphas):
        return _rollouts(self.problem, X, U, K, k, alphas)

    def backward_pass(self, X, U, hx, ghx, mu, mux, lambd, rho, reg):
        return _backward_pass(self.problem, X, U, hx, ghx, mu, mux, lambd, rho, reg)

    def total_cost(self, X, U, hx, mu, mux, lambd, rho):
        return _total_cost(self.problem, X, U, hx, mu, mux, lambd, rho)


def altro_solve(problem: Problem, evaluator=None, speculative: bool = True, verbose: bool = False,
                keep_history: bool = False, native: bool | None = None) -> AltroResult:
    """Run AL-iLQR on ``problem``.  ``evaluator(victim_poses[M, 6], want_grad) -> (alpha[M, n_obs],
    grad1[M, n_obs, 6] | None)`` defaults to the CUDA engine (:class:`EngineEvaluator`).  ``native``: run the
    per-pass host work (rollouts, Jacobians + Riccati, cost) in the C++ core; default: whenever the problem is one
    of its built-in systems (``problem.extra['native']``)."""
    t_start = time.perf_counter()
    clock = time.perf_counter
    tm = {"setup": 0.0, "constraints": 0.0, "backward": 0.0, "rollouts": 0.0, "cost": 0.0, "teardown": 0.0}
    own = evaluator is None
    if own:
        evaluator = EngineEvaluator(problem)
    if native is None:
        native = "native" in problem.extra
    if native:
        from .native import NativeCore
        core = NativeCore(problem)
    else:
        core = NumpyCore(problem)
    N, nx, nu, n_obs = problem.N, problem.nx, problem.nu, problem.n_obs
    X, U = problem.X0.copy(), problem.U0.copy()
    for t in range(N - 1):                                              # initial rollout, ALTRO.py:399-400
        X[t + 1] = core.rk4(X[t], U[t])[0]
    mu = np.zeros((N - 1, 2 * nu))
    mux = np.zeros((N, n_obs))
    lambd = np.zeros(nx)
    rho, reg = problem.rho, problem.reg_min
    ls_alphas = [0.5 ** i for i in range(problem.max_linesearch_iters)]
    hist, log = ([X.copy()] if keep_history else []), []
    penalty_updates, converged, passes, J = 0, False, 0, np.nan
    pair_solves = calls = 0

    tm["setup"] = clock() - t_start

    def constraints(Xs, want_grad, tolerant=False):
        nonlocal pair_solves, calls
        t0 = clock()
        try:
            return _constraints(Xs, want_grad, tolerant)
        finally:
            tm["constraints"] += clock() - t0

    def _constraints(Xs, want_grad, tolerant):
        """tolerant: failed solves do not raise; the third return value holds, per leading index of ``Xs`` (per
        trajectory), the status word of its first failed pair (0 = none).  Only evaluators with ``evaluate`` support it."""
        nonlocal pair_solves, calls
        lead = Xs.shape[:-1]
        poses = problem.pose_of_state(Xs).reshape(-1, 6)
        failed = None
        if tolerant and hasattr(evaluator, "evaluate"):
            alpha, g, status = evaluator.evaluate(poses, want_grad)
            st = status.reshape(lead[:-1] + (-1,)) if len(lead) > 1 else status.reshape(1, -1)
            first = np.argmax(st != 0, axis=-1)
            failed = np.take_along_axis(st, first[..., None], axis=-1)[..., 0]
        else:
            alpha, g = evaluator(poses, want_grad)
        pair_solves += poses.shape[0] * n_obs
        calls += 1
        hx = (1.0 - alpha).reshape(lead + (n_obs,))                     # inequality_constraints_x, e.g. piano_mover.py:66
        if not want_grad:
            return hx, None, failed
        Jp = problem.pose_jacobian(Xs).reshape(-1, 6, nx)               # d alpha / d x = g[0:6] . d(r, p)/dx
        ghx = -np.einsum("moi,mix->mox", g, Jp).reshape(lead + (n_obs, nx))
        return hx, ghx, failed

    for itr in range(problem.max_iters):
        passes = itr + 1
        hx, ghx, _ = constraints(X, True)                                # ONE batched solve: values + gradients
        t0 = clock()
        K, k, delta_J = core.backward_pass(X, U, hx, ghx, mu, mux, lambd, rho, reg)
        t1 = clock()
        old_cost = float(core.total_cost(X, U, hx, mu, mux, lambd, rho))
        tm["backward"] += t1 - t0
        tm["cost"] += clock() - t1
        alpha, accepted = 0.0, None
        if speculative:
            t0 = clock()
            Xn, Un = core.rollouts(X, U, K, k, ls_alphas)
            tm["rollouts"] += clock() - t0
            hxn, _, failed = constraints(Xn, False, tolerant=True)       # ONE batched solve for all step sizes
            t0 = clock()
            costs = core.total_cost(Xn, Un, np.nan_to_num(hxn, nan=0.0), mu, mux, lambd, rho)
            tm["cost"] += clock() - t0
            if failed is not None and failed.any():
                # The sequential loop (ALTRO.py:212-234) evaluates step sizes one by one and stops at the first that
                # lowers the cost: a PDIP failure at a LATER (smaller) step size is never seen by it and must not abort
                # the solve; one at or before the accepted step size raises there too.
                costs = np.where(failed != 0, np.inf, costs)
                ok_better = np.nonzero(costs < old_cost)[0]
                stop = int(ok_better[0]) if ok_better.size else len(ls_alphas) - 1
                seen = np.nonzero(failed[:stop + 1])[0]
                if seen.size:
                    from ..engine import raise_for_status
                    raise_for_status(int(failed[seen[0]]))
            better = np.nonzero(costs < old_cost)[0]
            if better.size:
                c = int(better[0])
                alpha, accepted = ls_alphas[c], (Xn[c].copy(), Un[c].copy(), hxn[c].copy(), float(costs[c]))
        else:
            for a in ls_alphas:
                Xn, Un = core.rollouts(X, U, K, k, [a])
                hxn, _, _ = constraints(Xn, False)
                cost = float(core.total_cost(Xn, Un, hxn, mu, mux, lambd, rho)[0])
                if cost < old_cost:
                    alpha, accepted = a, (Xn[0].copy(), Un[0].copy(), hxn[0].copy(), cost)
                    break
        if accepted is not None:
            X, U, hx, J = accepted
        else:
            J = old_cost                                                 # ALTRO.py:236-239
        if keep_history:
            hist.append(X.copy())
        # regularisation, update_reg ALTRO.py:48-75
        if alpha == 0.0:
            if reg == problem.reg_max:
                raise ValueError("Regularization parameter reached maximum value.")
            reg = min(problem.reg_max, reg * 10)
        elif alpha == 1.0:
            reg = max(problem.reg_min, reg / 10)
        kmax = float(np.max(np.linalg.norm(k, axis=1)))
        log.append((itr + 1, J, delta_J, kmax, alpha, reg, rho))
        if verbose:
            print(f"{itr + 1:3d}   {J:10.3e}  {delta_J:9.2e}  {kmax:9.2e}  {alpha:6.4f}   {reg:9.2e}   {rho:9.2e}", flush=True)
        if alpha > 0 and kmax < problem.atol:                            # ALTRO.py:444-483
            hu = np.concatenate([U - problem.u_max, -U + problem.u_min], axis=-1)
            mask_u = (mu > 0) | (hu > 0)
            mu = np.maximum(0.0, mu + rho * np.where(mask_u, hu, 0.0))
            convio = float(np.max(np.abs(hu + np.abs(hu)))) if hu.size else 0.0
            mask_x = (mux > 0) | (hx > 0)
            mux = np.maximum(0.0, mux + rho * np.where(mask_x, hx, 0.0))
            convio = max(convio, float(np.max(np.abs(hx + np.abs(hx)))))
            goal = X[-1] - problem.Xref[-1]
            lambd = lambd + rho * goal
            convio = max(convio, float(np.max(np.abs(goal))))
            if convio < problem.convio_tol:
                converged = True
                break
            rho *= problem.phi
            penalty_updates += 1
    t0 = clock()
    if own:
        evaluator.close()
    tm["teardown"] = clock() - t0
    return AltroResult(X=X, U=U, passes=passes, converged=converged, cost=float(J), rho=rho,
                       penalty_updates=penalty_updates, pair_solves=pair_solves, batched_calls=calls,
                       wall_s=time.perf_counter() - t_start, X_hist=hist, log=log, timing=tm)

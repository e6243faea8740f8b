"""Batched constraint hooks for the REFERENCE's own ``params`` dictionaries.

The reference's system scripts expose ``inequality_constraints_x(params, x)`` / ``..._grad(params, x)`` for ONE state
and loop over ``params['P_obs']`` inside (``systems/piano_mover.py:54-95``,
``systems/cluttered_hallway_quadrotor.py:120-167``, ``systems/cone_through_wall.py:122-167``).  A maintainer who wants
the batched path inside the reference's ``ALTRO.py`` replaces the per-knot loops (``ALTRO.py:120-139,265-314``) by one
call of :func:`make_batched_hooks`' functions over the whole trajectory:

    hooks = make_batched_hooks(params)                  # once, after initialize_<system>()
    HX  = hooks.constraints_x(X)                        # [N, n_obs]      == stack of inequality_constraints_x(params, X[t])
    HX, GX = hooks.constraints_x_with_grad(X)           # [N, n_obs, nx]  == stack of inequality_constraints_x_grad(params, X[t])

The state -> victim-pose maps are the ones of the three system scripts, selected by ``params['system']``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import numpy as np

from . import problems as _p


def _piano_pose(X):
    out = np.zeros(X.shape[:-1] + (6,))
    out[..., 0:2] = X[..., 0:2]
    out[..., 5] = np.tan(X[..., 4] / 4.0)
    return out


def _piano_jac(X):
    J = np.zeros(X.shape[:-1] + (6, X.shape[-1]))
    J[..., 0, 0] = 1.0
    J[..., 1, 1] = 1.0
    J[..., 5, 4] = 1.0 / (4.0 * np.cos(X[..., 4] / 4.0) ** 2)
    return J


_MAPS = {
    "piano_mover": (_piano_pose, _piano_jac),
    "quadrotor": (lambda X: np.concatenate([X[..., 0:3], X[..., 6:9]], axis=-1), _p._pose_jac_6dof),
    "coneThroughWall": (lambda X: np.concatenate([X[..., 0:3], X[..., 6:9]], axis=-1), _p._pose_jac_6dof),
}


@dataclass
class BatchedHooks:
    constraints_x: Callable
    constraints_x_with_grad: Callable
    evaluator: object


def make_batched_hooks(params: dict, evaluator=None) -> BatchedHooks:
    """``params``: a dictionary made by the reference's ``initialize_<system>()`` (uses ``system``, ``P_vic``,
    ``P_obs``, ``nx``).  ``evaluator`` defaults to the CUDA engine."""
    pose_of_state, pose_jacobian = _MAPS[params["system"]]
    n_obs, nx = len(params["P_obs"]), params["nx"]
    if evaluator is None:
        from .solver import EngineEvaluator

        class _Shim:      # what EngineEvaluator needs of a Problem
            victim, obstacles, n_obs = params["P_vic"], list(params["P_obs"]), len(params["P_obs"])
        evaluator = EngineEvaluator(_Shim)

    def constraints_x(X):
        X = np.asarray(X, dtype=float)
        alpha, _ = evaluator(pose_of_state(X).reshape(-1, 6), False)
        return (1.0 - alpha).reshape(X.shape[:-1] + (n_obs,))

    def constraints_x_with_grad(X):
        X = np.asarray(X, dtype=float)
        alpha, g = evaluator(pose_of_state(X).reshape(-1, 6), True)
        Jp = pose_jacobian(X).reshape(-1, 6, nx)
        ghx = -np.einsum("moi,mix->mox", g, Jp)
        return (1.0 - alpha).reshape(X.shape[:-1] + (n_obs,)), ghx.reshape(X.shape[:-1] + (n_obs, nx))

    return BatchedHooks(constraints_x, constraints_x_with_grad, evaluator)

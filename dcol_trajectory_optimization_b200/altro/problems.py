"""The three scenarios of the reference as vectorised problem definitions.

Each problem mirrors one ``initialize_*`` of the reference's ``systems/`` scripts — same dimensions,
weights, bounds, tolerances, obstacles and initial guess (data: ``data/scenes.npz``, extracted by
``oracle/gen_scene_data.py``) — but exposes its dynamics and its state -> victim-pose map in batched
form, so that the caller can evaluate every (candidate, knot, obstacle) collision constraint in ONE
call of the batched proximity engine instead of ``n_obs`` scalar calls per knot:

  piano_mover        systems/piano_mover.py:5-232            box vs 3 boxes, planar, N = 80
  cone_through_wall  systems/cone_through_wall.py:12-330     cone vs 4 boxes, 6-DOF rigid body, N = 60
  quadrotor          systems/cluttered_hallway_quadrotor.py:17-387   sphere vs 11 mixed primitives, N = 100
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable

import numpy as np

from ..primitives import (CapsuleMRP, ConeMRP, CylinderMRP, PolygonMRP, PolytopeMRP, SphereMRP, create_n_sided,
                          create_rect_prism)
from ..workloads import hallway_polytopes

_SCENES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "scenes.npz")


@dataclass
class Problem:
    name: str
    nx: int
    nu: int
    N: int
    dt: float
    Q: np.ndarray
    R: np.ndarray
    Qf: np.ndarray
    Xref: np.ndarray            # [N, nx]
    Uref: np.ndarray            # [>= N-1, nu]
    u_min: np.ndarray
    u_max: np.ndarray
    X0: np.ndarray              # [N, nx] initial state guess (every knot at x0)
    U0: np.ndarray              # [N-1, nu]
    victim: object
    obstacles: list
    dynamics: Callable          # (X[..., nx], U[..., nu]) -> Xdot[..., nx]
    pose_of_state: Callable     # X[..., nx] -> [..., 6]  victim (r, p)
    pose_jacobian: Callable     # X[..., nx] -> [..., 6, nx]  d (r, p) / d x
    atol: float = 1e-2
    convio_tol: float = 1e-4
    rho: float = 1.0
    phi: float = 10.0
    reg_min: float = 1e-6
    reg_max: float = 1e2
    max_iters: int = 3000
    max_linesearch_iters: int = 20
    extra: dict = field(default_factory=dict)   # extra["native"]: built-in system of the C++ host core (altro/native.py)

    @property
    def n_obs(self) -> int:
        return len(self.obstacles)


def _scenes():
    with np.load(_SCENES) as d:
        return {k: d[k] for k in d.files}


def _cross(a, b):
    return np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                     a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], axis=-1)


def _mrp_rate(p, omega):
    """p_dot = ((1 + |p|^2) / 4) (I + 2 ([p x]^2 + [p x]) / (1 + |p|^2)) omega
    (cluttered_hallway_quadrotor.py:60-61, cone_through_wall.py:41-44), written with cross products."""
    pp = np.sum(p * p, axis=-1, keepdims=True)
    pw = _cross(p, omega)
    return 0.25 * (1.0 + pp) * omega + 0.5 * (_cross(p, pw) + pw)


def _rotate_by_mrp(p, v):
    """Q(p) v with Q = I + (8 [p x]^2 + 4 (1 - |p|^2) [p x]) / (1 + |p|^2)^2 (problem_matrices.py:213-251)."""
    pp = np.sum(p * p, axis=-1, keepdims=True)
    pv = _cross(p, v)
    return v + (8.0 * _cross(p, pv) + 4.0 * (1.0 - pp) * pv) / (1.0 + pp) ** 2


def _place(prims, poses):
    for prim, pose in zip(prims, poses):
        prim.r = np.array(pose[:3])
        prim.p = np.array(pose[3:])
    return prims


def _pose_jac_6dof(X):
    """r = x[0:3], p = x[6:9] (cluttered_hallway_quadrotor.py:127-128, :161-163)."""
    J = np.zeros(X.shape[:-1] + (6, X.shape[-1]))
    for i in range(3):
        J[..., i, i] = 1.0
        J[..., 3 + i, 6 + i] = 1.0
    return J


def piano_mover() -> Problem:
    S = _scenes()
    nx, nu, N = 6, 3, 80

    def dynamics(X, U):                                      # piano_mover.py:5-23
        return np.concatenate([X[..., 2:4], U[..., 0:2], X[..., 5:6], U[..., 2:3] / 100.0], axis=-1)

    def pose_of_state(X):                                    # piano_mover.py:64-65
        out = np.zeros(X.shape[:-1] + (6,))
        out[..., 0:2] = X[..., 0:2]
        out[..., 5] = np.tan(X[..., 4] / 4.0)
        return out

    def pose_jacobian(X):                                    # piano_mover.py:83-94
        J = np.zeros(X.shape[:-1] + (6, nx))
        J[..., 0, 0] = 1.0
        J[..., 1, 1] = 1.0
        J[..., 5, 4] = 1.0 / (4.0 * np.cos(X[..., 4] / 4.0) ** 2)
        return J

    obstacles = _place([create_rect_prism(3.0, 3.0, 1.0), create_rect_prism(4.0, 1.0, 1.0),
                        create_rect_prism(1.0, 5.0, 1.1)], S["piano_obs_pose"])
    return Problem(name="piano_mover", nx=nx, nu=nu, N=N, dt=0.1, Q=np.eye(nx), R=np.diag([1.0, 1.0, 0.001]),
                   Qf=np.eye(nx), Xref=S["piano_Xref"], Uref=S["piano_Uref"], u_min=-200.0 * np.ones(nu),
                   u_max=200.0 * np.ones(nu), X0=S["piano_X0"], U0=S["piano_U0"],
                   victim=create_rect_prism(2.5, 0.15, 0.01), obstacles=obstacles, dynamics=dynamics,
                   pose_of_state=pose_of_state, pose_jacobian=pose_jacobian, atol=4e-2,
                   extra={"native": {"system": "piano"}})


def cone_through_wall() -> Problem:
    S = _scenes()
    nx, nu, N = 12, 6, 60
    mass, Jd = float(S["cone_mass"]), np.diag(S["cone_inertia"]).copy()

    def dynamics(X, U):                                      # cone_through_wall.py:19-47
        v, p, w = X[..., 3:6], X[..., 6:9], X[..., 9:12]
        wdot = (U[..., 3:6] - _cross(w, Jd * w)) / Jd
        return np.concatenate([v, U[..., 0:3] / mass, _mrp_rate(p, w), wdot], axis=-1)

    obstacles = _place([create_rect_prism(10.0, 10.0, 1.0), create_rect_prism(10.0, 10.0, 1.0),
                        create_rect_prism(4.1, 4.1, 1.1), create_rect_prism(4.1, 4.1, 1.1)], S["cone_obs_pose"])
    return Problem(name="coneThroughWall", nx=nx, nu=nu, N=N, dt=0.1, Q=np.eye(nx),
                   R=np.diag([1.0, 1.0, 1.0, 100.0, 100.0, 100.0]), Qf=np.eye(nx), Xref=S["cone_Xref"],
                   Uref=S["cone_Uref"], u_min=-20.0 * np.ones(nu), u_max=20.0 * np.ones(nu), X0=S["cone_X0"],
                   U0=S["cone_U0"], victim=ConeMRP(2.0, np.deg2rad(22)), obstacles=obstacles, dynamics=dynamics,
                   pose_of_state=lambda X: np.concatenate([X[..., 0:3], X[..., 6:9]], axis=-1),
                   pose_jacobian=_pose_jac_6dof, atol=1e-1,
                   extra={"mass": mass, "inertia": Jd, "native": {"system": "rigid_body", "mass": mass, "inertia": Jd}})


def quadrotor() -> Problem:
    S = _scenes()
    nx, nu, N = 12, 4, 100
    mass, Jd = 0.5, np.array([0.0023, 0.0023, 0.004])
    L, kf, km = 0.1750, 1.0, 0.0245
    gravity = np.array([0.0, 0.0, -9.81])

    def dynamics(X, U):                                      # cluttered_hallway_quadrotor.py:17-74
        v, p, w = X[..., 3:6], X[..., 6:9], X[..., 9:12]
        F = np.maximum(0.0, kf * U)
        M = km * U
        thrust = np.zeros(X.shape[:-1] + (3,))
        thrust[..., 2] = F[..., 0] + F[..., 1] + F[..., 2] + F[..., 3]
        tau = np.stack([L * (F[..., 1] - F[..., 3]), L * (F[..., 2] - F[..., 0]),
                        M[..., 0] - M[..., 1] + M[..., 2] - M[..., 3]], axis=-1)
        f_world = mass * gravity + _rotate_by_mrp(p, thrust)
        wdot = (tau - _cross(w, Jd * w)) / Jd
        return np.concatenate([v, f_world / mass, _mrp_rate(p, w), wdot], axis=-1)

    _, _, A2, b2 = hallway_polytopes()
    ngon = create_n_sided(5, 0.6)
    obstacles = _place([CylinderMRP(0.6, 3.0), CapsuleMRP(0.2, 5.0), SphereMRP(0.8), ConeMRP(2.0, np.deg2rad(22)),
                        PolytopeMRP(A2, b2), PolygonMRP(ngon["A"], ngon["b"], 0.2), CylinderMRP(1.1, 2.3),
                        CapsuleMRP(0.8, 1.0), SphereMRP(0.5), create_rect_prism(20, 5, 0.2),
                        create_rect_prism(20, 5, 0.2)], S["quad_obs_pose"])
    return Problem(name="quadrotor", nx=nx, nu=nu, N=N, dt=0.08, Q=np.eye(nx), R=np.eye(nu), Qf=np.eye(nx),
                   Xref=S["quad_Xref"], Uref=S["quad_Uref"], u_min=-2000.0 * np.ones(nu), u_max=2000.0 * np.ones(nu),
                   X0=S["quad_X0"], U0=S["quad_U0"], victim=SphereMRP(0.25), obstacles=obstacles, dynamics=dynamics,
                   pose_of_state=lambda X: np.concatenate([X[..., 0:3], X[..., 6:9]], axis=-1),
                   pose_jacobian=_pose_jac_6dof, atol=1e-2,
                   extra={"native": {"system": "quadrotor", "mass": mass, "inertia": Jd, "arm": L, "kf": kf, "km": km}})


PROBLEMS = {"piano_mover": piano_mover, "coneThroughWall": cone_through_wall, "quadrotor": quadrotor}

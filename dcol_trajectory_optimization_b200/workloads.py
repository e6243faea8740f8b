"""Synthetic pair batches for the benchmark and the parity tests.

``config4``  — "synthetic batched proximity sweep": every reference-supported ordered type
               pair over a 7-shape table, random poses.
``config5``  — "scaled quadrotor hallway": a sphere victim against the quadrotor scene's 11
               obstacle shapes cycled to ``n_obs`` obstacles x knots x rollout candidates.

Shapes follow the reference's scenes (``systems/cluttered_hallway_quadrotor.py:281-307``);
poses are synthetic.  Pure numpy, host side only.
"""
from __future__ import annotations

import os

import numpy as np

from .primitives import (CapsuleMRP, ConeMRP, CylinderMRP, PolygonMRP, PolytopeMRP, SphereMRP,
                         create_n_sided, create_rect_prism)
from .shapes import kind_of, pair_supported

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "polytopes.npz")


def hallway_polytopes():
    """``(A1[14,3], b1, A2[8,3], b2)``: the two polytopes of the reference's
    ``systems/polytopes.jld2`` (stored there as 3xf; only ``A2, b2`` are used by the scene)."""
    with np.load(_DATA) as d:
        return d["A1"].T.copy(), d["b1"].copy(), d["A2"].T.copy(), d["b2"].copy()


def config4_shapes():
    """The 7-shape table of config 4, in the order SURVEY.md section 8(d) lists it."""
    _, _, A2, b2 = hallway_polytopes()
    ngon = create_n_sided(5, 0.6)
    return [
        create_rect_prism(1.0, 2.0, 3.0),
        PolytopeMRP(A2, b2),
        CapsuleMRP(0.3, 1.2),
        CylinderMRP(0.4, 1.5),
        ConeMRP(2.0, np.deg2rad(22)),
        SphereMRP(0.5),
        PolygonMRP(ngon["A"], ngon["b"], 0.2),
    ]


def supported_type_pairs(shapes):
    """Row-major ordered shape-index pairs minus those the reference cannot assemble."""
    kinds = [kind_of(s) for s in shapes]
    return [(i, j) for i in range(len(shapes)) for j in range(len(shapes))
            if pair_supported(kinds[i], kinds[j])]


def config4_poses_exact(n_pairs: int, seed: int = 1234):
    """Poses drawn pair by pair with exactly the call sequence of SURVEY.md section 8(d)
    (slow; used for the committed parity sample).  Returns ``pose1[n,6], pose2[n,6]``."""
    rng = np.random.default_rng(seed)
    pose1 = np.empty((n_pairs, 6))
    pose2 = np.empty((n_pairs, 6))
    for k in range(n_pairs):
        u = rng.normal(size=3)
        pose1[k, :3] = u / np.linalg.norm(u) * rng.uniform(0, 1)
        pose1[k, 3:] = rng.normal(size=3) * 0.5
        u = rng.normal(size=3)
        pose2[k, :3] = u / np.linalg.norm(u) * rng.uniform(0, 6)
        pose2[k, 3:] = rng.normal(size=3) * 0.5
    return pose1, pose2


def config4_poses(n_pairs: int, seed: int = 1234):
    """Same distribution as :func:`config4_poses_exact`, drawn vectorised (a different
    stream of the same generator) so that 2^20..2^26 pairs take seconds, not minutes."""
    rng = np.random.default_rng(seed)

    def ball(radius_hi):
        u = rng.normal(size=(n_pairs, 3))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        return u * rng.uniform(0, radius_hi, size=(n_pairs, 1))

    pose1 = np.concatenate([ball(1.0), rng.normal(size=(n_pairs, 3)) * 0.5], axis=1)
    pose2 = np.concatenate([ball(6.0), rng.normal(size=(n_pairs, 3)) * 0.5], axis=1)
    return pose1, pose2


def config4_batch(n_pairs: int, seed: int = 1234, exact: bool = False):
    """``(shapes, idx1[n], idx2[n], pose1[n,6], pose2[n,6])``; pair k uses the
    ``k mod 40``-th supported ordered type pair."""
    shapes = config4_shapes()
    pairs = np.asarray(supported_type_pairs(shapes), dtype=np.int32)
    sel = np.arange(n_pairs) % len(pairs)
    pose1, pose2 = (config4_poses_exact if exact else config4_poses)(n_pairs, seed)
    return shapes, pairs[sel, 0].copy(), pairs[sel, 1].copy(), pose1, pose2


def quadrotor_obstacle_shapes():
    """The 11 obstacle shapes of the cluttered hallway
    (``systems/cluttered_hallway_quadrotor.py:281-307``), poses left at the origin."""
    _, _, A2, b2 = hallway_polytopes()
    ngon = create_n_sided(5, 0.6)
    return [
        CylinderMRP(0.6, 3.0), CapsuleMRP(0.2, 5.0), SphereMRP(0.8), ConeMRP(2.0, np.deg2rad(22)),
        PolytopeMRP(A2, b2), PolygonMRP(ngon["A"], ngon["b"], 0.2), CylinderMRP(1.1, 2.3),
        CapsuleMRP(0.8, 1.0), SphereMRP(0.5), create_rect_prism(20, 5, 0.2), create_rect_prism(20, 5, 0.2),
    ]


def config5_batch(n_obs: int = 1024, n_knots: int = 100, n_cand: int = 256, seed: int = 2,
                  cand_slice: slice | None = None):
    """Scaled quadrotor hallway.  Shape 0 is the victim sphere (R = 0.25); shapes 1..11 are the
    obstacle shapes, obstacle ``j`` using shape ``1 + j mod 11``.

    Obstacle poses: ``r ~ U([-8,8]x[-2.5,2.5]x[1,6])``, ``p ~ N(0, 0.5^2)``; knots interpolate
    x0=(-8,0,4) -> xg=(8,0,4); candidate ``c`` perturbs every knot by ``N(0, 0.3^2)``, ``p1 = 0``.
    Pair order is ``[candidate][knot][obstacle]``; ``cand_slice`` restricts to a range of
    candidates (multi-GPU shards own whole trajectories).
    Returns ``(shapes, idx1, idx2, pose1[n,6], pose2[n,6])``.
    """
    rng = np.random.default_rng(seed)
    shapes = [SphereMRP(0.25)] + quadrotor_obstacle_shapes()
    lo = np.array([-8.0, -2.5, 1.0])
    hi = np.array([8.0, 2.5, 6.0])
    obs_r = rng.uniform(lo, hi, size=(n_obs, 3))
    obs_p = rng.normal(size=(n_obs, 3)) * 0.5
    obs_shape = 1 + (np.arange(n_obs) % 11)
    knots = np.linspace([-8.0, 0.0, 4.0], [8.0, 0.0, 4.0], n_knots)
    noise = rng.normal(size=(n_cand, n_knots, 3)) * 0.3
    cands = np.arange(n_cand)[cand_slice] if cand_slice is not None else np.arange(n_cand)
    vic_r = (knots[None, :, :] + noise[cands])                       # [c, t, 3]
    n = len(cands) * n_knots * n_obs
    pose1 = np.zeros((len(cands), n_knots, n_obs, 6))
    pose1[..., :3] = vic_r[:, :, None, :]
    pose2 = np.empty((len(cands), n_knots, n_obs, 6))
    pose2[..., :3] = obs_r
    pose2[..., 3:] = obs_p
    idx1 = np.zeros(n, dtype=np.int32)
    idx2 = np.broadcast_to(obs_shape.astype(np.int32), (len(cands), n_knots, n_obs)).reshape(n).copy()
    return shapes, idx1, idx2, pose1.reshape(n, 6), pose2.reshape(n, 6)

"""TEST INFRASTRUCTURE, generation time only: goldens for the two EXTENSIONS,
produced by the reference's OWN solver rather than by anything written for this repository.

What runs is the unmodified ``proximity_gradient`` of /root/reference — its ``problem_matrices`` per primitive, its
``solve_lp_pdip`` (pdip.py:373-470) with its NT scaling, its finite-difference gradient — with exactly two patches to
the ASSEMBLY, which is the only thing the extensions touch:

  case 4   ``combine_problem_matrices`` (combine_problem_matrices.py:58-67) builds the second primitive's blocks with the
           column layout [x, alpha, extras1, extras2] but forgets to pad the first primitive's blocks with ``v2 - 4``
           zero columns, so ``np.vstack`` raises.  The patch pads ``G_ort1`` / ``G_soc1`` and stacks — the layout the
           reference's own code intends.
  ellipsoid  absent from the code; Report.pdf section 3.1.5 eq. 27: ``|| U Q'^T (x - r') || <= alpha`` with
           ``U = diag(1/a, 1/b, 1/c)``, one SOC(4) block written like the sphere's (problem_matrices.py:151-178):
           ``G_soc = [[0 0 0 -1], [-U Q'^T, 0]]``, ``h_soc = [0; -U Q'^T r']``, no orthant rows.

    python oracle/gen_golden_extensions.py        # writes tests/golden/extensions.npz (about a minute)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as G          # noqa: E402  (helpers: reference import, per-pair recording, fixture layout)
from gen_golden import S, W, P  # noqa: E402


class EllipsoidRef:
    """An ellipsoid in the style of the reference's primitive classes (attribute bag)."""

    def __init__(self, a, b, c):
        self.semi_axes = (a, b, c)
        self.r = np.zeros(3)
        self.p = np.zeros(3)
        self.r_offset = np.zeros(3)
        self.Q_offset = np.eye(3)


def patch_reference():
    r = G.ref()
    import primitives.problem_matrices as pm
    orig_pm, dcm = pm.problem_matrices, pm.dcm_from_mrp

    def problem_matrices(shape, rr, p):
        if isinstance(shape, EllipsoidRef):
            Q = dcm(p)
            rp = rr + Q @ shape.r_offset
            Qp = Q @ shape.Q_offset
            U = np.diag(1.0 / np.asarray(shape.semi_axes, dtype=float))
            G_soc = np.zeros((4, 4))
            G_soc[0, 3] = -1.0
            G_soc[1:, :3] = -(U @ Qp.T)
            h_soc = np.concatenate([[0.0], -(U @ Qp.T @ rp)])
            return np.empty((0, 4)), np.empty((0,)), G_soc, h_soc
        return orig_pm(shape, rr, p)

    orig_combine = r.combine.combine_problem_matrices

    def combine(G_ort1, h_ort1, G_soc1, h_soc1, G_ort2, h_ort2, G_soc2, h_soc2):
        v1, v2 = G_ort1.shape[1], G_ort2.shape[1]
        if v1 > 4 and v2 > 4:       # case 4 with the first primitive padded, everything else as lines 58-67
            n_ort1, n_ort2, n_soc1, n_soc2 = G_ort1.shape[0], G_ort2.shape[0], G_soc1.shape[0], G_soc2.shape[0]
            e1, e2 = v1 - 4, v2 - 4
            G_ort_top = np.hstack([G_ort1, np.zeros((n_ort1, e2))])
            G_soc_top = np.hstack([G_soc1, np.zeros((n_soc1, e2))])
            G_ort_bot = np.hstack([G_ort2[:, :4], np.zeros((n_ort2, e1)), G_ort2[:, 4:]])
            G_soc_bot = np.hstack([G_soc2[:, :4], np.zeros((n_soc2, e1)), G_soc2[:, 4:]])
            Gm = np.vstack([G_ort_top, G_ort_bot, G_soc_top, G_soc_bot])
            h = np.hstack([h_ort1, h_ort2, h_soc1, h_soc2])
            n_ort = n_ort1 + n_ort2
            c = np.zeros(v1 + v2 - 4)
            c[3] = 1.0
            return (c, Gm, h, np.arange(0, n_ort), np.arange(n_ort, n_ort + n_soc1),
                    np.arange(n_ort + n_soc1, n_ort + n_soc1 + n_soc2))
        return orig_combine(G_ort1, h_ort1, G_soc1, h_soc1, G_ort2, h_ort2, G_soc2, h_soc2)

    for mod in (r.proximity, r.proximity_gradient):     # both bind the two names at import time
        mod.problem_matrices = problem_matrices
        mod.combine_problem_matrices = combine


_to_ref_prim = G.to_ref_prim


def to_ref_prim(prim, pose=None):
    if S.kind_of(prim) == S.ELLIPSOID:
        out = EllipsoidRef(*prim.semi_axes)
        out.r_offset = np.array(prim.r_offset, dtype=float)
        out.Q_offset = np.array(prim.Q_offset, dtype=float)
        pose = S.pose_of(prim) if pose is None else pose
        out.r, out.p = np.array(pose[:3], dtype=float), np.array(pose[3:], dtype=float)
        return out
    return _to_ref_prim(prim, pose)


def main():
    patch_reference()
    G.to_ref_prim = to_ref_prim                       # the pool workers fork after this point and inherit both patches
    shapes = W.config4_shapes() + [P.EllipsoidMRP(0.7, 0.4, 0.25), P.EllipsoidMRP(0.5, 0.5, 0.5)]
    off = P.EllipsoidMRP(0.9, 0.3, 0.6)               # one ellipsoid with body-frame offsets
    off.r_offset = np.array([0.2, -0.1, 0.3])
    c, s = np.cos(0.4), np.sin(0.4)
    off.Q_offset = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    shapes.append(off)
    kinds = [S.kind_of(p) for p in shapes]
    extras = {S.CAPSULE, S.CYLINDER, S.POLYGON}
    pairs = [(i, j) for i in range(len(shapes)) for j in range(len(shapes))
             if (kinds[i] in extras and kinds[j] in extras) or S.ELLIPSOID in (kinds[i], kinds[j])]
    per = 40
    n = per * len(pairs)
    sel = np.arange(n) % len(pairs)
    pr = np.asarray(pairs, dtype=np.int32)
    idx1, idx2 = pr[sel, 0].copy(), pr[sel, 1].copy()
    pose1, pose2 = W.config4_poses(n, seed=4242)
    res = G.run_batch(shapes, idx1, idx2, pose1, pose2, tol=1e-6)
    G.save("extensions", shapes, idx1, idx2, pose1, pose2, 1e-6, res, keep_sz=0,
           extra={"n_type_pairs": np.int32(len(pairs)), "case4": np.array([kinds[i] in extras and kinds[j] in extras
                                                                         for i, j in zip(idx1, idx2)])})


if __name__ == "__main__":
    main()

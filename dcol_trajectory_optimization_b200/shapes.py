"""Host-side flattening of primitive objects into the C-ABI shape table.

The reference dispatches on ``isinstance`` inside ``problem_matrices``
(``primitives/problem_matrices.py:255-364``) every time a pair is evaluated.  The
batched path instead flattens each distinct primitive *shape* once into a
``dcol_shape`` record (``include/dcol.h``) plus packed half-space arrays, and a
pair becomes two shape indices and two 6-vectors ``(r, p)``.

Pure numpy; no device code here.
"""
from __future__ import annotations

import numpy as np

# primitive kinds, in the order of SURVEY.md section 8 / include/dcol.h
POLYTOPE, CAPSULE, CYLINDER, CONE, SPHERE, POLYGON, ELLIPSOID = range(7)
N_KINDS = 7      # ELLIPSOID is an extension: the reference's code has no such primitive (only its report does)
KIND_NAMES = ("polytope", "capsule", "cylinder", "cone", "sphere", "polygon", "ellipsoid")

#: extra decision variables a primitive adds beyond (x, alpha)
KIND_EXTRAS = (0, 1, 1, 0, 0, 2, 0)
#: second-order-cone dimension of a primitive (0 = none)
KIND_SOC = (0, 4, 4, 3, 4, 4, 4)

#: most half-spaces one polytope / polygon may carry (device row arrays are sized for it)
MAX_FACES = 32

_CLASS_TO_KIND = {
    "PolytopeMRP": POLYTOPE, "CapsuleMRP": CAPSULE, "CylinderMRP": CYLINDER,
    "ConeMRP": CONE, "SphereMRP": SPHERE, "PolygonMRP": POLYGON, "EllipsoidMRP": ELLIPSOID,
}

#: numpy mirror of ``struct dcol_shape`` (144 bytes, natural alignment)
SHAPE_DTYPE = np.dtype([
    ("type", "<i4"), ("n_faces", "<i4"), ("face_off", "<i4"), ("reserved", "<i4"),
    ("R", "<f8"), ("L", "<f8"), ("H", "<f8"), ("beta", "<f8"),
    ("r_offset", "<f8", (3,)), ("Q_offset", "<f8", (9,)),
], align=True)
assert SHAPE_DTYPE.itemsize == 144

# per-pair status words (include/dcol.h)
STATUS_OK, STATUS_MAX_ITER, STATUS_NON_FINITE, STATUS_NOT_PD, STATUS_UNSUPPORTED = range(5)


def kind_of(prim) -> int:
    """Primitive kind from the class *name*, so the reference's own classes are accepted."""
    for klass in type(prim).__mro__:
        k = _CLASS_TO_KIND.get(klass.__name__)
        if k is not None:
            return k
    raise TypeError(f"not a DCOL primitive: {type(prim).__name__}")


def n_ort_of(kind: int, n_faces: int) -> int:
    """Orthant rows a primitive contributes (problem_matrices.py:37-42,78-85,109,146,166,202)."""
    return (n_faces, 2, 4, 1, 0, n_faces, 0)[kind]


def pair_supported(kind1: int, kind2: int) -> bool:
    """False for the pairs the reference cannot assemble (both carry extra variables;
    combine_problem_matrices.py:58-67 raises ValueError)."""
    return not (KIND_EXTRAS[kind1] > 0 and KIND_EXTRAS[kind2] > 0)


def pose_of(prim) -> np.ndarray:
    """``[r(3), p(3)]`` of a primitive as float64 (``.r``/``.p`` may be lists)."""
    out = np.empty(6)
    out[:3] = np.asarray(prim.r, dtype=float).reshape(3)
    out[3:] = np.asarray(prim.p, dtype=float).reshape(3)
    return out


def flatten_shapes(prims):
    """Flatten primitive objects into ``(records[SHAPE_DTYPE], A[nf,3], b[nf])``.

    Polytope faces keep all three columns of ``A``; polygon faces use the first two
    (third column zero).  Pose (``r``, ``p``) is *not* part of the shape.
    """
    recs = np.zeros(len(prims), dtype=SHAPE_DTYPE)
    A_rows, b_rows = [], []
    off = 0
    for i, prim in enumerate(prims):
        kind = kind_of(prim)
        rec = recs[i]
        rec["type"] = kind
        rec["r_offset"] = np.asarray(prim.r_offset, dtype=float).reshape(3)
        rec["Q_offset"] = np.asarray(prim.Q_offset, dtype=float).reshape(9)
        if kind in (POLYTOPE, POLYGON):
            A = np.asarray(prim.A, dtype=float)
            b = np.asarray(prim.b, dtype=float).reshape(-1)
            cols = 3 if kind == POLYTOPE else 2
            if A.ndim != 2 or A.shape[1] != cols or A.shape[0] != b.shape[0]:
                raise ValueError(f"{KIND_NAMES[kind]}: A must be (f,{cols}) and b (f,), got {A.shape}, {b.shape}")
            f = A.shape[0]
            if not 1 <= f <= MAX_FACES:
                raise ValueError(f"{KIND_NAMES[kind]} with {f} faces: supported range is 1..{MAX_FACES}")
            A3 = np.zeros((f, 3))
            A3[:, :cols] = A
            A_rows.append(A3)
            b_rows.append(b)
            rec["n_faces"] = f
            rec["face_off"] = off
            off += f
        if kind in (CAPSULE, CYLINDER, SPHERE, POLYGON):
            rec["R"] = float(prim.R)
        if kind in (CAPSULE, CYLINDER):
            rec["L"] = float(prim.L)
        if kind == CONE:
            rec["H"] = float(prim.H)
            rec["beta"] = float(prim.beta)
        if kind == ELLIPSOID:               # semi-axes along the body axes travel in (R, L, H)
            rec["R"], rec["L"], rec["H"] = (float(v) for v in prim.semi_axes)
            if min(rec["R"], rec["L"], rec["H"]) <= 0:
                raise ValueError("ellipsoid semi-axes must be positive")
    A_packed = np.concatenate(A_rows) if A_rows else np.zeros((0, 3))
    b_packed = np.concatenate(b_rows) if b_rows else np.zeros((0,))
    return recs, np.ascontiguousarray(A_packed), np.ascontiguousarray(b_packed)


def problem_dims(rec1, rec2):
    """``(m_ort, q1, q2, n)`` of the conic program a shape pair assembles
    (combine_problem_matrices.py:22-32); ``n`` counts x(3), alpha and the extras."""
    k1, k2 = int(rec1["type"]), int(rec2["type"])
    m_ort = n_ort_of(k1, int(rec1["n_faces"])) + n_ort_of(k2, int(rec2["n_faces"]))
    return m_ort, KIND_SOC[k1], KIND_SOC[k2], 4 + KIND_EXTRAS[k1] + KIND_EXTRAS[k2]


def flop_model(m_ort: int, q1: int, q2: int, n: int, f1: int, f2: int, is_poly1: bool, is_poly2: bool,
               iters: int) -> int:
    """Algorithmic FP64 flops of one solve+gradient (the contract figure).

    add/sub/mul/div/sqrt = 1, FMA = 2, compares = 0.  ``flops = F_A + F_0 + iters*F_it + F_T + F_G``.
    """
    socs = [q for q in (q1, q2) if q > 0]
    m = m_ort + q1 + q2
    Q2 = q1 * q1 + q2 * q2
    aW = m_ort + 2 * Q2
    cp = m_ort + sum(5 * q - 4 for q in socs)
    icp = m_ort + sum(9 * q - 3 for q in socs)
    ls = 2 * m_ort + sum(10 * q + 7 for q in socs)
    nt = 2 * m_ort + sum(3 * q * q + 9 * q + 16 for q in socs)
    n3 = n ** 3 // 3
    F_it = (nt + Q2 + aW + cp + (4 * m * n + 4 * m + n + 1) + n * aW + m * n * (n + 1) + (n3 + 2 * n)
            + 2 * (icp + 5 * aW + 4 * m * n + 2 * n * n + 3 * m + n) + 4 * ls + (6 * m + 5)
            + (2 * aW + cp + 3 * m) + (4 * m + 2 * n + 2))
    F_T = nt + aW + cp + (4 * m * n + 4 * m + n + 1)
    F_0 = m * n * (n + 1) + n3 + 2 * n + 6 * m * n + 3 * n * n + n + m + 2 * (m_ort + sum(2 * q for q in socs) + m)
    F_A = 80 + (20 * f1 + 60 if is_poly1 else 90) + (20 * f2 + 60 if is_poly2 else 90)
    F_G = 24 * m + 360
    return F_A + F_0 + iters * F_it + F_T + F_G

"""Primitive data model accepted by the proximity drop-in.

Mirrors the attribute bags of the reference's
``primitives/misc_primitive_constructor.py:4-88`` (same class names, constructor
arguments and attribute names, so the reference's ``systems/*.py`` scripts can
build their scenes with either module) plus its two shape constructors
``create_rect_prism`` (``:91-142``) and ``create_n_sided`` (``:145-164``).

Every primitive carries a pose ``r`` (position), ``p`` (modified Rodrigues
parameters) and a body-frame offset ``r_offset`` / ``Q_offset``.  Callers mutate
``.r`` / ``.p`` in place (sometimes with plain Python lists, e.g.
``systems/piano_mover.py:176-178``); the host flattening in ``shapes.py``
therefore reads them through ``numpy.asarray`` at call time.
"""
import numpy as np

__all__ = [
    "PolytopeMRP", "CapsuleMRP", "CylinderMRP", "ConeMRP", "SphereMRP", "PolygonMRP", "EllipsoidMRP",
    "create_rect_prism", "create_n_sided",
]


class _PrimitiveMRP:
    """Pose + body-frame offset shared by all primitive kinds."""

    def __init__(self):
        self.r = np.zeros(3)
        self.p = np.zeros(3)
        self.r_offset = np.zeros(3)
        self.Q_offset = np.eye(3)


class PolytopeMRP(_PrimitiveMRP):
    """Convex polytope ``{y : A y <= alpha * b}`` in the body frame (f faces)."""

    def __init__(self, A, b, length=0, width=0, height=0):
        super().__init__()
        self.A = A
        self.b = b
        self.length = length
        self.width = width
        self.height = height


class EllipsoidMRP(_PrimitiveMRP):
    """EXTENSION (not in the reference's code; its report describes it, sec. 3.1.5): ellipsoid with semi-axes
    ``(a, b, c)`` along the body axes, ``{y : |diag(1/a, 1/b, 1/c) y| <= alpha}``.  A general ellipsoid
    ``y' P y <= 1`` is expressed by putting P's eigenvectors into ``Q_offset``."""

    def __init__(self, a, b, c):
        super().__init__()
        self.semi_axes = (a, b, c)


class CapsuleMRP(_PrimitiveMRP):
    """Capsule of radius ``R`` around a segment of length ``L`` on the body x axis."""

    def __init__(self, radius, height):
        super().__init__()
        self.R = radius
        self.L = height


class CylinderMRP(_PrimitiveMRP):
    """Cylinder of radius ``R`` and length ``L`` along the body x axis."""

    def __init__(self, radius, height):
        super().__init__()
        self.R = radius
        self.L = height


class ConeMRP(_PrimitiveMRP):
    """Cone of height ``H`` and half angle ``beta`` along the body x axis."""

    def __init__(self, height, beta):
        super().__init__()
        self.H = height
        self.beta = beta


class SphereMRP(_PrimitiveMRP):
    """Sphere of radius ``R``."""

    def __init__(self, radius):
        super().__init__()
        self.R = radius


class PolygonMRP(_PrimitiveMRP):
    """Planar polygon ``{y in R^2 : A y <= alpha * b}`` (body x-y plane) padded by radius ``R``."""

    def __init__(self, A, b, radius):
        super().__init__()
        self.A = A
        self.b = b
        self.R = radius


def create_rect_prism(length=20.0, width=20.0, height=2.0, attitude="MRP"):
    """Axis-aligned box as a 6-face polytope, faces ordered +x,+y,+z,-x,-y,-z.

    Same face order and ``b = half extent`` as the reference
    (``misc_primitive_constructor.py:106-130``); only the MRP attitude exists.
    """
    if attitude != "MRP":
        raise ValueError("Attitude must be 'MRP'")
    half = 0.5 * np.array([length, width, height], dtype=float)
    A = np.zeros((6, 3))
    for k in range(3):
        A[k, k] = 1.0
        A[k + 3, k] = -1.0
    b = np.concatenate([half, half])
    # the reference forms b as dot(normal, centre); -1 * -half == half exactly
    return PolytopeMRP(A, b, length=length, width=width, height=height)


def create_n_sided(N, d):
    """Regular N-gon: outward normals at angles 2*pi*k/N, every side at distance ``d``.

    Returns the same ``{"A": (N,2), "b": (N,)}`` dictionary as the reference
    (``misc_primitive_constructor.py:145-164``).
    """
    angles = np.linspace(0, 2 * np.pi, N, endpoint=False)
    A = np.stack([np.cos(angles), np.sin(angles)], axis=1)
    return {"A": A, "b": np.full(N, d)}

"""Primitive data model (mirror of the reference's ``primitives`` package, data classes only)."""
from .misc_primitive_constructor import (  # noqa: F401
    PolytopeMRP, CapsuleMRP, CylinderMRP, ConeMRP, SphereMRP, PolygonMRP, EllipsoidMRP,
    create_rect_prism, create_n_sided,
)

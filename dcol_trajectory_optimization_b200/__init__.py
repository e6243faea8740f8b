"""B200-native batched DCOL proximity solve + gradient.

Importing the package does not load the CUDA library; the first engine does, and raises if the
extension has not been built or no CUDA device is visible (there is no CPU fallback)."""
from .engine import (BatchResult, Plan, ProximityEngine, SceneResult, measure_fp64_peak, pinned_empty, pinned_free, release_cached,  # noqa: F401
                     proximity_batch, raise_for_status)
from .primitives import (CapsuleMRP, ConeMRP, CylinderMRP, EllipsoidMRP, PolygonMRP, PolytopeMRP, SphereMRP,  # noqa: F401
                         create_n_sided, create_rect_prism)

__version__ = "0.1.0"

"""``proximity.proximity_gradient`` of the reference (proximity/proximity_gradient.py:91-138), backed by the CUDA library."""
from . import _REPO  # noqa: F401  (sys.path bootstrap)
from dcol_trajectory_optimization_b200.proximity.proximity_gradient import proximity_gradient  # noqa: E402,F401

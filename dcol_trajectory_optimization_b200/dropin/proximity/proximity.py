"""``proximity.proximity`` of the reference (proximity/proximity.py:6-54), backed by the CUDA library."""
from . import _REPO  # noqa: F401  (sys.path bootstrap)
from dcol_trajectory_optimization_b200.proximity.proximity import proximity_mrp  # noqa: E402,F401

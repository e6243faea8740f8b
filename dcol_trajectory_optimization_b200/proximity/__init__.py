"""Mirror of the reference's ``proximity`` package: same module and function names
(``proximity.proximity.proximity_mrp``, ``proximity.proximity_gradient.proximity_gradient``), backed
by the CUDA library.  For zero-edit use from the reference's scripts put
``dcol_trajectory_optimization_b200/dropin`` ahead of the reference on ``sys.path``."""
from .proximity import proximity_mrp  # noqa: F401
from .proximity_gradient import proximity_gradient  # noqa: F401

"""Zero-edit drop-in for the reference's ``proximity`` package.

Put THIS directory's parent (``dcol_trajectory_optimization_b200/dropin``) ahead of the reference on
``sys.path``: the reference's ``systems/*.py`` do ``from proximity.proximity import proximity_mrp`` and
``from proximity.proximity_gradient import proximity_gradient`` (e.g. ``systems/piano_mover.py:2-3``) and then
get the CUDA path, while ``primitives``, ``ALTRO``, ``utils`` and ``systems`` keep resolving to the reference
(its primitive classes are accepted as they are: shapes are classified by class name)."""
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _REPO not in sys.path:          # make the real package importable by its own name
    sys.path.append(_REPO)

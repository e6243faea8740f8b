"""TEST INFRASTRUCTURE, generation time only: import the read-only Python reference.

Used by ``oracle/gen_golden.py`` (and nothing that runs on the GPU box) to produce the
committed fixtures under ``tests/golden/``.  ``/root/reference`` does not exist on the GPU
box, so no test, ``smoke()`` or ``bench.py`` imports this module.

The reference's hot path (L0-L2 of SURVEY.md) needs only numpy + scipy.  ``ALTRO.py`` and
the systems additionally import matplotlib / h5py / meshcat, which are absent here; they
are stubbed (recipe: SURVEY.md appendix B).
"""
import os
import sys
import types
import warnings
from unittest.mock import MagicMock

import numpy as np

REFERENCE_ROOT = os.environ.get("DCOL_REFERENCE_ROOT", "/root/reference")


class _FakeH5File:
    """Reads the four compact datasets of systems/polytopes.jld2 by byte offset."""

    def __init__(self, *a, **k):
        raw = open(os.path.join(REFERENCE_ROOT, "systems", "polytopes.jld2"), "rb").read()

        def f8(off, n):
            return np.frombuffer(raw[off:off + 8 * n], "<f8").copy()

        self._d = {"A1": f8(630, 42).reshape(3, 14), "b1": f8(1031, 14),
                   "A2": f8(1216, 24).reshape(3, 8), "b2": f8(1473, 8)}

    def __enter__(self):
        return self._d

    def __exit__(self, *a):
        return False


def import_reference():
    """Put the reference on sys.path (with plotting / h5py stubs) and return its hot-path modules."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "mpl_toolkits",
                 "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d", "meshcat"):
        sys.modules.setdefault(name, MagicMock())
    if "h5py" not in sys.modules:
        h5 = types.ModuleType("h5py")
        h5.File = _FakeH5File
        sys.modules["h5py"] = h5
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        import proximity.pdip as pdip
        import proximity.proximity as prox
        import proximity.proximity_gradient as prox_grad
        import primitives.misc_primitive_constructor as prims
        import primitives.problem_matrices as pm
        import primitives.combine_problem_matrices as cpm
    return types.SimpleNamespace(pdip=pdip, proximity=prox, proximity_gradient=prox_grad,
                                 primitives=prims, problem_matrices=pm, combine=cpm)

"""``proximity_mrp`` — drop-in for the reference's ``proximity/proximity.py:6-54``."""
import numpy as np

from ._scalar import solve_one


def proximity_mrp(prim1, prim2, pdip_tol=1e-6, verbose=False):
    """Proximity value ``alpha`` and contact point of two primitives.

    Same signature, argument meaning and return as the reference: ``(alpha, x[:3])`` with
    ``alpha = x[3]`` of the conic program's solution (``proximity.py:51-54``); ``verbose`` is ignored
    there too.  ``alpha > 1``: separated, ``alpha < 1``: penetrating.  Failures raise the exception
    classes of the reference (max iterations -> ``Exception``, non-finite -> ``ValueError``, Cholesky
    -> ``numpy.linalg.LinAlgError``, both primitives with extra variables -> ``ValueError``)."""
    res = solve_one(prim1, prim2, pdip_tol, want_grad=False)
    return np.float64(res.alpha[0]), res.contact[0].copy()

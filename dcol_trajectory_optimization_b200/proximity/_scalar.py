"""Shared machinery of the two scalar drop-ins: a small cache of single-pair engines keyed by the
two shapes (the reference re-dispatches on isinstance at every call; scenes re-use a handful of
shapes thousands of times)."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from ..engine import ProximityEngine, raise_for_status
from ..shapes import flatten_shapes, pose_of

_CACHE: "OrderedDict[bytes, ProximityEngine]" = OrderedDict()
_CACHE_MAX = 256
_IDX0 = np.zeros(1, dtype=np.int32)
_IDX1 = np.ones(1, dtype=np.int32)


def _engine_for(prim1, prim2) -> ProximityEngine:
    rec, A, b = flatten_shapes([prim1, prim2])
    key = rec.tobytes() + A.tobytes() + b.tobytes()
    eng = _CACHE.get(key)
    if eng is None:
        eng = ProximityEngine((rec, A, b))
        _CACHE[key] = eng
        if len(_CACHE) > _CACHE_MAX:
            _CACHE.popitem(last=False)[1].close()
    else:
        _CACHE.move_to_end(key)
    return eng


def solve_one(prim1, prim2, pdip_tol, want_grad):
    """One pair through the host entry point of the C ABI; raises what the reference raises."""
    eng = _engine_for(prim1, prim2)
    res = eng.solve_host(_IDX0, _IDX1, pose_of(prim1)[None, :], pose_of(prim2)[None, :], tol=pdip_tol,
                         want_grad=want_grad, want_contact=True)
    raise_for_status(int(res.status[0]))
    return res

"""The C-ABI library builds, loads and exports every symbol include/dcol.h declares; host-side logic
(shape flattening, workloads, error convention) — no compute calls, no GPU needed."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from dcol_trajectory_optimization_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_library_exports_every_declared_symbol(L):
    from dcol_trajectory_optimization_b200 import _lib
    header = open(os.path.join(ROOT, "include", "dcol.h")).read()
    declared = set(re.findall(r"\b(dcol_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(L, name) is not None
    assert b"sm_100a" in L.dcol_version()


def test_no_gpu_fails_loudly(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import dcol_trajectory_optimization_b200 as d
    assert L.dcol_device_count() == 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.ProximityEngine([d.SphereMRP(1.0)])
    from dcol_trajectory_optimization_b200.proximity import proximity_mrp
    with pytest.raises(RuntimeError):
        proximity_mrp(d.SphereMRP(1.0), d.SphereMRP(1.0))


def test_shape_struct_layout_matches_header():
    from dcol_trajectory_optimization_b200.shapes import SHAPE_DTYPE
    assert SHAPE_DTYPE.itemsize == 144
    assert [SHAPE_DTYPE.fields[k][1] for k in ("type", "n_faces", "face_off", "R", "L", "H", "beta", "r_offset",
                                               "Q_offset")] == [0, 4, 8, 16, 24, 32, 40, 48, 72]


def test_flatten_shapes_and_workloads():
    import dcol_trajectory_optimization_b200 as d
    from dcol_trajectory_optimization_b200 import shapes as S, workloads as W
    shp = W.config4_shapes()
    rec, A, b = S.flatten_shapes(shp)
    assert list(rec["type"]) == [0, 0, 1, 2, 3, 4, 5] and list(rec["n_faces"]) == [6, 8, 0, 0, 0, 0, 5]
    assert A.shape == (19, 3) and np.all(A[14:, 2] == 0)         # polygon faces use two columns
    assert len(W.supported_type_pairs(shp)) == 40
    _, i1, i2, p1, p2 = W.config4_batch(1000)
    assert p1.shape == (1000, 6) and np.all(np.linalg.norm(p1[:, :3], axis=1) <= 1.0 + 1e-12)
    _, j1, j2, q1, q2 = W.config5_batch(n_obs=22, n_knots=5, n_cand=3)
    assert len(j1) == 330 and np.all(j1 == 0) and set(j2) == set(range(1, 12))
    box = d.create_rect_prism(1, 2, 3)
    box.r = [1, 2, 3]                                             # callers assign lists (piano_mover.py:176)
    assert np.array_equal(S.pose_of(box), [1, 2, 3, 0, 0, 0])
    with pytest.raises(ValueError):
        S.flatten_shapes([d.PolytopeMRP(np.zeros((40, 3)), np.zeros(40))])   # > DCOL_MAX_FACES
    with pytest.raises(TypeError):
        S.flatten_shapes([object()])


def test_status_to_exception_mapping():
    from dcol_trajectory_optimization_b200 import raise_for_status
    raise_for_status(0)
    with pytest.raises(Exception, match="Maximum number of iterations reached, PDIP failed"):
        raise_for_status(1)
    with pytest.raises(ValueError):
        raise_for_status(2)
    with pytest.raises(np.linalg.LinAlgError):
        raise_for_status(3)
    with pytest.raises(ValueError):
        raise_for_status(4)


def test_flop_model_examples():
    """F_it of box x box = 1,585, cone x box = 1,851, sphere x sphere = 2,741."""
    from dcol_trajectory_optimization_b200.shapes import flop_model
    fit = lambda *a: flop_model(*a, 1) - flop_model(*a, 0)
    assert fit(12, 0, 0, 4, 6, 6, True, True) == 1585
    assert fit(7, 3, 0, 4, 0, 6, False, True) == 1851
    assert fit(0, 4, 4, 4, 0, 0, False, False) == 2741


def test_batch_result_summary():
    from dcol_trajectory_optimization_b200 import BatchResult
    r = BatchResult(alpha=np.zeros(6), contact=None, grad=None, iters=np.array([5, 6, 6, 0, 50, 7], np.int32),
                    status=np.array([0, 0, 0, 4, 1, 0], np.int32))
    s = r.summary()
    assert s["pairs"] == 6 and s["status"] == {"ok": 4, "max_iter": 1, "non_finite": 0, "not_pd": 0, "unsupported": 1}
    assert s["iters_hist"] == {5: 1, 6: 2, 7: 1} and s["iters_max"] == 7 and abs(s["iters_mean"] - 6.0) < 1e-12

"""Randomised shapes: the kernel header (host twin) against the oracle on polytopes with 4..32 random faces, polygons
with 3..12 sides, random body-frame offsets (r_offset, Q_offset) and random sizes of every primitive kind — the
runtime-face-count specialisations and the offset algebra, which the reference's scenes (6/8 faces, zero offsets)
never exercise.  The reference's data model carries all of these (misc_primitive_constructor.py:4-88,
problem_matrices.py:272-364); the oracle restates them operation by operation and is pinned to reference-generated
goldens that include offset cases (tests/golden/edge_cases.npz)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "host_twin"))

from dcol_trajectory_optimization_b200.primitives import (CapsuleMRP, ConeMRP, CylinderMRP, EllipsoidMRP, PolygonMRP,  # noqa: E402
                                                          PolytopeMRP, SphereMRP, create_rect_prism)
from dcol_trajectory_optimization_b200.shapes import flatten_shapes  # noqa: E402


def _random_rotation(rng):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    return q * np.sign(np.linalg.det(q))


def _random_polytope(rng, nf):
    """nf half-spaces a_i . y <= b_i with b_i > 0 (origin inside); the first six directions are a rotated +-axes
    frame when nf >= 6 so that the body is bounded, a random tetrahedron otherwise."""
    if nf >= 6:
        R = _random_rotation(rng)
        dirs = np.vstack([R, -R, rng.normal(size=(nf - 6, 3))])
    else:
        base = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], float)
        dirs = np.vstack([base @ _random_rotation(rng).T, rng.normal(size=(nf - 4, 3))])
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    return PolytopeMRP(dirs, rng.uniform(0.4, 1.6, size=nf))


def _random_polygon(rng, n):
    ang = np.sort((np.arange(n) + rng.uniform(-0.3, 0.3, size=n)) * 2 * np.pi / n)
    A = np.stack([np.cos(ang), np.sin(ang)], axis=1)
    return PolygonMRP(A, rng.uniform(0.3, 0.9, size=n), rng.uniform(0.05, 0.4))


def _shape_zoo(rng, offsets):
    shapes = [_random_polytope(rng, nf) for nf in (4, 5, 6, 7, 8, 9, 14, 21, 32)]
    shapes += [_random_polygon(rng, n) for n in (3, 4, 5, 6, 9, 12)]
    shapes += [create_rect_prism(*rng.uniform(0.3, 3.0, size=3)), CapsuleMRP(rng.uniform(0.1, 0.6), rng.uniform(0.5, 3.0)),
               CylinderMRP(rng.uniform(0.2, 0.8), rng.uniform(0.5, 3.0)), ConeMRP(rng.uniform(0.8, 3.0), rng.uniform(0.2, 0.6)),
               SphereMRP(rng.uniform(0.2, 1.2)), EllipsoidMRP(*rng.uniform(0.3, 1.5, size=3))]
    if offsets:
        for s in shapes:
            s.r_offset = rng.normal(size=3) * 0.3
            s.Q_offset = _random_rotation(rng)
    return shapes


def _batch(rng, shapes, B):
    ns = len(shapes)
    i1 = rng.integers(0, ns, size=B).astype(np.int32)
    i2 = rng.integers(0, ns, size=B).astype(np.int32)
    u = rng.normal(size=(B, 3))
    p1 = np.concatenate([rng.normal(size=(B, 3)) * 0.5, rng.normal(size=(B, 3)) * 0.5], axis=1)
    p2 = np.concatenate([u / np.linalg.norm(u, axis=1, keepdims=True) * rng.uniform(0.0, 6.0, size=(B, 1)),
                         rng.normal(size=(B, 3)) * 0.5], axis=1)
    return i1, i2, p1, p2


@pytest.fixture(scope="module")
def twin():
    import twin as T
    T.build()
    return T


@pytest.mark.parametrize("offsets", [False, True])
def test_twin_follows_oracle_on_random_shapes(twin, oracle, offsets):
    rng = np.random.default_rng(2024 + offsets)
    shapes = _shape_zoo(rng, offsets)
    rec, A, b = flatten_shapes(shapes)
    i1, i2, p1, p2 = _batch(rng, shapes, 30_000)
    ref = oracle.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=oracle.GRAD_EXACT, fix_case4=True)
    out = twin.solve_batch(rec, A, b, i1, i2, p1, p2, fix_case4=True)
    assert np.array_equal(out["status"], ref["status"])
    ok = ref["status"] == 0
    assert ok.mean() > 0.995                         # thin random slivers may fail in the reference too; same flag
    flips = out["iters"][ok] != ref["iters"][ok]
    assert flips.mean() < 2e-4, flips.sum()          # a mu that lands on the tolerance to the last bits
    same = ok & (out["iters"] == ref["iters"])
    rel = np.abs(out["alpha"][same] - ref["alpha"][same]) / np.maximum(np.abs(ref["alpha"][same]), 1.0)
    assert rel.max() < 1e-8, rel.max()
    gscale = np.maximum(np.abs(ref["grad"][same]).max(axis=1), 1e-12)
    gerr = np.abs(out["grad"][same] - ref["grad"][same]).max(axis=1) / gscale
    assert np.median(gerr) < 1e-9 and np.quantile(gerr, 0.999) < 1e-6, (np.median(gerr), np.quantile(gerr, 0.999))
    cscale = np.maximum(np.abs(ref["contact"][same]).max(axis=1), 1.0)
    cerr = np.abs(out["contact"][same] - ref["contact"][same]).max(axis=1) / cscale
    assert np.quantile(cerr, 0.999) < 1e-7            # non-unique contact points (parallel faces) move freely


def test_unsupported_pairs_without_the_extension_flag(twin, oracle):
    rng = np.random.default_rng(5)
    shapes = _shape_zoo(rng, True)
    rec, A, b = flatten_shapes(shapes)
    i1, i2, p1, p2 = _batch(rng, shapes, 4000)
    ref = oracle.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=oracle.GRAD_NONE)
    out = twin.solve_batch(rec, A, b, i1, i2, p1, p2, want_grad=False)
    assert np.array_equal(out["status"], ref["status"]) and (out["status"] == 4).sum() > 100
    assert np.isnan(out["alpha"][out["status"] == 4]).all()


def test_jacobian_on_random_shapes_matches_dense_kkt(twin, oracle):
    from test_jacobian import dense_jacobian
    rng = np.random.default_rng(9)
    shapes = _shape_zoo(rng, True)
    rec, A, b = flatten_shapes(shapes)
    i1, i2, p1, p2 = _batch(rng, shapes, 300)
    keep = np.array([not (rec["type"][a] in (1, 2, 5) and rec["type"][c] in (1, 2, 5)) for a, c in zip(i1, i2)])
    i1, i2, p1, p2 = i1[keep], i2[keep], p1[keep], p2[keep]     # the oracle's assembly entry point has no case-4 layout
    out = twin.solve_batch(rec, A, b, i1, i2, p1, p2, want_jac=True)
    errs = []
    for k in np.flatnonzero(out["status"] == 0):
        r = twin.solve_pair(rec, A, b, i1[k], i2[k], p1[k], p2[k])
        Jd = dense_jacobian(rec, A, b, i1[k], i2[k], p1[k], p2[k], r["x"], r["s"], r["z"])
        errs.append(np.abs(out["jac"][k] - Jd).max() / max(1.0, np.abs(Jd).max()))
    errs = np.array(errs)
    assert len(errs) > 150 and np.median(errs) < 1e-8 and np.quantile(errs, 0.98) < 1e-5, (np.median(errs), errs.max())


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("offsets", [False, True])
def test_cuda_follows_oracle_on_random_shapes(oracle, offsets):
    """The same property test through the C ABI on the GPU: runtime-face-count classes (``polyn`` 4..32 faces,
    ``pgonn`` 3..12 sides), body-frame offsets, random sizes, all 21 x 21 ordered shape pairs, against the oracle."""
    import dcol_trajectory_optimization_b200 as d
    rng = np.random.default_rng(2024 + offsets)
    shapes = _shape_zoo(rng, offsets)
    rec, A, b = flatten_shapes(shapes)
    i1, i2, p1, p2 = _batch(rng, shapes, 100_000)
    ref = oracle.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=oracle.GRAD_EXACT, fix_case4=True)
    eng = d.ProximityEngine((rec, A, b))
    out = eng.solve_host(i1, i2, p1, p2, fix_case4=True)
    eng.close()
    assert np.array_equal(out.status, ref["status"])
    ok = ref["status"] == 0
    assert ok.mean() > 0.995
    flips = out.iters[ok] != ref["iters"][ok]
    assert flips.mean() < 2e-4, flips.sum()
    same = ok & (out.iters == ref["iters"])
    rel = np.abs(out.alpha[same] - ref["alpha"][same]) / np.maximum(np.abs(ref["alpha"][same]), 1.0)
    # 1e-8 is the bar on the reference's shapes; random slivers (faces at 0.4..1.6 around a point, 32 of them) produce a few
    # ill-conditioned pairs per 1e5 in which rounding alone moves alpha by a few 1e-8 at an unchanged iteration count
    assert np.quantile(rel, 0.9999) < 1e-8 and rel.max() < 1e-6, (np.quantile(rel, 0.9999), rel.max())
    gscale = np.maximum(np.abs(ref["grad"][same]).max(axis=1), 1e-12)
    gerr = np.abs(out.grad[same] - ref["grad"][same]).max(axis=1) / gscale
    assert np.median(gerr) < 1e-9 and np.quantile(gerr, 0.999) < 1e-6, (np.median(gerr), np.quantile(gerr, 0.999))
    cscale = np.maximum(np.abs(ref["contact"][same]).max(axis=1), 1.0)
    cerr = np.abs(out.contact[same] - ref["contact"][same]).max(axis=1) / cscale
    assert np.quantile(cerr, 0.999) < 1e-7


@pytest.mark.gpu
def test_cuda_face_count_limit_is_reported():
    """DCOL_MAX_FACES = 32 half-spaces per polytope / polygon (the reference has no limit; its largest shape, A1 of
    systems/polytopes.jld2, has 14): a larger shape is rejected when the table is created, with DCOL_E_SHAPE."""
    import dcol_trajectory_optimization_b200 as d
    from dcol_trajectory_optimization_b200._lib import DcolError
    rng = np.random.default_rng(1)
    ok = d.ProximityEngine([_random_polytope(rng, 32), d.SphereMRP(0.5)])
    r = ok.solve_host([0], [1], np.zeros((1, 6)), np.array([[4.0, 0, 0, 0, 0, 0]]))
    assert r.status[0] == 0
    ok.close()
    with pytest.raises(ValueError):                     # the Python host refuses to flatten it ...
        d.ProximityEngine([_random_polytope(rng, 33), d.SphereMRP(0.5)])
    rec, A, b = flatten_shapes([_random_polytope(rng, 32), d.SphereMRP(0.5)])
    rec = rec.copy()
    rec["n_faces"][0] = 33                              # ... and so does the C ABI for a raw record
    A33, b33 = np.vstack([A, A[:1]]), np.concatenate([b, b[:1]])
    with pytest.raises(DcolError) as e:
        d.ProximityEngine((rec, A33, b33))
    assert e.value.code == -2

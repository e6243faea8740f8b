"""dcol_plan_refine: re-ordering a plan by the previous solve's iteration counts changes which thread takes which
pair and nothing else (GPU only: the plan lives on the device)."""
import numpy as np
import pytest

from dcol_trajectory_optimization_b200 import workloads as W
from dcol_trajectory_optimization_b200.shapes import flatten_shapes


@pytest.mark.gpu
def test_refine_keeps_results_and_sorts_groups_by_iterations():
    import torch
    import dcol_trajectory_optimization_b200 as d
    B = 200_003                                   # ragged: not a multiple of the plan tile or the warp
    shapes, i1, i2, p1, p2 = W.config4_batch(B, seed=33)
    rec, A, b = flatten_shapes(shapes)
    eng = d.ProximityEngine((rec, A, b), device=0)
    plan = eng.plan(i1, i2)
    P1, P2 = torch.as_tensor(p1, device="cuda"), torch.as_tensor(p2, device="cuda")
    r0 = eng.solve(plan, P1, P2)
    torch.cuda.synchronize()
    perm0 = plan.perm().cpu().numpy()
    key = i1.astype(np.int64) * len(rec) + i2
    it0 = r0.iters.cpu().numpy()

    def check_perm(perm, iters_used):
        assert np.array_equal(np.sort(perm), np.arange(B))
        k = key[perm]
        assert (np.diff(k) >= 0).all()                                  # groups untouched, in the same order
        assert np.array_equal(k, key[perm0])
        same = np.diff(k) == 0
        assert (np.diff(np.clip(iters_used[perm], 0, 63))[same] <= 0).all()   # inside every group: longest first

    for _ in range(2):                                                  # the two permutation buffers swap
        plan.refine(r0.iters)
        check_perm(plan.perm().cpu().numpy(), it0)
        r1 = eng.solve(plan, P1, P2)
        torch.cuda.synchronize()
        for name in ("alpha", "contact", "grad", "iters", "status"):
            a, c = getattr(r0, name).cpu().numpy(), getattr(r1, name).cpu().numpy()
            assert np.array_equal(a, c, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, c), name
    # arbitrary keys are clamped, never trusted
    junk = torch.randint(-5, 500, (B,), dtype=torch.int32, device="cuda")
    plan.refine(junk)
    check_perm(plan.perm().cpu().numpy(), junk.cpu().numpy())
    r2 = eng.solve(plan, P1, P2)
    torch.cuda.synchronize()
    assert np.array_equal(r2.iters.cpu().numpy(), it0) and np.array_equal(r2.alpha.cpu().numpy(), r0.alpha.cpu().numpy())
    with pytest.raises(ValueError):
        plan.refine(r0.iters[:10])
    eng.close()


@pytest.mark.gpu
def test_refine_single_group_and_tiny_plans():
    import torch
    import dcol_trajectory_optimization_b200 as d
    from dcol_trajectory_optimization_b200.primitives import SphereMRP, create_rect_prism
    rec, A, b = flatten_shapes([SphereMRP(0.5), create_rect_prism(1, 2, 3)])
    eng = d.ProximityEngine((rec, A, b), device=0)
    rng = np.random.default_rng(0)
    for B in (1, 31, 5000):
        i1, i2 = np.zeros(B, np.int32), np.ones(B, np.int32)
        p1 = np.concatenate([rng.normal(size=(B, 3)) * 3, rng.normal(size=(B, 3)) * 0.3], axis=1)
        p2 = np.zeros((B, 6))
        plan = eng.plan(i1, i2)
        P1, P2 = torch.as_tensor(p1, device="cuda"), torch.as_tensor(p2, device="cuda")
        r0 = eng.solve(plan, P1, P2)
        plan.refine(r0.iters)
        r1 = eng.solve(plan, P1, P2)
        torch.cuda.synchronize()
        perm = plan.perm().cpu().numpy()
        assert np.array_equal(np.sort(perm), np.arange(B))
        assert (np.diff(r0.iters.cpu().numpy()[perm]) <= 0).all()
        assert np.array_equal(r0.alpha.cpu().numpy(), r1.alpha.cpu().numpy(), equal_nan=True)
        plan.close()
    eng.close()

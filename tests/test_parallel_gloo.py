"""The N > 1 path on CPU: world_size-2 gloo processes shard a batch, fill the packed record buffers
and exchange them with the single all-gather.  The local solve is injected (here: the oracle, as the
checker) because the product itself has no CPU path."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_pairs, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import torch.distributed as dist
    import oracle as O
    from dcol_trajectory_optimization_b200 import parallel, workloads as W
    from dcol_trajectory_optimization_b200.shapes import flatten_shapes
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes, i1, i2, p1, p2 = W.config4_batch(n_pairs, seed=3)
    rec, A, b = flatten_shapes(shapes)

    def solve_local(lo, hi, out):
        r = O.solve_batch(rec, A, b, i1[lo:hi], i2[lo:hi], p1[lo:hi], p2[lo:hi], grad_mode=O.GRAD_EXACT, threads=2)
        out.alpha.copy_(torch.from_numpy(r["alpha"]))
        out.grad.copy_(torch.from_numpy(r["grad"]))
        out.iters.copy_(torch.from_numpy(r["iters"]))
        out.status.copy_(torch.from_numpy(r["status"]))

    parts = parallel.sharded_solve(solve_local, n_pairs, rank, world, torch.device("cpu"))
    alpha = torch.cat([p.alpha for _, _, p in parts]).numpy()
    grad = torch.cat([p.grad for _, _, p in parts]).numpy()
    iters = torch.cat([p.iters for _, _, p in parts]).numpy()
    status = torch.cat([p.status for _, _, p in parts]).numpy()
    full = O.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=O.GRAD_EXACT, threads=2)
    ok = (np.array_equal(alpha, full["alpha"]) and np.array_equal(grad, full["grad"])
          and np.array_equal(iters, full["iters"]) and np.array_equal(status, full["status"])
          and [(lo, hi) for lo, hi, _ in parts] == [parallel.shard_bounds(n_pairs, r, world) for r in range(world)])
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_solve_world2_gloo():
    import torch.multiprocessing as mp
    world, n_pairs = 2, 1001        # odd: ranks get 501 / 500 pairs, padded to one fixed-size gather
    port = 29500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_pairs, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_shard_bounds_and_packed_views():
    from dcol_trajectory_optimization_b200 import parallel
    for n, w in [(0, 4), (7, 8), (1001, 2), (1 << 20, 8), (13, 3)]:
        b = [parallel.shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
    flat, v = parallel.alloc_packed(5, "cpu")
    assert flat.numel() == 5 * parallel.WORDS_PER_PAIR
    assert parallel.WORDS_PER_PAIR * 8 == 112 and v.contact is None
    v.alpha.fill_(1.0); v.grad.fill_(2.0); v.iters.fill_(4); v.status.fill_(5)
    w = parallel.packed_views(flat.clone(), 5)
    assert float(w.alpha.sum()) == 5 and float(w.grad.sum()) == 120
    assert w.iters.tolist() == [4] * 5 and w.status.tolist() == [5] * 5
    with pytest.raises(ValueError):
        parallel.packed_views(flat[:-1], 5)

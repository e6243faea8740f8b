"""Run under torchrun (one process per GPU): checks parallel.PeerRecordGather, the all-gather fused into
the solve kernel's epilogue over CUDA-IPC peer mappings.  Used by tests/test_gpu_parity.py."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dcol_trajectory_optimization_b200 as d  # noqa: E402
from dcol_trajectory_optimization_b200 import parallel, workloads as W  # noqa: E402
from dcol_trajectory_optimization_b200.engine import records_to_result  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B = 50_001
    # every rank can rebuild every rank's inputs (seed = 100 + rank), so it can check all gathered slots
    batches = [W.config4_batch(B, seed=100 + r) for r in range(world)]
    shapes, i1, i2, p1, p2 = batches[rank]
    eng = d.ProximityEngine(shapes, device=local)
    plan = eng.plan(i1, i2)
    d1, d2 = torch.from_numpy(p1).to(dev), torch.from_numpy(p2).to(dev)
    fabric = sys.argv[1] if len(sys.argv) > 1 else "unicast"
    if fabric == "multicast":
        try:
            pg = parallel.MulticastRecordGather(B, rank, world, local)
        except Exception as exc:
            print("MULTICAST_UNAVAILABLE", rank, repr(exc)[:200], flush=True)
            dist.barrier()
            dist.destroy_process_group()
            return
    else:
        pg = parallel.PeerRecordGather(B, rank, world, local)
    pg._both.fill_(-1.0)
    dist.barrier()
    torch.cuda.synchronize()
    eng.solve_records(plan, d1, d2, pg.begin_step(), multicast=pg.multicast)
    pg.handshake()
    torch.cuda.synchronize()
    # every rank's perm differs: exchange them once (static per plan)
    perms = [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(world)]
    dist.all_gather(perms, plan.perm())
    for r in range(world):
        _, j1, j2, q1, q2 = batches[r]
        pl = eng.plan(j1, j2)
        want = eng.solve(pl, torch.from_numpy(q1).to(dev), torch.from_numpy(q2).to(dev))
        got = records_to_result(pg.gathered[r], perms[r])
        torch.cuda.synchronize()
        assert torch.equal(got.iters, want.iters) and torch.equal(got.status, want.status), f"rank {rank} slot {r}"
        assert torch.equal(got.alpha, want.alpha) and torch.equal(got.grad, want.grad), f"rank {rank} slot {r}"
    # the user-facing form: same global batch on every rank, results of all pairs everywhere, in pair order
    from dcol_trajectory_optimization_b200 import workloads as W2
    shapes_g, g1, g2, q1, q2 = W2.config4_batch(20_003, seed=55)           # odd size: shards of 10,002 / 10,001
    eng_g = d.ProximityEngine(shapes_g, device=local)
    fs = parallel.FusedShardedSolver(eng_g, g1, g2, rank, world, fabric=fabric)
    got = fs.solve(torch.from_numpy(q1).to(dev), torch.from_numpy(q2).to(dev))
    pl = eng_g.plan(g1, g2)
    want = eng_g.solve(pl, torch.from_numpy(q1).to(dev), torch.from_numpy(q2).to(dev))
    torch.cuda.synchronize()
    assert torch.equal(got.iters, want.iters) and torch.equal(got.status, want.status)
    assert torch.equal(got.alpha, want.alpha) and torch.equal(got.grad, want.grad)
    # back-to-back solves with drifting poses, nothing synchronised in between, and one rank's CONSUMER delayed after
    # every handshake (its reads of step n are still pending while the other rank is already storing step n+1): with a
    # single gathered buffer this returns records of the wrong step (write-after-read); the double buffer must not
    n_steps = 60
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)                       # same drift on every rank
    base1, base2 = torch.from_numpy(q1).to(dev), torch.from_numpy(q2).to(dev)
    drift = 0.01 * torch.randn(base1.shape, generator=gen, device=dev, dtype=base1.dtype)
    if world > 1:
        plain_handshake = fs.gather.handshake

        def skewed_handshake():
            plain_handshake()
            if rank == world - 1:
                torch.cuda._sleep(3_000_000)    # ~1.5 ms at 1.9 GHz, longer than a whole solve of this size
        fs.gather.handshake = skewed_handshake
    outs = [fs.solve(base1 + t * drift, base2) for t in range(n_steps)]
    torch.cuda.synchronize()
    for t, got in enumerate(outs):
        want = eng_g.solve(pl, base1 + t * drift, base2)
        torch.cuda.synchronize()
        assert torch.equal(got.iters, want.iters) and torch.equal(got.status, want.status), f"rank {rank} step {t}"
        assert torch.equal(got.alpha, want.alpha) and torch.equal(got.grad, want.grad), f"rank {rank} step {t}"
    print("BACK_TO_BACK_OK", rank, n_steps, flush=True)
    fs.close()
    print("PEER_GATHER_OK", rank, flush=True)
    dist.barrier()
    pg.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

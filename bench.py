#!/usr/bin/env python
"""Benchmark of the batched DCOL proximity solve + gradient (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl reference]

One step = one pass of the hot path (alpha + contact point + d alpha / d pose for every pair) over
one batch of the config-4 workload ("synthetic batched proximity sweep": all 40 reference-supported
ordered type pairs over the 7-shape table, random poses), P pairs per GPU (weak scaling), followed
for N > 1 by the single all-gather of every rank's records.  Rank 0 prints ONE JSON line.

 value          pairs/s, whole job, inputs resident in HBM, timed with CUDA events, max over ranks
 e2e            the same metric through the host entry point of the C ABI (dcol_proximity_batch_host):
                page-locked NumPy buffers in and out, host<->device copies inside the timed region
 roofline       FP64: algorithmic flops (actual iteration counts) of the
                pair_kernel launches / their CUDA-event duration / the FP64 FMA peak measured in
                this run (dcol_measure_fp64_peak; MEASURED_PEAKS.json has no FP64 entry)
 cpu_baseline   the oracle (C port of the reference's NumPy path, all host threads) on a bounded
                sample of the same workload; rank 0, N = 1 only
 --impl reference   times that CPU path alone, same metric / config
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "proximity solves+grads/sec"
UNIT = "pairs/s"
WORKLOAD = ("config4: synthetic batched proximity sweep, 40 ordered type pairs x random poses, alpha + d alpha/d pose [12] "
            "(what proximity_gradient returns) + iters + status per pair")


WORKLOAD5 = "config5: scaled quadrotor hallway, sphere victim vs 1024 obstacles (11 shapes) x 100 knots x candidates, alpha + d alpha/d pose [12] + iters + status per pair"


def make_batch(n_pairs, seed, workload="config4"):
    from dcol_trajectory_optimization_b200 import workloads as W
    from dcol_trajectory_optimization_b200.shapes import flatten_shapes
    if workload == "config5":       # whole trajectories: n_pairs is rounded down to a multiple of 1024 x 100
        n_cand = max(1, n_pairs // (1024 * 100))
        shapes, i1, i2, p1, p2 = W.config5_batch(n_obs=1024, n_knots=100, n_cand=n_cand, seed=seed)
    else:
        shapes, i1, i2, p1, p2 = W.config4_batch(n_pairs, seed=seed)
    return flatten_shapes(shapes), i1, i2, p1, p2


def flop_model_total(rec, i1, i2, iters):
    """Sum over pairs of the algorithmic flop model with each pair's actual iteration count."""
    from dcol_trajectory_optimization_b200.shapes import POLYTOPE, flop_model, problem_dims
    ns = len(rec)
    key = i1.astype(np.int64) * ns + i2
    cnt = np.bincount(key, minlength=ns * ns)
    its = np.bincount(key, weights=iters.astype(np.float64), minlength=ns * ns)
    total = 0.0
    for k in np.nonzero(cnt)[0]:
        r1, r2 = rec[k // ns], rec[k % ns]
        m_ort, q1, q2, n = problem_dims(r1, r2)
        args = (m_ort, q1, q2, n, int(r1["n_faces"]), int(r2["n_faces"]), int(r1["type"]) == POLYTOPE,
                int(r2["type"]) == POLYTOPE)
        f0 = flop_model(*args, 0)
        fit = flop_model(*args, 1) - f0
        total += cnt[k] * f0 + its[k] * fit
    return total


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.1 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


def cpu_baseline(n_target_seconds=12.0, threads=None, seed=1234):
    """The oracle (C port of the reference path, FD gradient as in the reference) on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    threads = threads or os.cpu_count() or 1
    (rec, A, b), i1, i2, p1, p2 = make_batch(1 << 16, seed)
    t = time.perf_counter()
    O.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=O.GRAD_FD, threads=threads)
    rate = (1 << 16) / (time.perf_counter() - t)
    n = int(min(max(rate * n_target_seconds, 1 << 16), 1 << 24))
    (rec, A, b), i1, i2, p1, p2 = make_batch(n, seed)
    t = time.perf_counter()
    r = O.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=O.GRAD_FD, threads=threads)
    dt = time.perf_counter() - t
    assert int((r["status"] != 0).sum()) == 0
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {n} pairs of the config-4 batch (seed {seed}), oracle/dcol_oracle.c with the reference's "
                      f"13-evaluation finite-difference gradient, {threads} POSIX threads, {dt:.1f} s; the Python reference "
                      f"itself runs ~1e2 calls/s/core (BASELINE.md section 2), this C port ~6e4"}


def cpu_baseline_python(single=1000, per_worker=200, timeout=600):
    """The UNMODIFIED Python reference (baseline/_ref, installed by baseline/install_ref.py) timed on this box's host
    cores in a subprocess: proximity_gradient on the head of the config-4 batch, one process and a fork pool of
    os.cpu_count() workers.  Returns (dict for the JSON line, path of the saved outputs)."""
    import tempfile
    script = os.path.join(ROOT, "baseline", "time_reference.py")
    check = os.path.join(tempfile.gettempdir(), f"dcol_ref_check_{os.getpid()}.npz")
    try:
        r = subprocess.run([sys.executable, "-W", "ignore", script, "--single", str(single), "--per-worker", str(per_worker),
                            "--check", check], capture_output=True, text=True, timeout=timeout)
        out = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as exc:
        return {"unavailable": f"baseline/time_reference.py failed: {exc!r}"[:300]}, None
    return out, (check if os.path.exists(check) else None)


def parity_vs_python_reference(check_path, device=0):
    """alpha / gradient of the CUDA path against what the unmodified Python reference just computed on this box for
    the same pairs (the head of the exact-sequence config-4 batch)."""
    import dcol_trajectory_optimization_b200 as d
    from dcol_trajectory_optimization_b200 import workloads as W
    from dcol_trajectory_optimization_b200.shapes import flatten_shapes
    ref = np.load(check_path)
    n = len(ref["alpha"])
    shapes, i1, i2, p1, p2 = W.config4_batch(n, seed=1234, exact=True)
    eng = d.ProximityEngine(flatten_shapes(shapes), device=device)
    res = eng.solve_host(i1, i2, p1, p2)
    eng.close()
    a_err = np.abs(res.alpha - ref["alpha"]) / np.maximum(np.abs(ref["alpha"]), 1.0)
    g_err = np.abs(res.grad - ref["grad"]).max(axis=1) / np.abs(ref["grad"]).max(axis=1)
    return {"pairs": n, "failed_pairs": int((res.status != 0).sum()), "max_alpha_rel_err": float(a_err.max()),
            "max_grad_rel_err": float(g_err.max()), "bars": {"alpha": 1e-8, "grad": 1e-6},
            "note": "reference gradient is a forward finite difference (h = 2^-26): its own noise is up to 5.4e-7"}


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of the CURRENT kernels
    (profiles/r02_ncu_traffic.json, written by tools/ncu_summary.py --json from `ncu --set full` of this bench command);
    null when no capture of the current kernels is committed."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    try:
        t = json.load(open(path))
        return {"traffic": t["mean_dram_bytes_per_launch"],
                "traffic_detail": {"unit": "B per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                                   "pairs_per_launch": t["pairs_per_launch"],
                                   "algorithmic_bytes_per_launch": t["pairs_per_launch"] * 212,
                                   "launches_captured": t["launches"], "source": "profiles/r02_ncu_traffic.json <- " + t["source"]}}
    except Exception:
        return {"traffic": None, "traffic_detail": {"note": "no ncu capture of the current kernels committed"}}


def parity_sample(eng_shapes, device, n=1 << 18, seed=4321):
    """Parity of the CUDA path with the oracle on a fresh sample of the workload, reported in the line: status /
    iteration-count mismatches, alpha and gradient errors, and how many pairs are ROUNDING-SENSITIVE in the reference's
    own arithmetic (its gradient moves by more than 1e-6 when the oracle is merely compiled with fused multiply-adds;
    those pairs are held to 1e-3 by the tests, every other pair to 1e-6)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import dcol_trajectory_optimization_b200 as d
    (rec, A, b), i1, i2, p1, p2 = make_batch(n, seed)
    ref = O.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=O.GRAD_EXACT)
    fma = O.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=O.GRAD_EXACT, fma=True)
    eng = d.ProximityEngine((rec, A, b), device=device)
    res = eng.solve_host(i1, i2, p1, p2)
    eng.close()
    gscale = np.abs(ref["grad"]).max(axis=1)
    sens = (np.abs(fma["grad"] - ref["grad"]).max(axis=1) / gscale) > 1e-6
    gerr = np.abs(res.grad - ref["grad"]).max(axis=1) / gscale
    aerr = np.abs(res.alpha - ref["alpha"]) / np.maximum(np.abs(ref["alpha"]), 1.0)
    return {"pairs": n, "status_mismatches": int((res.status != ref["status"]).sum()),
            "iteration_count_mismatches": int((res.iters != ref["iters"]).sum()),
            "oracle_fma_build_iteration_count_mismatches": int((fma["iters"] != ref["iters"]).sum()),
            "max_alpha_rel_err": float(aerr.max()), "pairs_with_grad_err_above_1e-6": int((gerr > 1e-6).sum()),
            "rounding_sensitive_pairs": int(sens.sum()),
            "max_grad_rel_err_excluding_rounding_sensitive": float(gerr[~sens].max()),
            "max_grad_rel_err": float(gerr.max()),
            "note": "oracle = C restatement of the reference pinned to reference-generated goldens; gradient vs the exact "
                    "derivative of the reference's frozen-(x, z) Lagrangian"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference
    itself is pure Python and cannot be compiled into oracle/_ref), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    threads = os.cpu_count() or 1
    n = args.ref_pairs
    (rec, A, b), i1, i2, p1, p2 = make_batch(n, 1234)
    for _ in range(args.warmup):
        O.solve_batch(rec, A, b, i1[:n // 8], i2[:n // 8], p1[:n // 8], p2[:n // 8], grad_mode=O.GRAD_FD, threads=threads)
    t = time.perf_counter()
    for _ in range(args.steps):
        O.solve_batch(rec, A, b, i1, i2, p1, p2, grad_mode=O.GRAD_FD, threads=threads)
    dt = (time.perf_counter() - t) / args.steps
    v = n / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_step": n, "note": "bounded sample of the same workload"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n} pairs per step, oracle/dcol_oracle.c (C port of the NumPy reference, FD "
                                       f"gradient), {threads} threads"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_python_reference:
        # the reference itself (pure Python, ~1e2 calls/s/core): reported next to the port, which stays the arm's
        # `value` because it is the STRONGER baseline (same algorithm in C on all cores)
        line["cpu_baseline_python"], _ = cpu_baseline_python()
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=1 << 23, help="pairs per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"], help="N > 1: how records reach all ranks")
    ap.add_argument("--fabric", default="auto", choices=["auto", "multicast", "unicast"], help="--gather fused: NVLS multicast or peer stores")
    ap.add_argument("--chunks", type=int, default=1, help="--gather nccl: chunks per step (gather/solve overlap)")
    ap.add_argument("--ref-pairs", type=int, default=1 << 21, help="pairs per step of the reference arm")
    ap.add_argument("--workload", default="config4", choices=["config4", "config5"])
    ap.add_argument("--no-altro", action="store_true", help="skip the three ALTRO scenario solves (N = 1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-python-reference", action="store_true", help="skip timing the unmodified Python reference (baseline/_ref)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sizes", action="store_true", help="skip the config5_full / config4_2p26 extra measurements")
    ap.add_argument("--no-parity-sample", action="store_true", help="skip the CUDA-vs-oracle parity sample (N = 1 only)")
    ap.add_argument("--no-coherent", action="store_true", help="skip the coherent re-solve line (N = 1 only)")
    ap.add_argument("--no-jacobian", action="store_true", help="skip the solution-Jacobian throughput line (N = 1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import dcol_trajectory_optimization_b200 as d
    from dcol_trajectory_optimization_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    B = args.pairs

    (rec, A, b), i1, i2, p1, p2 = make_batch(B, 1234 + rank, args.workload)
    B = len(i1)
    eng = d.ProximityEngine((rec, A, b), device=local_rank)
    fp64_peak = d.measure_fp64_peak(local_rank)
    d1, d2 = torch.from_numpy(p1).to(dev), torch.from_numpy(p2).to(dev)
    # N > 1, default: the all-gather is FUSED into the solve — the kernel's epilogue stores every 112-byte record
    # into slot `rank` of every rank's gathered buffer (local + NVLink peer mappings); a 4-byte all-reduce on
    # the same stream is the completion handshake.  --gather nccl: solve, then all_gather_into_tensor.
    mode = "none" if world == 1 else args.gather
    peer = None
    def agree(ok):      # every rank must take the same path
        flag = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return float(flag) == 1.0

    if mode == "fused":
        # 1st choice: NVLink multicast (one multimem.st per 16 B reaches every rank); 2nd: unicast peer stores through
        # CUDA-IPC mappings; last: solve, then NCCL all-gather
        if args.fabric in ("auto", "multicast"):
            try:
                peer = parallel.MulticastRecordGather(B, rank, world, local_rank)
            except Exception as exc:
                peer = None
                if rank == 0:
                    print(f"# multicast gather unavailable ({exc})", file=sys.stderr, flush=True)
            if not agree(peer is not None):
                peer = None
        if peer is None:
            try:
                peer = parallel.PeerRecordGather(B, rank, world, local_rank)
            except Exception as exc:            # no CUDA IPC between the ranks on this box
                peer = None
                if rank == 0:
                    print(f"# peer-store gather unavailable ({exc}); using NCCL all-gather", file=sys.stderr, flush=True)
            if not agree(peer is not None):
                peer, mode = None, "nccl"
    n_chunks = args.chunks if mode == "nccl" else 1
    bounds = [parallel.shard_bounds(B, c, n_chunks) for c in range(n_chunks)]
    plans = [eng.plan(i1[lo:hi], i2[lo:hi]) for lo, hi in bounds]
    n_launches = sum(p.n_launches for p in plans)
    pipe = None
    if mode == "fused":
        plan_perm = plans[0].perm()

        def step():
            eng.solve_records(plans[0], d1, d2, peer.begin_step(), multicast=peer.multicast)   # double-buffered slots
            peer.handshake()
    else:
        pipe = parallel.GatherPipeline(bounds, world if mode == "nccl" else 1, dev, with_contact=False)

        def launch(c, out):
            lo, hi = bounds[c]
            eng.solve(plans[c], d1[lo:hi], d2[lo:hi], out=out)

        def step():
            pipe.step(launch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t0 = time.time()
    ev[0].record()
    for s in range(args.steps):
        kev[s][0].record()
        step()
        kev[s][1].record()
    ev[1].record()
    barrier()
    t1 = time.time()
    ms = torch.tensor([ev[0].elapsed_time(ev[1])], dtype=torch.float64, device=dev)
    kms = torch.tensor([sum(a.elapsed_time(bb) for a, bb in kev) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    value = B * world / (ms_per_step * 1e-3)

    gather_check = None
    if mode == "fused":
        # Is rank s's slot of MY gathered buffer exactly what rank s computed?  Every rank solves its own batch once more
        # through the NON-fused array-mode call (separate output arrays, no peer stores), packs those results into records
        # in plan order, and publishes (a) two 64-bit checksums of all B records and (b) its first 65,536 records, through
        # NCCL.  Every rank then compares EVERY slot of its gathered buffer with the owner's checksums (all records) and
        # sample (bit for bit); the verdict is the AND over ranks.
        from dcol_trajectory_optimization_b200.engine import records_to_result

        def checksums(rec):
            x = rec.contiguous().view(torch.int64).reshape(-1)
            wgt = torch.arange(x.numel(), device=x.device, dtype=torch.int64) % 1000003 + 1
            return torch.stack([x.sum(), (x * wgt).sum()])          # int64 arithmetic wraps

        ref = eng.solve(plans[0], d1, d2, want_contact=False)
        pl = plan_perm.long()
        rec_ref = torch.empty((B, parallel.WORDS_PER_PAIR), dtype=torch.float64, device=dev)
        rec_ref[:, 0], rec_ref[:, 1:13] = ref.alpha[pl], ref.grad[pl]
        rec_ref[:, 13] = (ref.iters[pl].long() | (ref.status[pl].long() << 32)).view(torch.float64)
        n_s = min(B, 1 << 16)
        cs_all = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
        sm_all = [torch.empty((n_s, parallel.WORDS_PER_PAIR), dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(cs_all, checksums(rec_ref))
        dist.all_gather(sm_all, rec_ref[:n_s].contiguous())
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        for r in range(world):              # slot r of my gathered buffer was filled by rank r's kernel
            slot = peer.gathered[r]
            good = torch.equal(checksums(slot), cs_all[r]) and torch.equal(slot[:n_s].contiguous().view(torch.int64),
                                                                           sm_all[r].view(torch.int64))
            if not good:
                ok.zero_()
                print(f"# rank {rank}: slot {r} of the gathered buffer differs from rank {r}'s own results", file=sys.stderr, flush=True)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        gather_check = {"gather_verified": bool(int(ok)), "slots_checked": world,
                        "how": f"every rank compared every slot of its gathered buffer with the owner's non-fused array-mode "
                               f"re-solve: two 64-bit checksums over all {B} records per slot + the first {n_s} records bit for bit"}
        mine = [records_to_result(peer.gathered[rank], plan_perm)]
        del ref, rec_ref, sm_all
    else:
        mine = pipe.rank_results(rank)      # nccl: this rank's records out of the gathered buffers
    iters = torch.cat([r.iters for r in mine]).cpu().numpy()
    n_fail = int(sum(int((r.status != 0).sum()) for r in mine))
    flops = flop_model_total(rec, i1, i2, iters)
    kernel_ms = float(kms)
    achieved = flops / (kernel_ms * 1e-3)
    bytes_alg = B * (96 + 4 + 112)   # poses + perm in, alpha + grad + iters + status out
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)

    # ---- end to end through the host entry point of the C ABI (pinned host buffers) ----
    e2e = None
    if not args.no_e2e:
        Be = B
        hi1, hi2 = d.pinned_empty(Be, np.int32), d.pinned_empty(Be, np.int32)
        hp1, hp2 = d.pinned_empty((Be, 6)), d.pinned_empty((Be, 6))
        hi1[:], hi2[:], hp1[:], hp2[:] = i1, i2, p1, p2
        hout = d.BatchResult(alpha=d.pinned_empty(Be), contact=None, grad=d.pinned_empty((Be, 12)),
                             iters=d.pinned_empty(Be, np.int32), status=d.pinned_empty(Be, np.int32))
        for _ in range(2):
            eng.solve_host(hi1, hi2, hp1, hp2, out=hout)
        barrier()
        t = time.perf_counter()
        n_e2e = max(10, args.steps)      # ~22 ms each: the mean over a quarter of a second is stable against host hiccups
        for _ in range(n_e2e):
            eng.solve_host(hi1, hi2, hp1, hp2, out=hout)
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t) / n_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert np.array_equal(hout.iters, iters)
        e2e = {"value": Be * world / float(dt), "unit": UNIT, "h2d_bytes_per_step": Be * (48 + 48 + 4 + 4),
               "d2h_bytes_per_step": Be * (8 + 96 + 4 + 4), "ms_per_step": float(dt) * 1e3,
               "api": "dcol_proximity_batch_host (ProximityEngine.solve_host), page-locked NumPy buffers, "
                      "includes the per-chunk plan (counting sort)"}

    # ---- end to end on the scene workload (config 5) through the scene entry point: only poses cross PCIe on the way in ----
    e2e_scene = None
    if not args.no_e2e and world == 1:
        from dcol_trajectory_optimization_b200 import workloads as Wl
        from dcol_trajectory_optimization_b200.shapes import flatten_shapes as _flat
        n_obs, n_knots = 1024, 100
        n_cand = max(1, min(B, 1 << 23) // (n_obs * n_knots))
        rng5 = np.random.default_rng(2)
        shapes5 = [Wl.SphereMRP(0.25)] + Wl.quadrotor_obstacle_shapes()
        obs_pose = np.concatenate([rng5.uniform([-8.0, -2.5, 1.0], [8.0, 2.5, 6.0], size=(n_obs, 3)),
                                   rng5.normal(size=(n_obs, 3)) * 0.5], axis=1)
        obs_shape = (1 + (np.arange(n_obs) % 11)).astype(np.int32)
        knots = np.linspace([-8.0, 0.0, 4.0], [8.0, 0.0, 4.0], n_knots)
        Ms = n_cand * n_knots
        hv = d.pinned_empty((Ms, 6))
        hv[:] = 0.0
        hv[:, :3] = (knots[None] + rng5.normal(size=(n_cand, n_knots, 3)) * 0.3).reshape(Ms, 3)
        eng5 = d.ProximityEngine(_flat(shapes5), device=local_rank)
        sout = d.SceneResult(alpha=d.pinned_empty((Ms, n_obs)), grad1=d.pinned_empty((Ms, n_obs, 6)),
                             iters=d.pinned_empty((Ms, n_obs), np.int32), status=d.pinned_empty((Ms, n_obs), np.int32))
        for _ in range(2):
            eng5.solve_scene_host(0, hv, obs_shape, obs_pose, out=sout)
        t = time.perf_counter()
        n_rep = max(10, args.steps)
        for _ in range(n_rep):
            eng5.solve_scene_host(0, hv, obs_shape, obs_pose, out=sout)
        dt5 = (time.perf_counter() - t) / n_rep
        # the same pairs with inputs resident on the device (plan reused), for the e2e / device ratio
        i1s, i2s = np.zeros(Ms * n_obs, np.int32), np.tile(obs_shape, Ms)
        plan5 = eng5.plan(i1s, i2s)
        q1 = torch.from_numpy(np.repeat(np.asarray(hv), n_obs, axis=0)).to(dev)
        q2 = torch.from_numpy(np.tile(obs_pose, (Ms, 1))).to(dev)
        o5 = None
        ev5 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for s_ in range(2 + n_rep):
            if s_ == 2:
                ev5[0].record()
            o5 = eng5.solve(plan5, q1, q2, want_contact=False, grad1=True, out=o5)
        ev5[1].record()
        torch.cuda.synchronize()
        dev5 = Ms * n_obs / (ev5[0].elapsed_time(ev5[1]) * 1e-3 / n_rep)
        assert np.array_equal(sout.iters.ravel(), o5.iters.cpu().numpy())
        e2e_scene = {"value": Ms * n_obs / dt5, "unit": UNIT, "pairs_per_step": Ms * n_obs, "ms_per_step": dt5 * 1e3,
                     "h2d_bytes_per_step": int(Ms * 48 + n_obs * 52), "d2h_bytes_per_step": int(Ms * n_obs * (8 + 48 + 4 + 4)),
                     "device_resident_value": dev5, "e2e_over_device": (Ms * n_obs / dt5) / dev5,
                     "failed_pairs": int((sout.status != 0).sum()),
                     "workload": f"config5: sphere victim, {n_cand} candidates x {n_knots} knots x {n_obs} obstacles",
                     "api": "dcol_proximity_scene_host (ProximityEngine.solve_scene_host): victim poses [M][6] + obstacle poses "
                            "[n_obs][6] in, alpha + d alpha/d(victim pose) [6] + iters + status out (64 B/pair), page-locked buffers"}
        plan5.close()
        eng5.close()
        del q1, q2, o5

    # ---- BASELINE.json configs at their stated sizes, as extra keys (not `value`) ----
    extra_sizes = {}
    if not args.no_sizes:
        del_keep = (d1, d2)      # the main batch stays resident; these runs allocate their own inputs on the device

        def timed(engine, plan, q1, q2, reps=3):
            out = None
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for s_ in range(2 + reps):
                if s_ == 2:
                    barrier()
                    evs[0].record()
                out = engine.solve(plan, q1, q2, want_contact=False, out=out)
            evs[1].record()
            barrier()
            t_ms = torch.tensor([evs[0].elapsed_time(evs[1]) / reps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
            fails = (out.status != 0).sum().to(torch.float64).reshape(1)
            if world > 1:
                dist.all_reduce(fails)
            return float(t_ms), int(fails), float(out.iters.double().mean())

        # config 5 at its full size: 1024 obstacles x 100 knots x 256 candidates = 26,214,400 pairs, candidates sharded
        from dcol_trajectory_optimization_b200 import workloads as Wl
        from dcol_trajectory_optimization_b200.shapes import flatten_shapes as _flat
        n_obs, n_knots, n_cand = 1024, 100, 256
        clo, chi = parallel.shard_bounds(n_cand, rank, world)
        rng5 = np.random.default_rng(2)
        obs = torch.from_numpy(np.concatenate([rng5.uniform([-8.0, -2.5, 1.0], [8.0, 2.5, 6.0], size=(n_obs, 3)),
                                               rng5.normal(size=(n_obs, 3)) * 0.5], axis=1)).to(dev)
        knots = np.linspace([-8.0, 0.0, 4.0], [8.0, 0.0, 4.0], n_knots)
        vic = np.zeros((n_cand, n_knots, 6))
        vic[..., :3] = knots[None] + rng5.normal(size=(n_cand, n_knots, 3)) * 0.3
        vic = torch.from_numpy(vic[clo:chi].reshape(-1, 6)).to(dev)
        Ms = vic.shape[0]
        q1 = vic.repeat_interleave(n_obs, dim=0).contiguous()
        q2 = obs.repeat(Ms, 1).contiguous()
        eng5 = d.ProximityEngine(_flat([Wl.SphereMRP(0.25)] + Wl.quadrotor_obstacle_shapes()), device=local_rank)
        plan5 = eng5.plan(torch.zeros(Ms * n_obs, dtype=torch.int32), (1 + torch.arange(n_obs, dtype=torch.int32) % 11).repeat(Ms))
        t5, f5, it5 = timed(eng5, plan5, q1, q2)
        extra_sizes["config5_full"] = {"pairs": n_obs * n_knots * n_cand, "n_gpus": world, "ms_per_step": t5,
                                       "value": n_obs * n_knots * n_cand / (t5 * 1e-3), "unit": UNIT, "failed_pairs": f5,
                                       "mean_pdip_iters": it5, "scaling": "strong (fixed 26,214,400 pairs, candidates sharded)",
                                       "outputs": "alpha + grad[12] + iters + status, device resident, no gather"}
        plan5.close()
        eng5.close()
        del q1, q2, vic, obs
        # config 4 at 2^26 pairs per GPU (the top of BASELINE.json's 1M-64M sweep): poses drawn on the device from the same
        # distribution (torch generator), 40 type pairs round-robin
        n26 = 1 << 26
        gen = torch.Generator(device=dev)
        gen.manual_seed(99 + rank)

        def ball(radius):
            u = torch.randn((n26, 3), generator=gen, device=dev, dtype=torch.float64)
            u /= u.norm(dim=1, keepdim=True)
            return u * (radius * torch.rand((n26, 1), generator=gen, device=dev, dtype=torch.float64))
        q1 = torch.cat([ball(1.0), 0.5 * torch.randn((n26, 3), generator=gen, device=dev, dtype=torch.float64)], dim=1)
        q2 = torch.cat([ball(6.0), 0.5 * torch.randn((n26, 3), generator=gen, device=dev, dtype=torch.float64)], dim=1)
        pairs40 = torch.from_numpy(np.asarray(Wl.supported_type_pairs(Wl.config4_shapes()), dtype=np.int32))
        sel = torch.arange(n26) % len(pairs40)
        eng4 = d.ProximityEngine(_flat(Wl.config4_shapes()), device=local_rank)
        plan26 = eng4.plan(pairs40[sel, 0].contiguous(), pairs40[sel, 1].contiguous())
        t26, f26, it26 = timed(eng4, plan26, q1, q2, reps=2)
        extra_sizes["config4_2p26"] = {"pairs_per_gpu": n26, "n_gpus": world, "ms_per_step": t26,
                                       "value": n26 * world / (t26 * 1e-3), "unit": UNIT, "failed_pairs": f26,
                                       "mean_pdip_iters": it26, "scaling": "weak"}
        plan26.close()
        eng4.close()
        del q1, q2, sel

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.workload == "config4" else WORKLOAD5, "pairs_per_gpu": B, "pairs_per_step": B * world,
                       "reference_arm_pairs_per_step": args.ref_pairs,
                       "same_config_note": "the --impl reference arm runs a bounded sample of the same workload per step (the metric is a rate)",
                       "type_pairs": plans[0].n_groups,
                       "mean_pdip_iters": float(iters.mean()), "failed_pairs": n_fail,
                       "l2": "inputs larger than L2 (805 MB of poses per step at the default size)",
                       "plan": "pairs grouped by shape pair once (device counting sort), plan reused by every step, as ALTRO "
                               "re-evaluates a fixed pair list; the e2e figure re-plans every chunk",
                       "collective": {"none": "none",
                                      "fused": "all-gather fused into the solve: the kernel epilogue stores each 112 B record "
                                               "(alpha, grad[12], iters, status) to every rank's buffer over NVLink ("
                                               + ("one multimem.st per 16 B through the NVSwitch multicast address"
                                                  if (peer is not None and peer.multicast) else "unicast stores to CUDA-IPC peer mappings")
                                               + "); 4-byte NCCL all-reduce as completion handshake",
                                      "nccl": f"all_gather_into_tensor of the 112 B/pair records, {n_chunks} chunk(s) per step"}[mode],
                       "parallelism": f"batch sharded over {world} GPU(s), one process per GPU"},
            "clocks": clocks,
            **(gather_check or {}),
            "e2e": e2e,
            "e2e_scene": e2e_scene,
            **extra_sizes,
            "gpu_launches": args.steps * n_launches,
            "roofline": {"bound": "fp64", "achieved": achieved / 1e12, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak,
                         **measured_traffic(),
                         "kernel": f"dcol::pair_kernel<P1,P2> ({plans[0].n_groups} specialisations, {n_launches} launches per step)",
                         "kernel_ms_per_step": kernel_ms, "model_flops_per_pair": flops / B,
                         "peak_source": "dcol_measure_fp64_peak in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "hbm": {"achieved": bytes_alg / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": bytes_alg / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                                 "bytes_per_pair": 212}},
        }
        if world == 1 and not args.no_altro:
            # BASELINE.json metric, second half: "ALTRO solve time" of the reference's three scenarios through the
            # batched caller (altro/solver.py), every collision constraint on this GPU
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from run_altro import run as run_altro
            line["altro"] = run_altro()
        if world == 1 and not args.no_jacobian:
            # extension: the same solve plus the solution Jacobian d(contact, alpha)/d pose
            # [4][12], from the separate jacobian kernels (dcol_proximity_batch_jacobian); not part of `value`
            nj = min(B, 1 << 21)
            jplan = eng.plan(i1[:nj], i2[:nj])
            jout = None
            jev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for s_ in range(5):
                if s_ == 2:
                    jev[0].record()
                jout = eng.solve(jplan, d1[:nj], d2[:nj], want_contact=False, want_jac=True, out=jout)
            jev[1].record()
            torch.cuda.synchronize()
            jms = jev[0].elapsed_time(jev[1]) / 3
            line["jacobian"] = {"value": nj / (jms * 1e-3), "unit": UNIT, "pairs_per_step": nj, "ms_per_step": jms,
                                "finite": bool(torch.isfinite(jout.jac).all()),
                                "outputs": "alpha + grad[12] + jac[4][12] + iters + status (496 B/pair)"}
        if world == 1 and not args.no_coherent:
            # (new) temporally coherent re-solves, the ALTRO access pattern: the SAME pair list is solved again and again
            # with slowly drifting poses (random walk, sigma per step), and after every solve the plan is re-ordered by that
            # solve's iteration counts (dcol_plan_refine), so warps hold pairs of nearly equal count.  Every step solves
            # poses it has never seen; the refine kernels are inside the timed region.  Not part of `value`.
            nc, sigma, n_seq, n_warm = min(B, 1 << 21), 0.01, 8, 3
            gen = torch.Generator(device=dev)
            gen.manual_seed(7)
            seq, cur = [], d1[:nc].clone()
            for _ in range(n_seq + n_warm):
                cur = cur + sigma * torch.randn(cur.shape, generator=gen, device=dev, dtype=cur.dtype)
                seq.append(cur)
            rates = {}
            for mode in ("plain", "refined"):
                cplan = eng.plan(i1[:nc], i2[:nc])
                cout = None
                cev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                for s_, poses in enumerate(seq):
                    if s_ == n_warm:
                        cev[0].record()
                    cout = eng.solve(cplan, poses, d2[:nc], want_contact=False, out=cout)
                    if mode == "refined":
                        cplan.refine(cout.iters)
                cev[1].record()
                torch.cuda.synchronize()
                rates[mode] = nc * n_seq / (cev[0].elapsed_time(cev[1]) * 1e-3)
                cplan.close()
            line["coherent_resolve"] = {"value": rates["refined"], "value_without_refine": rates["plain"], "unit": UNIT,
                                        "pairs_per_step": nc, "steps": n_seq, "pose_drift_sigma_per_step": sigma,
                                        "failed_pairs_last_step": int((cout.status != 0).sum()),
                                        "note": "fixed pair list, poses drift by a random walk; plan re-ordered after every "
                                                "solve by that solve's iteration counts (dcol_plan_refine, inside the timed region)"}
        if world == 1 and not args.no_parity_sample:
            line["parity_sample"] = parity_sample(None, local_rank)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
            if not args.no_python_reference:
                line["cpu_baseline_python"], check = cpu_baseline_python()
                if check:
                    line["parity_vs_python_reference"] = parity_vs_python_reference(check, local_rank)
                    os.remove(check)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

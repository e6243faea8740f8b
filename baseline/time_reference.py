#!/usr/bin/env python
"""CPU baseline, the real thing: time the UNMODIFIED Python reference's ``proximity_gradient``
(``baseline/_ref/proximity/proximity_gradient.py:91-138``) on the head of the config-4 batch, on this machine's
host cores: (a) one process, BLAS pinned to one thread; (b) a fork pool with one worker
per core over disjoint contiguous slices.  Prints ONE JSON line.  Run as a subprocess of ``bench.py`` so that the
reference's top-level packages (``proximity``, ``primitives``) never enter the benchmark process.

    python baseline/time_reference.py [--single N] [--per-worker M] [--check out.npz]
"""
import argparse
import json
import os
import sys
import time

for v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[v] = "1"

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def _import_reference():
    import types
    import warnings
    from unittest.mock import MagicMock
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "mpl_toolkits.mplot3d.art3d", "meshcat", "h5py"):
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        from proximity.proximity_gradient import proximity_gradient
        import primitives.misc_primitive_constructor as prims
    return types.SimpleNamespace(proximity_gradient=proximity_gradient, prims=prims)


def _reference_shapes(ref):
    """The 7-shape table of config 4 as objects of the REFERENCE's own classes."""
    d = np.load(os.path.join(ROOT, "dcol_trajectory_optimization_b200", "data", "polytopes.npz"))
    A2, b2 = d["A2"].T.copy(), d["b2"].copy()
    P = ref.prims
    ngon = P.create_n_sided(5, 0.6)
    return [P.create_rect_prism(1.0, 2.0, 3.0), P.PolytopeMRP(A2, b2), P.CapsuleMRP(0.3, 1.2), P.CylinderMRP(0.4, 1.5),
            P.ConeMRP(2.0, np.deg2rad(22)), P.SphereMRP(0.5), P.PolygonMRP(ngon["A"], ngon["b"], 0.2)]


def _batch(n):
    sys.path.insert(0, ROOT)
    from dcol_trajectory_optimization_b200 import workloads as W
    _, i1, i2, p1, p2 = W.config4_batch(n, seed=1234, exact=True)      # the exact call sequence of SURVEY.md 8(d)
    return i1, i2, p1, p2


_G = {}


def _run_slice(bounds):
    lo, hi = bounds
    ref, shapes, (i1, i2, p1, p2) = _G["ref"], _G["shapes"], _G["batch"]
    import copy
    out_a, out_g = np.empty(hi - lo), np.empty((hi - lo, 12))
    t = time.perf_counter()
    for k in range(lo, hi):
        a, b = copy.copy(shapes[i1[k]]), copy.copy(shapes[i2[k]])
        a.r, a.p, b.r, b.p = p1[k, :3].copy(), p1[k, 3:].copy(), p2[k, :3].copy(), p2[k, 3:].copy()
        out_a[k - lo], out_g[k - lo] = ref.proximity_gradient(a, b)
    return time.perf_counter() - t, out_a, out_g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--single", type=int, default=1000, help="pairs of the single-process leg")
    ap.add_argument("--per-worker", type=int, default=200, help="pairs per pool worker")
    ap.add_argument("--check", default=None, help="also save alpha / grad of the single-process leg here (.npz)")
    args = ap.parse_args()
    if not os.path.isdir(REF):
        print(json.dumps({"unavailable": "baseline/_ref is not installed (python baseline/install_ref.py)"}))
        return
    import multiprocessing as mp
    import scipy
    cores = os.cpu_count() or 1
    n_pool = args.per_worker * cores
    _G["ref"] = ref = _import_reference()
    _G["shapes"] = _reference_shapes(ref)
    _G["batch"] = _batch(max(args.single, n_pool))
    _run_slice((0, 20))                                     # warm-up (imports, first-call costs)
    dt, alpha, grad = _run_slice((0, args.single))
    if args.check:
        np.savez(args.check, alpha=alpha, grad=grad)
    single = args.single / dt
    t = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:        # workers inherit the imported reference and the batch
        parts = pool.map(_run_slice, [(w * args.per_worker, (w + 1) * args.per_worker) for w in range(cores)])
    wall = time.perf_counter() - t
    model = ""
    try:
        model = [ln.split(":", 1)[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")][0]
    except Exception:
        pass
    print(json.dumps({
        "kind": "reference", "unit": "pairs/s", "value": n_pool / wall, "cores": cores,
        "single_process": {"value": single, "pairs": args.single, "seconds": dt},
        "pool": {"value": n_pool / wall, "pairs": n_pool, "workers": cores, "wall_seconds": wall,
                 "busy_seconds_max": max(p[0] for p in parts)},
        "cpu_model": model, "numpy": np.__version__, "scipy": scipy.__version__,
        "sample": f"unmodified proximity_gradient of baseline/_ref on the head of the config-4 batch (default_rng(1234), "
                  f"exact call sequence): first {args.single} pairs in one process, first {n_pool} pairs over a fork pool of "
                  f"{cores} workers; BLAS pinned to 1 thread"}))


if __name__ == "__main__":
    main()

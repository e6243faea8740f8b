"""TEST INFRASTRUCTURE: generate the committed golden fixtures under tests/golden/.

Runs the UNMODIFIED Python reference (imported from /root/reference, see _refimport.py) on
seeded inputs and stores inputs + outputs as small .npz files.  The reference has no tests
or golden vectors of its own, so these files — outputs of the
reference itself, run in the build container — are what pins the oracle (oracle/dcol_oracle.c)
and, through it, the CUDA path.

    python oracle/gen_golden.py            # all fixtures (about 3 minutes on 8 cores)
    python oracle/gen_golden.py scenarios  # one of: scenarios config4 config5 edge

Per pair the reference is driven through its public ``proximity_gradient`` (which also
returns alpha); the PDIP iteration count is the number of ``calc_NT_scalings`` calls minus
one, (x, s, z) are captured at the ``solve_lp_pdip`` return, and exceptions are mapped to the
status words of include/dcol.h.
"""
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from dcol_trajectory_optimization_b200 import shapes as S          # noqa: E402
from dcol_trajectory_optimization_b200 import workloads as W       # noqa: E402
from dcol_trajectory_optimization_b200 import primitives as P      # noqa: E402
from _refimport import import_reference                            # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
MAX_M = 48      # padded row count in the fixtures
MAX_N = 8
MAX_TRACE = 52

_ref = None


def ref():
    global _ref
    if _ref is None:
        _ref = import_reference()
    return _ref


def to_ref_prim(prim, pose=None):
    """Rebuild one of our primitive objects as the reference's own class instance."""
    R = ref().primitives
    kind = S.kind_of(prim)
    if kind == S.POLYTOPE:
        out = R.PolytopeMRP(np.array(prim.A, dtype=float), np.array(prim.b, dtype=float))
    elif kind == S.CAPSULE:
        out = R.CapsuleMRP(prim.R, prim.L)
    elif kind == S.CYLINDER:
        out = R.CylinderMRP(prim.R, prim.L)
    elif kind == S.CONE:
        out = R.ConeMRP(prim.H, prim.beta)
    elif kind == S.SPHERE:
        out = R.SphereMRP(prim.R)
    else:
        out = R.PolygonMRP(np.array(prim.A, dtype=float), np.array(prim.b, dtype=float), prim.R)
    out.r_offset = np.array(prim.r_offset, dtype=float)
    out.Q_offset = np.array(prim.Q_offset, dtype=float)
    if pose is None:
        pose = S.pose_of(prim)
    out.r = np.array(pose[:3], dtype=float)
    out.p = np.array(pose[3:], dtype=float)
    return out


def run_reference_pair(prim1, prim2, tol=1e-6):
    """One reference proximity_gradient call -> dict(alpha, x, s, z, iters, status, grad, mu_trace)."""
    r = ref()
    rec = {"nt": 0, "mu": [], "xsz": None}
    orig_nt = r.pdip.calc_NT_scalings
    orig_solve = r.proximity_gradient.solve_lp_pdip

    def nt_counting(s, z, idx_ort, idx_soc1, idx_soc2):
        rec["nt"] += 1
        deg = len(idx_ort) + (len(idx_soc1) > 0) + (len(idx_soc2) > 0)
        rec["mu"].append(float(np.dot(s, z) / deg))
        return orig_nt(s, z, idx_ort, idx_soc1, idx_soc2)

    def solve_recording(*a, **k):
        out = orig_solve(*a, **k)
        rec["xsz"] = tuple(np.array(v, dtype=float) for v in out)
        return out

    r.pdip.calc_NT_scalings = nt_counting
    r.proximity_gradient.solve_lp_pdip = solve_recording
    res = dict(alpha=np.nan, x=np.full(MAX_N, np.nan), s=np.full(MAX_M, np.nan), z=np.full(MAX_M, np.nan),
               iters=-1, status=-1, grad=np.full(12, np.nan), mu=np.full(MAX_TRACE, np.nan), n=0, m=0)
    try:
        with np.errstate(all="ignore"):
            alpha, grad = r.proximity_gradient.proximity_gradient(prim1, prim2, pdip_tol=tol)
        x, s, z = rec["xsz"]
        res.update(alpha=float(alpha), grad=np.array(grad, dtype=float), status=S.STATUS_OK,
                   iters=rec["nt"] - 1, n=len(x), m=len(s))
        res["x"][:len(x)] = x
        res["s"][:len(s)] = s
        res["z"][:len(z)] = z
    except np.linalg.LinAlgError:
        res.update(status=S.STATUS_NOT_PD, iters=max(rec["nt"] - 1, 0))
    except ValueError as e:
        if "infs or NaNs" in str(e):
            res.update(status=S.STATUS_NON_FINITE, iters=max(rec["nt"] - 1, 0))
        else:       # np.vstack dimension mismatch: combine_problem_matrices case 4
            res.update(status=S.STATUS_UNSUPPORTED, iters=0)
    except Exception as e:  # noqa: BLE001 - the reference raises bare Exception on max-iter
        if "Maximum number of iterations" not in str(e):
            raise
        res.update(status=S.STATUS_MAX_ITER, iters=rec["nt"])
    finally:
        r.pdip.calc_NT_scalings = orig_nt
        r.proximity_gradient.solve_lp_pdip = orig_solve
    mu = rec["mu"][:MAX_TRACE]
    res["mu"][:len(mu)] = mu
    return res


# ----------------------------------------------------------------------------------------------
_G = {}


def _init_worker(shape_list):
    _G["shapes"] = shape_list


def _work(task):
    i1, i2, pose1, pose2, tol = task
    p1 = to_ref_prim(_G["shapes"][i1], pose1)
    p2 = to_ref_prim(_G["shapes"][i2], pose2)
    return run_reference_pair(p1, p2, tol)


def run_batch(shape_list, idx1, idx2, pose1, pose2, tol=1e-6, procs=None):
    """Reference outputs for a flattened batch, in parallel over processes."""
    tols = np.broadcast_to(np.asarray(tol, dtype=float), (len(idx1),))
    tasks = [(int(idx1[k]), int(idx2[k]), pose1[k], pose2[k], float(tols[k])) for k in range(len(idx1))]
    with Pool(procs or os.cpu_count(), initializer=_init_worker, initargs=(shape_list,)) as pool:
        out = pool.map(_work, tasks, chunksize=max(1, len(tasks) // (8 * (procs or os.cpu_count()))))
    keys = ("alpha", "x", "s", "z", "iters", "status", "grad", "mu", "n", "m")
    return {k: np.array([o[k] for o in out]) for k in keys}


def save(name, shape_list, idx1, idx2, pose1, pose2, tol, res, keep_sz=None, extra=None):
    recs, A, b = S.flatten_shapes(shape_list)
    d = dict(shape_records=recs, A=A, b=b, idx1=np.asarray(idx1, np.int32), idx2=np.asarray(idx2, np.int32),
             pose1=np.asarray(pose1, float), pose2=np.asarray(pose2, float),
             tol=np.broadcast_to(np.asarray(tol, float), (len(idx1),)).copy(),
             alpha=res["alpha"], x=res["x"], iters=res["iters"].astype(np.int32),
             status=res["status"].astype(np.int32), grad=res["grad"], n=res["n"].astype(np.int32),
             m=res["m"].astype(np.int32))
    k = len(idx1) if keep_sz is None else keep_sz
    d["s"] = res["s"][:k]
    d["z"] = res["z"][:k]
    d["mu"] = res["mu"][:k]
    if extra:
        d.update(extra)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **d)
    st = np.bincount(res["status"].astype(int) + 1, minlength=6)
    ok = res["status"] == 0
    print(f"{name}: {len(idx1)} pairs -> {path} ({os.path.getsize(path) / 1024:.0f} KiB); "
          f"status counts [-1,0,1,2,3,4] = {st.tolist()}; iters mean {res['iters'][ok].mean() if ok.any() else float('nan'):.2f} "
          f"min {res['iters'][ok].min() if ok.any() else -1} max {res['iters'][ok].max() if ok.any() else -1}")


# ----------------------------------------------------------------------------------------------
def gen_scenarios():
    """Knot-0 collision pairs of the three shipped scenes."""
    import importlib
    import io
    import contextlib
    ref()
    cwd = os.getcwd()
    shape_list, idx1, idx2, pose1, pose2, names = [], [], [], [], [], []
    try:
        os.chdir("/root/reference")      # the quadrotor scene opens systems/polytopes.jld2 relative to cwd
        for mod, init, setter in (
            ("systems.piano_mover", "initialize_piano_mover", "inequality_constraints_x"),
            ("systems.cone_through_wall", "initialize_coneThroughWall", "inequality_constraints_x"),
            ("systems.cluttered_hallway_quadrotor", "initialize_quadrotor", "inequality_constraints_x"),
        ):
            m = importlib.import_module(mod)
            with contextlib.redirect_stdout(io.StringIO()):
                params, X, U = getattr(m, init)()
                getattr(m, setter)(params, X[0])       # puts the knot-0 pose into P_vic
            vic = params["P_vic"]
            v_idx = len(shape_list)
            shape_list.append(vic)
            for j, obs in enumerate(params["P_obs"]):
                shape_list.append(obs)
                idx1.append(v_idx)
                idx2.append(len(shape_list) - 1)
                pose1.append(S.pose_of(vic))
                pose2.append(S.pose_of(obs))
                names.append(f"{params['system']}[{j}]")
    finally:
        os.chdir(cwd)
    pose1, pose2 = np.array(pose1), np.array(pose2)
    res = run_batch(shape_list, idx1, idx2, pose1, pose2, 1e-6, procs=4)
    save("scenarios", shape_list, idx1, idx2, pose1, pose2, 1e-6, res, extra=dict(names=np.array(names)))
    for nm, a, it in zip(names, res["alpha"], res["iters"]):
        print(f"   {nm:22s} alpha={a!r} iters={it}")


def gen_config4(n_pairs=4000):
    shape_list, idx1, idx2, pose1, pose2 = W.config4_batch(n_pairs, exact=True)
    res = run_batch(shape_list, idx1, idx2, pose1, pose2, 1e-6)
    save("config4_sample", shape_list, idx1, idx2, pose1, pose2, 1e-6, res, keep_sz=400)


def gen_config5():
    shape_list, idx1, idx2, pose1, pose2 = W.config5_batch(n_obs=22, n_knots=10, n_cand=10)
    res = run_batch(shape_list, idx1, idx2, pose1, pose2, 1e-6)
    save("config5_sample", shape_list, idx1, idx2, pose1, pose2, 1e-6, res, keep_sz=220)


def gen_edge():
    """Edge behaviour: unsupported pairs, non-finite input, extreme separation, coincident centres,
    tolerance extremes, body-frame offsets, irregular polygon (quirk Q1), 14-face polytope."""
    rng = np.random.default_rng(77)
    base = W.config4_shapes()                      # 0 box 1 poly8 2 capsule 3 cylinder 4 cone 5 sphere 6 polygon5
    A1, b1, _, _ = W.hallway_polytopes()
    shape_list = list(base)
    shape_list.append(P.PolytopeMRP(A1, b1))                                   # 7: 14 faces
    irr = np.array([[1.0, 0.2], [0.1, 1.0], [-0.8, 0.6], [-0.5, -1.0], [0.7, -0.9]])
    irr /= np.linalg.norm(irr, axis=1, keepdims=True)
    shape_list.append(P.PolygonMRP(irr, np.array([0.5, 0.7, 0.4, 0.9, 0.6]), 0.15))   # 8: irregular 5-gon
    tri = P.create_n_sided(3, 0.4)
    shape_list.append(P.PolygonMRP(tri["A"], tri["b"], 0.1))                   # 9: triangle
    # body-frame offsets (never exercised by the shipped scenes, but part of the data model)
    def with_offset(prim, seed):
        g = np.random.default_rng(seed)
        prim.r_offset = g.normal(size=3) * 0.3
        q, _ = np.linalg.qr(g.normal(size=(3, 3)))
        prim.Q_offset = q * np.sign(np.linalg.det(q))
        return prim
    off0 = len(shape_list)
    shape_list += [with_offset(P.create_rect_prism(1.0, 2.0, 3.0), 1), with_offset(P.CapsuleMRP(0.3, 1.2), 2),
                   with_offset(P.CylinderMRP(0.4, 1.5), 3), with_offset(P.ConeMRP(2.0, np.deg2rad(22)), 4),
                   with_offset(P.SphereMRP(0.5), 5),
                   with_offset(P.PolygonMRP(P.create_n_sided(5, 0.6)["A"], P.create_n_sided(5, 0.6)["b"], 0.2), 6)]
    idx1, idx2, pose1, pose2, tol, tag = [], [], [], [], [], []

    def add(i, j, q1, q2, t=1e-6, name=""):
        idx1.append(i); idx2.append(j); pose1.append(np.asarray(q1, float)); pose2.append(np.asarray(q2, float))
        tol.append(t); tag.append(name)

    def rand_pose(scale_r):
        u = rng.normal(size=3)
        return np.concatenate([u / np.linalg.norm(u) * rng.uniform(0, scale_r), rng.normal(size=3) * 0.5])

    # (1) the nine unsupported pairs
    for i in (2, 3, 6):
        for j in (2, 3, 6):
            add(i, j, rand_pose(1), rand_pose(6), name="case4")
    # (2) non-finite input
    bad = rand_pose(1); bad[1] = np.nan
    add(0, 0, bad, rand_pose(6), name="nan_r")
    bad = rand_pose(6); bad[4] = np.nan
    add(5, 2, rand_pose(1), bad, name="nan_p")
    bad = rand_pose(6); bad[0] = np.inf
    add(4, 0, rand_pose(1), bad, name="inf_r")
    # (3) extreme separations, every kind against box and sphere
    for sep in (1e2, 1e4, 1e6, 1e9):
        for i in range(7):
            for j in (0, 5):
                if S.pair_supported(S.kind_of(shape_list[i]), S.kind_of(shape_list[j])):
                    q2 = rand_pose(1); q2[:3] = q2[:3] / np.linalg.norm(q2[:3]) * sep
                    add(i, j, rand_pose(1), q2, name=f"sep{sep:g}")
    # (4) coincident centres
    for i in range(7):
        for j in range(7):
            if S.pair_supported(S.kind_of(shape_list[i]), S.kind_of(shape_list[j])):
                q1 = rand_pose(1); q2 = rand_pose(1); q2[:3] = q1[:3]
                add(i, j, q1, q2, name="coincident")
    # (5) tolerance extremes
    for t in (1e-2, 1e-9, 1e-12, 0.0):
        for (i, j) in ((0, 0), (0, 1), (5, 5), (4, 0), (5, 3), (6, 1), (2, 4)):
            add(i, j, rand_pose(1), rand_pose(6), t=t, name=f"tol{t:g}")
    # (6) offsets, 14-face polytope, irregular polygons, triangle
    others = (0, 1, 4, 5)
    for k in range(6):
        for j in list(others) + [off0 + 0, off0 + 3, off0 + 4]:
            for rep in range(3):
                add(off0 + k, j, rand_pose(1), rand_pose(6), name="offset")
                add(j, off0 + k, rand_pose(1), rand_pose(6), name="offset")
    for special in (7, 8, 9):
        for j in (0, 1, 4, 5, 7):
            for rep in range(8):
                add(special, j, rand_pose(1), rand_pose(6), name=f"shape{special}")
                add(j, special, rand_pose(1), rand_pose(6), name=f"shape{special}")
    for j in (2, 3):        # 14-face polytope against the extras-carrying kinds
        for rep in range(8):
            add(7, j, rand_pose(1), rand_pose(6), name="shape7")
            add(j, 7, rand_pose(1), rand_pose(6), name="shape7")
    pose1, pose2 = np.array(pose1), np.array(pose2)
    res = run_batch(shape_list, idx1, idx2, pose1, pose2, np.array(tol))
    save("edge_cases", shape_list, idx1, idx2, pose1, pose2, np.array(tol), res, extra=dict(tag=np.array(tag)))
    for t in sorted(set(tag)):
        sel = np.array(tag) == t
        print(f"   {t:12s} n={sel.sum():4d} status={np.bincount(res['status'][sel], minlength=5).tolist()} "
              f"iters={sorted(set(res['iters'][sel].tolist()))[:12]}")


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    which = sys.argv[1:] or ["scenarios", "config4", "config5", "edge"]
    for w in which:
        {"scenarios": gen_scenarios, "config4": gen_config4, "config5": gen_config5, "edge": gen_edge}[w]()

"""The zero-edit drop-in of INTEGRATION.md section 1: with dropin/ ahead of the reference on sys.path, the
reference's own scenario scripts import OUR proximity functions and THEIR primitives.  Needs the reference
tree (present in the build container only); import-level, no GPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DCOL_REFERENCE_ROOT", "/root/reference")

SCRIPT = r'''
import sys, types
from unittest.mock import MagicMock
for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.mplot3d",
          "mpl_toolkits.mplot3d.art3d", "meshcat", "h5py"]:
    sys.modules[m] = MagicMock()
sys.path.insert(0, sys.argv[1]); sys.path.insert(1, sys.argv[2])
sys.dont_write_bytecode = True
import systems.piano_mover as pm, systems.cone_through_wall as cw, systems.cluttered_hallway_quadrotor as q
import ALTRO, primitives.misc_primitive_constructor as prims
for mod in (pm, cw, q):
    assert mod.proximity_mrp.__module__ == "dcol_trajectory_optimization_b200.proximity.proximity", mod.proximity_mrp.__module__
    assert mod.proximity_gradient.__module__ == "dcol_trajectory_optimization_b200.proximity.proximity_gradient"
assert prims.__file__.startswith(sys.argv[2])
# the reference's own primitive objects flatten into our shape table
from dcol_trajectory_optimization_b200.shapes import flatten_shapes
params, X, U = pm.initialize_piano_mover()
rec, A, b = flatten_shapes([params["P_vic"]] + params["P_obs"])
assert list(rec["type"]) == [0, 0, 0, 0] and A.shape == (24, 3)
print("DROPIN_OK")
'''


def test_reference_scripts_resolve_to_the_dropin():
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", SCRIPT, os.path.join(ROOT, "dcol_trajectory_optimization_b200", "dropin"), REF],
                       capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


HOOKS_SCRIPT = r'''
import sys, types
import numpy as np
from unittest.mock import MagicMock
for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.mplot3d",
          "mpl_toolkits.mplot3d.art3d", "meshcat"]:
    sys.modules[m] = MagicMock()
sys.path.insert(0, sys.argv[2] + "/oracle"); sys.path.insert(0, sys.argv[2])
from _refimport import import_reference
import_reference()                                   # the UNMODIFIED reference, its own proximity package
import oracle as O
from dcol_trajectory_optimization_b200.altro.reference_hooks import make_batched_hooks
from dcol_trajectory_optimization_b200.shapes import flatten_shapes, pose_of
import systems.piano_mover as pm, systems.cluttered_hallway_quadrotor as q, systems.cone_through_wall as cw
import os
os.chdir(sys.argv[1])
for mod, init in ((pm, pm.initialize_piano_mover), (q, q.initialize_quadrotor), (cw, cw.initialize_coneThroughWall)):
    params, X, U = init()
    prims = [params["P_vic"]] + list(params["P_obs"])
    rec, A, b = flatten_shapes(prims)
    obs = np.stack([pose_of(o) for o in params["P_obs"]]); n = len(obs)
    def ev(poses, want_grad):
        M = poses.shape[0]
        r = O.solve_batch(rec, A, b, np.zeros(M * n, np.int32), np.tile(np.arange(1, n + 1, dtype=np.int32), M),
                          np.repeat(poses, n, axis=0), np.tile(obs, (M, 1)), grad_mode=O.GRAD_EXACT if want_grad else O.GRAD_NONE)
        return r["alpha"].reshape(M, n), (r["grad"][:, :6].reshape(M, n, 6) if want_grad else None)
    hooks = make_batched_hooks(params, evaluator=ev)
    rng = np.random.default_rng(0)
    Xs = np.array(X[:4], dtype=float) + 0.05 * rng.normal(size=(4, params["nx"]))
    HX, GX = hooks.constraints_x_with_grad(Xs)
    for t in range(4):
        hx_ref = mod.inequality_constraints_x(params, Xs[t])
        gx_ref = mod.inequality_constraints_x_grad(params, Xs[t])
        assert np.abs(HX[t] - hx_ref).max() < 1e-9, (params["system"], np.abs(HX[t] - hx_ref).max())
        assert np.abs(GX[t] - gx_ref).max() < 2e-6 * max(1.0, np.abs(gx_ref).max()), (params["system"], np.abs(GX[t] - gx_ref).max())
print("HOOKS_OK")
'''


def test_batched_hooks_equal_the_reference_system_functions():
    """make_batched_hooks on the reference's own params dictionaries returns what its per-knot
    inequality_constraints_x / _grad return (values to 1e-9, gradients to the reference's FD noise)."""
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", HOOKS_SCRIPT, REF, ROOT], capture_output=True, text=True,
                       cwd=ROOT, timeout=600)
    assert r.returncode == 0 and "HOOKS_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-2500:]

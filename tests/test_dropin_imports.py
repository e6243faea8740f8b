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

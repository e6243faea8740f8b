"""TEST/DATA INFRASTRUCTURE, generation time only: extract the DATA of the reference's three scenarios
(obstacle poses pasted from the Julia original, initial control guesses, cone mass properties) by calling its
initialize_* functions, and store it as dcol_trajectory_optimization_b200/data/scenes.npz.  The scenario
definitions in altro/problems.py read that file; no reference code is copied."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refimport import import_reference  # noqa: E402


def main():
    import_reference()
    os.chdir("/root/reference")
    from systems.piano_mover import initialize_piano_mover
    from systems.cone_through_wall import initialize_coneThroughWall
    from systems.cluttered_hallway_quadrotor import initialize_quadrotor
    out = {}
    for name, init in (("piano", initialize_piano_mover), ("cone", initialize_coneThroughWall), ("quad", initialize_quadrotor)):
        params, X, U = init()
        out[name + "_U0"] = np.array(U, dtype=float)
        out[name + "_X0"] = np.array(X, dtype=float)
        out[name + "_obs_pose"] = np.array([np.concatenate([np.asarray(o.r, float), np.asarray(o.p, float)])
                                            for o in params["P_obs"]])
        out[name + "_Xref"] = np.array(params["Xref"], dtype=float)
        out[name + "_Uref"] = np.array(params["Uref"], dtype=float)
        if "m" in params:
            out[name + "_mass"] = float(params["m"])
            out[name + "_inertia"] = np.array(params["J"], dtype=float)
    path = os.path.join(os.path.dirname(HERE), "dcol_trajectory_optimization_b200", "data", "scenes.npz")
    np.savez_compressed(path, **out)
    print({k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()

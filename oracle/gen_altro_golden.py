"""TEST INFRASTRUCTURE, generation time only: run the UNMODIFIED reference ALTRO on one of its three
scenarios (main.py:39-52) in the build container and record the trajectory it converges to, so that a
batched drop-in can be checked for trajectory parity on the GPU box (where /root/reference does not exist).

    python oracle/gen_altro_golden.py piano_mover|coneThroughWall|quadrotor

Writes tests/golden/altro_<system>.npz: X_hist, U_hist (every iLQR pass), pass count, and the call
counts the authors' cProfile dumps pin (solve_lp_pdip, calc_NT_scalings).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refimport import import_reference  # noqa: E402


def main(system):
    ref = import_reference()
    scratch = os.path.join(os.path.dirname(HERE), "gpurun_out", "altro_ref_" + system)
    os.makedirs(scratch, exist_ok=True)
    os.chdir(scratch)                       # plots.py creates result_images/ under the cwd (matplotlib is stubbed)
    import ALTRO as A
    import proximity.pdip as pdip
    counts = {"nt": 0, "solve": 0}
    nt0, solve0 = pdip.calc_NT_scalings, pdip.solve_lp_pdip

    def nt(*a, **k):
        counts["nt"] += 1
        return nt0(*a, **k)

    def solve(*a, **k):
        counts["solve"] += 1
        return solve0(*a, **k)

    pdip.calc_NT_scalings = nt
    import proximity.proximity as pp
    import proximity.proximity_gradient as pg
    pp.solve_lp_pdip = solve
    pg.solve_lp_pdip = solve
    if system == "piano_mover":
        from systems.piano_mover import initialize_piano_mover as init
    elif system == "coneThroughWall":
        from systems.cone_through_wall import initialize_coneThroughWall as init
    elif system == "quadrotor":
        from systems.cluttered_hallway_quadrotor import initialize_quadrotor as init
    else:
        raise SystemExit("unknown system")
    params, X, U = init()
    X0, U0 = np.array(X, dtype=float), np.array(U, dtype=float)
    t = time.time()
    Xn, Un = A.ALTRO(params, X, U)
    wall = time.time() - t
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", f"altro_{system}.npz")
    np.savez_compressed(out, X=np.array(Xn, dtype=float), U=np.array(Un, dtype=float), X0=X0, U0=U0,
                        X_hist=np.array(params["X_hist"], dtype=float), U_hist=np.array(params["U_hist"], dtype=float),
                        n_passes=len(params["X_hist"]) - 1, n_nt=counts["nt"], n_solve=counts["solve"], wall_s=wall)
    print(system, "passes", len(params["X_hist"]) - 1, counts, f"{wall:.1f} s", flush=True)


if __name__ == "__main__":
    main(sys.argv[1])

"""ctypes binding of ``libdcol_altro.so`` (``include/dcol_altro.h``): the native host core of the batched
AL-iLQR caller — RK4 rollouts of all line-search candidates, forward-difference dynamics Jacobians + Riccati
recursion, augmented-Lagrangian cost — for the three systems of the reference
(``systems/piano_mover.py``, ``systems/cone_through_wall.py``, ``systems/cluttered_hallway_quadrotor.py``).

Host code only; the collision constraints still come from the CUDA engine.  ``altro_solve`` uses this core when
the problem names one of the built-in systems (``Problem.extra['native']``) and the NumPy implementation of the
same functions (``altro/solver.py``) for user-defined dynamics.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB_PATH = os.path.join(PKG, "libdcol_altro.so")
SYSTEMS = {"piano": 0, "rigid_body": 1, "quadrotor": 2}
E_NOT_PD = -3

#: every symbol include/dcol_altro.h declares
SYMBOLS = ("dcol_altro_version", "dcol_altro_dynamics", "dcol_altro_rk4", "dcol_altro_rollouts", "dcol_altro_jacobians",
           "dcol_altro_backward_pass", "dcol_altro_total_cost")

_lib = None


class _Problem(C.Structure):
    _fields_ = [("system", C.c_int32), ("nx", C.c_int32), ("nu", C.c_int32), ("N", C.c_int32), ("n_obs", C.c_int32),
                ("reserved", C.c_int32), ("dt", C.c_double), ("mass", C.c_double), ("inertia", C.c_double * 3),
                ("arm", C.c_double), ("kf", C.c_double), ("km", C.c_double), ("Q", C.c_void_p), ("R", C.c_void_p),
                ("Qf", C.c_void_p), ("Xref", C.c_void_p), ("Uref", C.c_void_p), ("u_min", C.c_void_p),
                ("u_max", C.c_void_p)]


def build() -> str:
    subprocess.run(["make", "-C", os.path.join(PKG, "csrc"), "../libdcol_altro.so"], check=True,
                   stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with make -C dcol_trajectory_optimization_b200/csrc")
        L = C.CDLL(LIB_PATH)
        vp, dp = C.c_void_p, C.c_void_p
        L.dcol_altro_version.restype = C.c_char_p
        L.dcol_altro_dynamics.argtypes = [vp, C.c_int64, dp, dp, dp]
        L.dcol_altro_rk4.argtypes = [vp, C.c_int64, dp, dp, dp]
        L.dcol_altro_rollouts.argtypes = [vp, dp, dp, dp, dp, dp, C.c_int32, dp, dp]
        L.dcol_altro_jacobians.argtypes = [vp, dp, dp, C.c_double, dp, dp]
        L.dcol_altro_backward_pass.argtypes = [vp, dp, dp, dp, dp, dp, dp, dp, C.c_double, C.c_double, dp, dp,
                                               C.POINTER(C.c_double)]
        L.dcol_altro_total_cost.argtypes = [vp, C.c_int32, dp, dp, dp, dp, dp, dp, C.c_double, dp]
        for name in SYMBOLS[1:]:
            getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class NativeCore:
    """The per-pass host work of one problem through the C ABI.  Keeps the problem's constant arrays alive."""

    def __init__(self, problem):
        spec = problem.extra["native"]
        self.N, self.nx, self.nu, self.n_obs = problem.N, problem.nx, problem.nu, problem.n_obs
        self._keep = [_f64(problem.Q), _f64(problem.R), _f64(problem.Qf), _f64(problem.Xref),
                      _f64(problem.Uref[:problem.N - 1]), _f64(problem.u_min), _f64(problem.u_max)]
        inertia = spec.get("inertia", (1.0, 1.0, 1.0))
        self._p = _Problem(SYSTEMS[spec["system"]], problem.nx, problem.nu, problem.N, problem.n_obs, 0, problem.dt,
                           float(spec.get("mass", 1.0)), (C.c_double * 3)(*[float(v) for v in inertia]),
                           float(spec.get("arm", 0.0)), float(spec.get("kf", 0.0)), float(spec.get("km", 0.0)),
                           *[a.ctypes.data for a in self._keep])
        self._ref = C.addressof(self._p)
        self._L = lib()

    @staticmethod
    def _check(rc):
        if rc == E_NOT_PD:           # scipy.linalg.cho_factor(Quu) in the reference, ALTRO.py:321
            raise np.linalg.LinAlgError("Quu is not positive definite")
        if rc != 0:
            raise ValueError(f"dcol_altro: bad argument ({rc})")

    def rk4(self, X, U):
        X, U = _f64(X).reshape(-1, self.nx), _f64(U).reshape(-1, self.nu)
        out = np.empty_like(X)
        self._check(self._L.dcol_altro_rk4(self._ref, X.shape[0], X.ctypes.data, U.ctypes.data, out.ctypes.data))
        return out

    def dynamics(self, X, U):
        X, U = _f64(X).reshape(-1, self.nx), _f64(U).reshape(-1, self.nu)
        out = np.empty_like(X)
        self._check(self._L.dcol_altro_dynamics(self._ref, X.shape[0], X.ctypes.data, U.ctypes.data, out.ctypes.data))
        return out

    def rollouts(self, X, U, K, k, alphas):
        X, U, K, k, alphas = _f64(X), _f64(U), _f64(K), _f64(k), _f64(alphas)
        Cn = alphas.shape[0]
        Xn = np.empty((Cn, self.N, self.nx))
        Un = np.empty((Cn, self.N - 1, self.nu))
        self._check(self._L.dcol_altro_rollouts(self._ref, X.ctypes.data, U.ctypes.data, K.ctypes.data, k.ctypes.data,
                                                alphas.ctypes.data, Cn, Xn.ctypes.data, Un.ctypes.data))
        return Xn, Un

    def jacobians(self, X, U, delta=1e-6):
        X, U = _f64(X), _f64(U)
        A = np.empty((self.N - 1, self.nx, self.nx))
        B = np.empty((self.N - 1, self.nx, self.nu))
        self._check(self._L.dcol_altro_jacobians(self._ref, X.ctypes.data, U.ctypes.data, float(delta), A.ctypes.data,
                                                 B.ctypes.data))
        return A, B

    def backward_pass(self, X, U, hx, ghx, mu, mux, lambd, rho, reg):
        X, U, hx, ghx, mu, mux, lambd = (_f64(a) for a in (X, U, hx, ghx, mu, mux, lambd))
        K = np.empty((self.N - 1, self.nu, self.nx))
        k = np.empty((self.N - 1, self.nu))
        dJ = C.c_double()
        self._check(self._L.dcol_altro_backward_pass(self._ref, X.ctypes.data, U.ctypes.data, hx.ctypes.data,
                                                     ghx.ctypes.data, mu.ctypes.data, mux.ctypes.data,
                                                     lambd.ctypes.data, float(rho), float(reg), K.ctypes.data,
                                                     k.ctypes.data, C.byref(dJ)))
        return K, k, dJ.value

    def total_cost(self, X, U, hx, mu, mux, lambd, rho):
        X, U, hx, mu, mux, lambd = (_f64(a) for a in (X, U, hx, mu, mux, lambd))
        lead = X.shape[:-2]
        Cn = int(np.prod(lead)) if lead else 1
        cost = np.empty(Cn)
        self._check(self._L.dcol_altro_total_cost(self._ref, Cn, X.ctypes.data, U.ctypes.data, hx.ctypes.data,
                                                  mu.ctypes.data, mux.ctypes.data, lambd.ctypes.data, float(rho),
                                                  cost.ctypes.data))
        return cost.reshape(lead) if lead else cost[0]

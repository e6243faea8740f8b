"""TEST INFRASTRUCTURE: ctypes front end of the host twin (tests/host_twin/host_twin.cpp), the
product's device solver header compiled as plain C++ for CPU-side debugging of the algorithm."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB_PATH = os.path.join(HERE, "libdcol_twin.so")
SOURCES = [os.path.join(HERE, "host_twin.cpp"),
           os.path.join(ROOT, "dcol_trajectory_optimization_b200", "csrc", "dcol_solver.cuh"),
           os.path.join(ROOT, "dcol_trajectory_optimization_b200", "csrc", "dcol_classes.cuh"),
           os.path.join(ROOT, "include", "dcol.h")]
_lib = None


def build(force=False, opt="-O0"):
    stale = (not os.path.exists(LIB_PATH)
             or any(os.path.getmtime(LIB_PATH) < os.path.getmtime(s) for s in SOURCES))
    if force or stale:
        subprocess.run(["g++", opt, "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-pthread",
                        "-o", LIB_PATH, SOURCES[0], "-lm"], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.dcol_twin_batch.restype = C.c_int
        L.dcol_twin_batch.argtypes = [C.c_void_p, dp, dp, ip, ip, dp, dp, C.c_int64, C.c_double, C.c_int, C.c_int,
                                      dp, dp, dp, ip, ip]
        L.dcol_twin_batch_jac.restype = C.c_int
        L.dcol_twin_batch_jac.argtypes = L.dcol_twin_batch.argtypes + [dp]
        L.dcol_twin_pair.restype = C.c_int
        L.dcol_twin_pair.argtypes = [C.c_void_p, dp, dp, C.c_int32, C.c_int32, dp, dp, C.c_double, dp, dp, dp, dp,
                                     ip, ip, ip, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _table(records, A, b):
    records = np.ascontiguousarray(records)
    A = np.ascontiguousarray(A, dtype=np.float64).reshape(-1, 3)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
    if A.shape[0] == 0:
        A, b = np.zeros((1, 3)), np.zeros(1)
    return records, A, b


def solve_batch(records, A, b, idx1, idx2, pose1, pose2, tol=1e-6, max_iter=50, threads=None, want_grad=True,
                fix_case4=False, want_jac=False):
    records, A, b = _table(records, A, b)
    idx1 = np.ascontiguousarray(idx1, dtype=np.int32)
    idx2 = np.ascontiguousarray(idx2, dtype=np.int32)
    pose1 = np.ascontiguousarray(pose1, dtype=np.float64).reshape(-1, 6)
    pose2 = np.ascontiguousarray(pose2, dtype=np.float64).reshape(-1, 6)
    B = idx1.shape[0]
    alpha, contact = np.empty(B), np.empty((B, 3))
    grad = np.empty((B, 12)) if want_grad else None
    iters, status = np.empty(B, np.int32), np.empty(B, np.int32)
    jac = np.empty((B, 4, 12)) if want_jac else None
    lib().dcol_twin_set_fix_case4(1 if fix_case4 else 0)
    args = [records.ctypes.data, _dp(A), _dp(b), _ip(idx1), _ip(idx2), _dp(pose1), _dp(pose2), B,
            float(tol), int(max_iter), int(threads or os.cpu_count() or 1), _dp(alpha), _dp(contact),
            _dp(grad) if want_grad else C.POINTER(C.c_double)(), _ip(iters), _ip(status)]
    if want_jac:
        lib().dcol_twin_batch_jac(*args, _dp(jac))
    else:
        lib().dcol_twin_batch(*args)
    lib().dcol_twin_set_fix_case4(0)
    return dict(alpha=alpha, contact=contact, grad=grad, iters=iters, status=status, jac=jac)


def solve_pair(records, A, b, i1, i2, pose1, pose2, tol=1e-6):
    records, A, b = _table(records, A, b)
    pose1 = np.ascontiguousarray(pose1, dtype=np.float64)
    pose2 = np.ascontiguousarray(pose2, dtype=np.float64)
    alpha = C.c_double()
    x, s, z = np.full(8, np.nan), np.full(72, np.nan), np.full(72, np.nan)
    grad, mu = np.full(12, np.nan), np.full(51, np.nan)
    n, m, iters = C.c_int32(), C.c_int32(), C.c_int32()
    st = lib().dcol_twin_pair(records.ctypes.data, _dp(A), _dp(b), int(i1), int(i2), _dp(pose1), _dp(pose2),
                              float(tol), C.byref(alpha), _dp(x), _dp(s), _dp(z), C.byref(n), C.byref(m),
                              C.byref(iters), _dp(grad), _dp(mu))
    return dict(alpha=alpha.value, x=x[:n.value], s=s[:m.value], z=z[:m.value], iters=iters.value, status=st,
                grad=grad, mu=mu, n=n.value, m=m.value)

"""TEST INFRASTRUCTURE: ctypes front end of the CPU oracle (oracle/dcol_oracle.c).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this
module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdcol_oracle.so")
LIB_PATH_FMA = os.path.join(HERE, "libdcol_oracle_fma.so")   # same source, fused multiply-adds allowed

GRAD_NONE, GRAD_FD, GRAD_EXACT = 0, 1, 2
MAX_M, MAX_N, MAX_TRACE = 72, 8, 51

_libs = {}


def build(force: bool = False) -> str:
    """Compile the oracle (both roundings) with the committed Makefile (gcc only, a second or two)."""
    src = os.path.join(HERE, "dcol_oracle.c")
    stale = any(not os.path.exists(p) or os.path.getmtime(p) < os.path.getmtime(src) for p in (LIB_PATH, LIB_PATH_FMA))
    if force or stale:
        subprocess.run(["make", "-C", HERE, "-s"] + (["-B"] if force else []), check=True)
    return LIB_PATH


def lib(fma: bool = False):
    _lib = _libs.get(fma)
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH_FMA if fma else LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.dcol_oracle_pair.restype = C.c_int
        L.dcol_oracle_pair.argtypes = [C.c_void_p, dp, dp, C.c_int32, C.c_int32, dp, dp, C.c_double, C.c_int,
                                       dp, dp, dp, dp, ip, ip, ip, dp, dp]
        L.dcol_oracle_batch.restype = C.c_int
        L.dcol_oracle_batch.argtypes = [C.c_void_p, dp, dp, ip, ip, dp, dp, C.c_int64, C.c_double, C.c_int, C.c_int,
                                        dp, dp, dp, ip, ip]
        L.dcol_oracle_assemble.restype = C.c_int
        L.dcol_oracle_assemble.argtypes = [C.c_void_p, dp, dp, C.c_int32, C.c_int32, dp, dp, dp, dp, dp, ip]
        L.dcol_oracle_set_fix_case4.restype = None
        L.dcol_oracle_set_fix_case4.argtypes = [C.c_int]
        L.dcol_oracle_dcm.restype = None
        L.dcol_oracle_dcm.argtypes = [dp, dp, dp]
        _lib = _libs[fma] = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _table(records, A, b):
    records = np.ascontiguousarray(records)
    A = np.ascontiguousarray(A, dtype=np.float64).reshape(-1, 3)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
    if A.shape[0] == 0:      # keep the pointers valid
        A = np.zeros((1, 3))
        b = np.zeros(1)
    return records, A, b


def solve_pair(records, A, b, i1, i2, pose1, pose2, tol=1e-6, grad_mode=GRAD_FD):
    """One pair -> dict(alpha, x[n], s[m], z[m], iters, status, grad[12], mu[51])."""
    records, A, b = _table(records, A, b)
    pose1 = np.ascontiguousarray(pose1, dtype=np.float64)
    pose2 = np.ascontiguousarray(pose2, dtype=np.float64)
    alpha = C.c_double()
    x, s, z = np.full(MAX_N, np.nan), np.full(MAX_M, np.nan), np.full(MAX_M, np.nan)
    grad, mu = np.full(12, np.nan), np.full(MAX_TRACE, np.nan)
    n, m, iters = C.c_int32(), C.c_int32(), C.c_int32()
    st = lib().dcol_oracle_pair(records.ctypes.data, _dp(A), _dp(b), int(i1), int(i2), _dp(pose1), _dp(pose2),
                                float(tol), int(grad_mode), C.byref(alpha), _dp(x), _dp(s), _dp(z), C.byref(n),
                                C.byref(m), C.byref(iters), _dp(grad), _dp(mu))
    return dict(alpha=alpha.value, x=x[:n.value], s=s[:m.value], z=z[:m.value], iters=iters.value, status=st,
                grad=grad, mu=mu, n=n.value, m=m.value)


def solve_batch(records, A, b, idx1, idx2, pose1, pose2, tol=1e-6, grad_mode=GRAD_FD, threads=None,
                want_contact=True, fma=False, fix_case4=False):
    """Batch -> dict(alpha[B], contact[B,3], grad[B,12], iters[B], status[B]).
    ``fma=True`` runs the fused-multiply-add build (rounding-sensitivity probe, see Makefile)."""
    records, A, b = _table(records, A, b)
    idx1 = np.ascontiguousarray(idx1, dtype=np.int32)
    idx2 = np.ascontiguousarray(idx2, dtype=np.int32)
    pose1 = np.ascontiguousarray(pose1, dtype=np.float64).reshape(-1, 6)
    pose2 = np.ascontiguousarray(pose2, dtype=np.float64).reshape(-1, 6)
    B = idx1.shape[0]
    alpha = np.empty(B)
    contact = np.empty((B, 3)) if want_contact else None
    grad = np.empty((B, 12)) if grad_mode != GRAD_NONE else None
    iters = np.empty(B, dtype=np.int32)
    status = np.empty(B, dtype=np.int32)
    null = C.POINTER(C.c_double)()
    lib(fma).dcol_oracle_set_fix_case4(1 if fix_case4 else 0)   # extension switch, see dcol_oracle.c
    lib(fma).dcol_oracle_batch(records.ctypes.data, _dp(A), _dp(b), _ip(idx1), _ip(idx2), _dp(pose1), _dp(pose2), B,
                            float(tol), int(grad_mode), int(threads or os.cpu_count() or 1), _dp(alpha),
                            _dp(contact) if contact is not None else null, _dp(grad) if grad is not None else null,
                            _ip(iters), _ip(status))
    lib(fma).dcol_oracle_set_fix_case4(0)
    return dict(alpha=alpha, contact=contact, grad=grad, iters=iters, status=status)


def assemble(records, A, b, i1, i2, pose1, pose2):
    """Assembled conic program of one pair -> (status, c[n], G[m,n], h[m], (n, m, n_ort, q1, q2))."""
    records, A, b = _table(records, A, b)
    pose1 = np.ascontiguousarray(pose1, dtype=np.float64)
    pose2 = np.ascontiguousarray(pose2, dtype=np.float64)
    c, G, h = np.zeros(MAX_N), np.zeros((MAX_M, MAX_N)), np.zeros(MAX_M)
    dims = np.zeros(5, dtype=np.int32)
    st = lib().dcol_oracle_assemble(records.ctypes.data, _dp(A), _dp(b), int(i1), int(i2), _dp(pose1), _dp(pose2),
                                    _dp(c), _dp(G), _dp(h), _ip(dims))
    n, m = int(dims[0]), int(dims[1])
    return st, c[:n], G[:m, :n], h[:m], tuple(int(v) for v in dims)


def dcm(p):
    """(Q[3,3], dQ[3,3,3]) with dQ[k] = dQ/dp_k."""
    p = np.ascontiguousarray(p, dtype=np.float64)
    Q, dQ = np.zeros((3, 3)), np.zeros((3, 3, 3))
    lib().dcol_oracle_dcm(_dp(p), _dp(Q), _dp(dQ))
    return Q, dQ

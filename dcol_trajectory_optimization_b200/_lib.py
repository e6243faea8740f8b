"""ctypes binding of ``libdcol_b200.so`` (the C ABI declared in ``include/dcol.h``).

The library holds the hand-written sm_100a kernels; there is no CPU fallback.  ``lib()`` raises
``RuntimeError`` if the shared object has not been built, and every compute entry point of the
library itself fails with ``DCOL_E_NOGPU`` when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCOL_LIB") or os.path.join(HERE, "libdcol_b200.so")   # DCOL_LIB: build experiments
CSRC = os.path.join(HERE, "csrc")

WANT_CONTACT, WANT_GRAD, FIX_CASE4, DEST_MULTICAST, ONE_PAIR_PER_THREAD, WANT_GRAD1, LANE_REFILL = 1, 2, 4, 8, 16, 32, 64
MAX_ITER = 50
MAX_M, MAX_N = 72, 8
MAX_DEST, RECORD_WORDS = 8, 14

#: every symbol include/dcol.h declares (checked by the CPU test-suite against the built library)
SYMBOLS = (
    "dcol_version", "dcol_last_error", "dcol_device_count", "dcol_shape_table_create", "dcol_shape_table_destroy",
    "dcol_plan_create", "dcol_plan_destroy", "dcol_plan_size", "dcol_plan_n_groups", "dcol_plan_n_launches", "dcol_plan_refine",
    "dcol_proximity_batch_device", "dcol_proximity_batch_jacobian", "dcol_proximity_batch_host",
    "dcol_proximity_scene_host", "dcol_host_alloc",
    "dcol_host_free", "dcol_release_cached",
    "dcol_proximity_batch_records", "dcol_plan_perm", "dcol_device_alloc", "dcol_device_free", "dcol_ipc_export",
    "dcol_ipc_import", "dcol_ipc_close",
    "dcol_debug_trace_pair", "dcol_measure_fp64_peak",
)

_lib = None


class DcolError(RuntimeError):
    """A C-ABI call failed (argument error < 0, CUDA error > 0)."""

    def __init__(self, code, message):
        super().__init__(f"dcol error {code}: {message}")
        self.code = code


def build(jobs: int | None = None, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree with the committed Makefile (nvcc, sm_100a)."""
    cmd = ["make", "-C", CSRC, f"-j{jobs or os.cpu_count() or 4}"]
    subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def lib():
    """The loaded library with argument types set; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
            "g.build()' or make -C dcol_trajectory_optimization_b200/csrc).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses: device or host
    L.dcol_version.restype = C.c_char_p
    L.dcol_last_error.restype = C.c_char_p
    L.dcol_device_count.restype = C.c_int
    L.dcol_shape_table_create.restype = C.c_int
    L.dcol_shape_table_create.argtypes = [vp, C.c_int32, dp, dp, C.c_int32, C.c_int, C.POINTER(C.c_void_p)]
    L.dcol_shape_table_destroy.restype = None
    L.dcol_shape_table_destroy.argtypes = [vp]
    L.dcol_plan_create.restype = C.c_int
    L.dcol_plan_create.argtypes = [vp, ip, ip, C.c_int64, vp, C.POINTER(C.c_void_p)]
    L.dcol_plan_destroy.restype = None
    L.dcol_plan_destroy.argtypes = [vp]
    L.dcol_plan_size.restype = C.c_int64
    L.dcol_plan_size.argtypes = [vp]
    L.dcol_plan_n_groups.restype = C.c_int32
    L.dcol_plan_n_groups.argtypes = [vp]
    L.dcol_plan_n_launches.restype = C.c_int32
    L.dcol_plan_n_launches.argtypes = [vp]
    L.dcol_plan_refine.restype = C.c_int
    L.dcol_plan_refine.argtypes = [vp, ip, vp]
    L.dcol_proximity_batch_device.restype = C.c_int
    L.dcol_proximity_batch_device.argtypes = [vp, dp, dp, C.c_double, C.c_int32, C.c_uint32, dp, dp, dp, ip, ip, vp]
    L.dcol_proximity_batch_jacobian.restype = C.c_int
    L.dcol_proximity_batch_jacobian.argtypes = [vp, dp, dp, C.c_double, C.c_int32, C.c_uint32, dp, dp, dp, dp, ip, ip, vp]
    L.dcol_proximity_batch_host.restype = C.c_int
    L.dcol_proximity_batch_host.argtypes = [vp, ip, ip, dp, dp, C.c_int64, C.c_double, C.c_int32, C.c_uint32,
                                            dp, dp, dp, ip, ip]
    L.dcol_proximity_scene_host.restype = C.c_int
    L.dcol_proximity_scene_host.argtypes = [vp, C.c_int32, dp, C.c_int64, ip, dp, C.c_int32, C.c_double, C.c_int32, C.c_uint32,
                                            dp, dp, ip, ip]
    L.dcol_proximity_batch_records.restype = C.c_int
    L.dcol_proximity_batch_records.argtypes = [vp, dp, dp, C.c_double, C.c_int32, C.c_uint32, C.c_int32, C.POINTER(C.c_void_p),
                                               C.c_int64, dp, vp]
    L.dcol_plan_perm.restype = C.c_void_p
    L.dcol_plan_perm.argtypes = [vp]
    L.dcol_device_alloc.restype = C.c_int
    L.dcol_device_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]
    L.dcol_device_free.restype = None
    L.dcol_device_free.argtypes = [C.c_int, vp]
    L.dcol_ipc_export.restype = C.c_int
    L.dcol_ipc_export.argtypes = [C.c_int, vp, vp]
    L.dcol_ipc_import.restype = C.c_int
    L.dcol_ipc_import.argtypes = [C.c_int, vp, C.POINTER(C.c_void_p)]
    L.dcol_ipc_close.restype = None
    L.dcol_ipc_close.argtypes = [C.c_int, vp]
    L.dcol_host_alloc.restype = C.c_int
    L.dcol_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.dcol_host_free.restype = None
    L.dcol_host_free.argtypes = [vp]
    L.dcol_release_cached.restype = None
    L.dcol_release_cached.argtypes = []
    L.dcol_debug_trace_pair.restype = C.c_int
    L.dcol_debug_trace_pair.argtypes = [vp, C.c_int32, C.c_int32, dp, dp, C.c_double, dp, dp, dp, dp, ip, ip, ip, ip, dp]
    L.dcol_measure_fp64_peak.restype = C.c_int
    L.dcol_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise DcolError(rc, lib().dcol_last_error().decode())

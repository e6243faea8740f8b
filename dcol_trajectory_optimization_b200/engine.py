"""Batched proximity solve + gradient: the Python host above the C ABI.

New entry points (the reference evaluates pairs one by one in Python loops,
``systems/cluttered_hallway_quadrotor.py:131-133,155``):

* :class:`ProximityEngine` — owns a shape table on one GPU; ``plan()`` groups a batch of
  (shape, shape) index pairs once, ``solve()`` runs every pair of the plan on device buffers,
  ``solve_host()`` takes NumPy buffers through the chunked copy/solve pipeline of the C ABI.
* :func:`proximity_batch` — convenience over lists of primitive objects.

PyTorch is only the buffer carrier (device memory, streams); all arithmetic happens in
``libdcol_b200.so``.  Status words follow ``include/dcol.h``; :func:`raise_for_status` maps them to
the exception classes the reference raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .shapes import (STATUS_MAX_ITER, STATUS_NON_FINITE, STATUS_NOT_PD, STATUS_OK, STATUS_UNSUPPORTED, flatten_shapes,
                     pose_of)

WANT_CONTACT, WANT_GRAD, FIX_CASE4 = _lib.WANT_CONTACT, _lib.WANT_GRAD, _lib.FIX_CASE4
ONE_PAIR_PER_THREAD, WANT_GRAD1, LANE_REFILL = _lib.ONE_PAIR_PER_THREAD, _lib.WANT_GRAD1, _lib.LANE_REFILL


def raise_for_status(status: int):
    """Re-raise what the reference raises for a failed solve (pdip.py:470, SciPy check_finite,
    numpy/scipy Cholesky, combine_problem_matrices.py:58-67)."""
    if status == STATUS_OK:
        return
    if status == STATUS_MAX_ITER:
        raise Exception("Maximum number of iterations reached, PDIP failed")
    if status == STATUS_NON_FINITE:
        raise ValueError("array must not contain infs or NaNs")
    if status == STATUS_NOT_PD:
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    if status == STATUS_UNSUPPORTED:
        raise ValueError("all the input array dimensions except for the concatenation axis must match exactly "
                         "(both primitives carry extra decision variables)")
    raise RuntimeError(f"unknown DCOL status {status}")


@dataclass
class BatchResult:
    """Outputs of one batched solve (torch CUDA tensors from ``solve``, NumPy arrays from ``solve_host``)."""
    alpha: object      # [B]      proximity value, NaN where status != 0
    contact: object    # [B, 3]   x[0:3] of the solution, or None
    grad: object       # [B, 12]  d alpha / d [r1 p1 r2 p2], or None
    iters: object      # [B]      PDIP iterations taken
    status: object     # [B]      DCOL_STATUS_*
    jac: object = None  # [B, 4, 12] d(contact x, y, z, alpha) / d [r1 p1 r2 p2] (``solve(want_jac=True)`` only)

    def summary(self) -> dict:
        """Per-batch counters (the reference has no logging inside the solver): how many pairs
        ended in each status and the histogram of PDIP iteration counts of the converged ones."""
        status = self.status.cpu().numpy() if hasattr(self.status, "cpu") else np.asarray(self.status)
        iters = self.iters.cpu().numpy() if hasattr(self.iters, "cpu") else np.asarray(self.iters)
        names = ("ok", "max_iter", "non_finite", "not_pd", "unsupported")
        counts = np.bincount(status, minlength=5)
        ok = status == STATUS_OK
        hist = np.bincount(iters[ok], minlength=1) if ok.any() else np.zeros(1, dtype=np.int64)
        return {"pairs": int(status.shape[0]), "status": {n: int(c) for n, c in zip(names, counts)},
                "iters_mean": float(iters[ok].mean()) if ok.any() else float("nan"),
                "iters_max": int(iters[ok].max()) if ok.any() else 0,
                "iters_hist": {int(i): int(c) for i, c in enumerate(hist) if c}}


@dataclass
class SceneResult:
    """Outputs of ``ProximityEngine.solve_scene_host`` (NumPy arrays)."""
    alpha: object      # [M, n_obs]
    grad1: object      # [M, n_obs, 6] d alpha / d [r1 p1] (victim pose), or None
    iters: object      # [M, n_obs] or None
    status: object     # [M, n_obs]


class _DevArray:
    """A raw device address exposed through ``__cuda_array_interface__`` so torch can view it."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def device_view(ptr: int, shape, device, typestr="<f8"):
    """Torch view (no copy) of a raw device allocation, e.g. one made by ``dcol_device_alloc``."""
    import torch
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


def records_to_result(rec, perm=None) -> BatchResult:
    """Views of a ``[B, 14]`` float64 record tensor (plan order); with ``perm`` the fields are scattered back
    to pair order (copies)."""
    import torch
    alpha, grad = rec[:, 0], rec[:, 1:13]
    ints = rec[:, 13].contiguous().view(torch.int32).view(-1, 2)
    iters, status = ints[:, 0], ints[:, 1]
    if perm is None:
        return BatchResult(alpha=alpha, contact=None, grad=grad, iters=iters, status=status)
    p = perm.long()
    out = BatchResult(alpha=torch.empty_like(alpha), contact=None, grad=torch.empty_like(grad),
                      iters=torch.empty_like(iters), status=torch.empty_like(status))
    out.alpha[p], out.grad[p], out.iters[p], out.status[p] = alpha, grad, iters, status
    return out


class Plan:
    """A batch's pairs grouped by shape pair (device counting sort), reusable across solves."""

    def __init__(self, engine, idx1, idx2):
        import torch
        self.engine = engine
        self.idx1 = idx1.to(device=engine.device, dtype=torch.int32).contiguous()
        self.idx2 = idx2.to(device=engine.device, dtype=torch.int32).contiguous()
        if self.idx1.shape != self.idx2.shape or self.idx1.dim() != 1:
            raise ValueError("idx1 and idx2 must be 1-D and of equal length")
        self.size = int(self.idx1.shape[0])
        handle = C.c_void_p()
        stream = torch.cuda.current_stream(engine.device).cuda_stream
        _lib.check(_lib.lib().dcol_plan_create(engine._table, self.idx1.data_ptr(), self.idx2.data_ptr(), self.size,
                                               stream, C.byref(handle)))
        self._handle = handle
        self.n_groups = int(_lib.lib().dcol_plan_n_groups(handle))
        self.n_launches = int(_lib.lib().dcol_plan_n_launches(handle))

    def perm(self):
        """int32 CUDA tensor ``[size]``: ``perm[i]`` = index of the i-th pair in plan order (a copy)."""
        import torch
        ptr = _lib.lib().dcol_plan_perm(self._handle)
        if not ptr or self.size == 0:
            return torch.zeros(0, dtype=torch.int32, device=self.engine.device)
        return torch.as_tensor(_DevArray(ptr, (self.size,), "<i4"), device=self.engine.device).clone()

    def refine(self, iters):
        """Re-order the plan for the next solve of the same pair list: inside every group, pairs are sorted (longest first) by the
        iteration counts ``iters`` (int32 CUDA tensor ``[size]``, pair order — ``BatchResult.iters`` of the previous
        solve).  For callers that re-solve a fixed pair list with slowly changing poses (AL-iLQR passes): warps then
        hold pairs of nearly equal iteration count.  Results are unaffected.  Enqueues on the current stream."""
        import torch
        if not (iters.is_cuda and iters.dtype == torch.int32 and iters.is_contiguous() and iters.numel() == self.size):
            raise ValueError(f"iters must be a contiguous int32 CUDA tensor of {self.size} elements")
        self.engine._check_open(self)
        stream = torch.cuda.current_stream(self.engine.device).cuda_stream
        _lib.check(_lib.lib().dcol_plan_refine(self._handle, iters.data_ptr(), stream))

    def close(self):
        if getattr(self, "_handle", None):
            _lib.lib().dcol_plan_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ProximityEngine:
    """Shape table on one CUDA device + batched solves against it."""

    def __init__(self, shapes, device: int = 0):
        L = _lib.lib()
        if L.dcol_device_count() <= 0:
            raise RuntimeError("no CUDA device visible: the DCOL B200 engine has no CPU fallback")
        import torch
        self.device = torch.device("cuda", device)
        if isinstance(shapes, tuple) and len(shapes) == 3:
            self.records, self.A, self.b = shapes
        else:
            self.records, self.A, self.b = flatten_shapes(list(shapes))
        self.records = np.ascontiguousarray(self.records)
        self.A = np.ascontiguousarray(self.A, dtype=np.float64).reshape(-1, 3)
        self.b = np.ascontiguousarray(self.b, dtype=np.float64).reshape(-1)
        handle = C.c_void_p()
        nf = int(self.b.shape[0])
        _lib.check(L.dcol_shape_table_create(self.records.ctypes.data, len(self.records),
                                             self.A.ctypes.data if nf else None, self.b.ctypes.data if nf else None,
                                             nf, device, C.byref(handle)))
        self._table = handle

    def close(self):
        if getattr(self, "_table", None):
            _lib.lib().dcol_shape_table_destroy(self._table)
            self._table = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ device buffers
    def _check_open(self, plan=None):
        """A plan borrows the engine's shape table: using either after ``close()`` is an error, not undefined behaviour."""
        if getattr(self, "_table", None) is None:
            raise RuntimeError("this ProximityEngine has been closed")
        if plan is not None and (plan.engine is not self or getattr(plan, "_handle", None) is None):
            raise RuntimeError("the plan is closed or belongs to another engine")

    def plan(self, idx1, idx2) -> Plan:
        import torch
        self._check_open()
        return Plan(self, torch.as_tensor(idx1), torch.as_tensor(idx2))

    def solve(self, plan: Plan, pose1, pose2, tol: float = 1e-6, max_iter: int = 50, want_grad: bool = True,
              want_contact: bool = True, out: BatchResult | None = None, fix_case4: bool = False,
              want_jac: bool = False, grad1: bool = False, one_pair_per_thread: bool = False,
              lane_refill: bool = False) -> BatchResult:
        """Enqueue the solve of every pair of ``plan`` on the current CUDA stream.

        ``pose1``/``pose2``: float64 CUDA tensors ``[B, 6]`` (rows ``r, p``).  Returns CUDA tensors;
        the call does not synchronise.  ``want_jac`` (extension) also returns the
        solution Jacobian ``jac[B, 4, 12]`` = d(contact point, alpha) / d[r1 p1 r2 p2], computed inside the solve
        kernel by an adjoint solve with the factor of the final reduced KKT matrix
        (``dcol_proximity_batch_jacobian``).  ``grad1``: the gradient is ``[B, 6]``, d alpha / d[r1 p1] only
        (``DCOL_WANT_GRAD1``: what the reference's callers consume).  ``one_pair_per_thread``: use the kernels
        without lane refill, ``lane_refill``: force the lane-refill kernels (neither: the library's default)."""
        import torch
        self._check_open(plan)
        B = plan.size
        if grad1 and want_jac:
            raise ValueError("grad1 is not available together with want_jac")
        for name, t in (("pose1", pose1), ("pose2", pose2)):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == (B, 6)):
                raise ValueError(f"{name} must be a contiguous float64 CUDA tensor of shape ({B}, 6)")
        if out is None:
            dev = self.device
            out = BatchResult(
                alpha=torch.empty(B, dtype=torch.float64, device=dev),
                contact=torch.empty((B, 3), dtype=torch.float64, device=dev) if want_contact else None,
                grad=torch.empty((B, 6 if grad1 else 12), dtype=torch.float64, device=dev) if want_grad else None,
                iters=torch.empty(B, dtype=torch.int32, device=dev),
                status=torch.empty(B, dtype=torch.int32, device=dev),
                jac=torch.empty((B, 4, 12), dtype=torch.float64, device=dev) if want_jac else None)
        if out.grad is not None and tuple(out.grad.shape) != (B, 6 if grad1 else 12):
            raise ValueError(f"out.grad must have shape ({B}, {6 if grad1 else 12})")
        flags = ((WANT_CONTACT if out.contact is not None else 0) | (WANT_GRAD if out.grad is not None else 0)
                 | (FIX_CASE4 if fix_case4 else 0) | (WANT_GRAD1 if (grad1 and out.grad is not None) else 0)
                 | (ONE_PAIR_PER_THREAD if one_pair_per_thread else 0) | (LANE_REFILL if lane_refill else 0))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if want_jac:
            if out.jac is None:
                raise ValueError("want_jac needs out.jac ([B, 4, 12] float64 CUDA tensor)")
            _lib.check(_lib.lib().dcol_proximity_batch_jacobian(
                plan._handle, pose1.data_ptr(), pose2.data_ptr(), float(tol), int(max_iter), flags, out.alpha.data_ptr(),
                out.contact.data_ptr() if out.contact is not None else None,
                out.grad.data_ptr() if out.grad is not None else None, out.jac.data_ptr(), out.iters.data_ptr(),
                out.status.data_ptr(), stream))
            return out
        _lib.check(_lib.lib().dcol_proximity_batch_device(
            plan._handle, pose1.data_ptr(), pose2.data_ptr(), float(tol), int(max_iter), flags, out.alpha.data_ptr(),
            out.contact.data_ptr() if out.contact is not None else None,
            out.grad.data_ptr() if out.grad is not None else None, out.iters.data_ptr(), out.status.data_ptr(), stream))
        return out

    def solve_records(self, plan: Plan, pose1, pose2, dest_ptrs, record_offset: int = 0, tol: float = 1e-6,
                      max_iter: int = 50, contact=None, fix_case4: bool = False, multicast: bool = False,
                      one_pair_per_thread: bool = False, lane_refill: bool = False):
        """Record mode: every pair's 112-byte record ``{alpha, grad[12], iters, status}`` is written, in plan
        order, to each of the raw device addresses ``dest_ptrs`` (local buffers or peer-GPU buffers mapped with
        CUDA IPC — the all-gather of the results fused into the solve).  Enqueues on the current stream."""
        import torch
        self._check_open(plan)
        B = plan.size
        for name, t in (("pose1", pose1), ("pose2", pose2)):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == (B, 6)):
                raise ValueError(f"{name} must be a contiguous float64 CUDA tensor of shape ({B}, 6)")
        arr = (C.c_void_p * len(dest_ptrs))(*[int(p) for p in dest_ptrs])
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().dcol_proximity_batch_records(
            plan._handle, pose1.data_ptr(), pose2.data_ptr(), float(tol), int(max_iter),
            (FIX_CASE4 if fix_case4 else 0) | (_lib.DEST_MULTICAST if multicast else 0)
            | (ONE_PAIR_PER_THREAD if one_pair_per_thread else 0) | (LANE_REFILL if lane_refill else 0), len(dest_ptrs), arr,
            int(record_offset), contact.data_ptr() if contact is not None else None, stream))

    # ------------------------------------------------------------------ host buffers
    def solve_host(self, idx1, idx2, pose1, pose2, tol: float = 1e-6, max_iter: int = 50, want_grad: bool = True,
                   want_contact: bool = True, out: BatchResult | None = None, fix_case4: bool = False,
                   grad1: bool = False, one_pair_per_thread: bool = False) -> BatchResult:
        """The reference-facing call: NumPy buffers in, NumPy buffers out, copies inside.  ``grad1``: ``grad`` is
        ``[B, 6]`` (d alpha / d[r1 p1], ``DCOL_WANT_GRAD1``), which halves the bytes that come back over PCIe."""
        idx1 = np.ascontiguousarray(idx1, dtype=np.int32)
        idx2 = np.ascontiguousarray(idx2, dtype=np.int32)
        pose1 = np.ascontiguousarray(pose1, dtype=np.float64).reshape(-1, 6)
        pose2 = np.ascontiguousarray(pose2, dtype=np.float64).reshape(-1, 6)
        B = int(idx1.shape[0])
        if not (idx2.shape[0] == B and pose1.shape[0] == B and pose2.shape[0] == B):
            raise ValueError("idx1, idx2, pose1, pose2 must describe the same number of pairs")
        if out is None:
            out = BatchResult(alpha=np.empty(B), contact=np.empty((B, 3)) if want_contact else None,
                              grad=np.empty((B, 6 if grad1 else 12)) if want_grad else None, iters=np.empty(B, np.int32),
                              status=np.empty(B, np.int32))
        if out.grad is not None and tuple(out.grad.shape) != (B, 6 if grad1 else 12):
            raise ValueError(f"out.grad must have shape ({B}, {6 if grad1 else 12})")
        flags = ((WANT_CONTACT if out.contact is not None else 0) | (WANT_GRAD if out.grad is not None else 0)
                 | (FIX_CASE4 if fix_case4 else 0) | (WANT_GRAD1 if (grad1 and out.grad is not None) else 0)
                 | (ONE_PAIR_PER_THREAD if one_pair_per_thread else 0))
        _lib.check(_lib.lib().dcol_proximity_batch_host(
            self._table, idx1.ctypes.data, idx2.ctypes.data, pose1.ctypes.data, pose2.ctypes.data, B, float(tol),
            int(max_iter), flags, out.alpha.ctypes.data, out.contact.ctypes.data if out.contact is not None else None,
            out.grad.ctypes.data if out.grad is not None else None, out.iters.ctypes.data, out.status.ctypes.data))
        return out

    def solve_scene_host(self, victim_shape: int, victim_poses, obstacle_shapes, obstacle_poses, tol: float = 1e-6,
                         max_iter: int = 50, want_grad: bool = True, want_iters: bool = True, out: "SceneResult | None" = None,
                         fix_case4: bool = False, one_pair_per_thread: bool = False, lane_refill: bool = False) -> "SceneResult":
        """Every (victim pose, obstacle) pair of a scene in one call (``dcol_proximity_scene_host``): ``victim_poses``
        ``[M, 6]`` of shape ``victim_shape`` against ``obstacle_shapes[j]`` at ``obstacle_poses[j]`` (``[n_obs, 6]``).  NumPy
        in, NumPy out: ``alpha[M, n_obs]``, ``grad1[M, n_obs, 6]`` = d alpha / d(victim r, p) — the part of the gradient the
        reference's systems keep (``cluttered_hallway_quadrotor.py:161-163``) — ``iters``, ``status``.  Only the poses
        cross PCIe on the way in; they are broadcast into pairs on the device."""
        vp = np.ascontiguousarray(victim_poses, dtype=np.float64).reshape(-1, 6)
        osh = np.ascontiguousarray(obstacle_shapes, dtype=np.int32).reshape(-1)
        op = np.ascontiguousarray(obstacle_poses, dtype=np.float64).reshape(-1, 6)
        M, n_obs = int(vp.shape[0]), int(osh.shape[0])
        if op.shape[0] != n_obs:
            raise ValueError("obstacle_shapes and obstacle_poses must describe the same number of obstacles")
        if out is None:
            out = SceneResult(alpha=np.empty((M, n_obs)), grad1=np.empty((M, n_obs, 6)) if want_grad else None,
                              iters=np.empty((M, n_obs), np.int32) if want_iters else None,
                              status=np.empty((M, n_obs), np.int32))
        if out.alpha.shape != (M, n_obs) or out.status.shape != (M, n_obs):
            raise ValueError(f"out.alpha / out.status must have shape ({M}, {n_obs})")
        flags = ((FIX_CASE4 if fix_case4 else 0) | (ONE_PAIR_PER_THREAD if one_pair_per_thread else 0)
                 | (LANE_REFILL if lane_refill else 0))
        _lib.check(_lib.lib().dcol_proximity_scene_host(
            self._table, int(victim_shape), vp.ctypes.data, M, osh.ctypes.data, op.ctypes.data, n_obs, float(tol),
            int(max_iter), flags, out.alpha.ctypes.data, out.grad1.ctypes.data if out.grad1 is not None else None,
            out.iters.ctypes.data if out.iters is not None else None, out.status.ctypes.data))
        return out

    # ------------------------------------------------------------------ debugging
    def trace_pair(self, i1: int, i2: int, pose1, pose2, tol: float = 1e-6) -> dict:
        """One pair with the per-iteration ``mu`` trace and the final world-frame ``(x, s, z)``."""
        pose1 = np.ascontiguousarray(pose1, dtype=np.float64).reshape(6)
        pose2 = np.ascontiguousarray(pose2, dtype=np.float64).reshape(6)
        alpha = C.c_double()
        x, s, z = np.full(_lib.MAX_N, np.nan), np.full(_lib.MAX_M, np.nan), np.full(_lib.MAX_M, np.nan)
        mu = np.full(_lib.MAX_ITER + 1, np.nan)
        n, m, iters, status = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(_lib.lib().dcol_debug_trace_pair(
            self._table, int(i1), int(i2), pose1.ctypes.data, pose2.ctypes.data, float(tol),
            C.cast(C.byref(alpha), C.c_void_p), x.ctypes.data, s.ctypes.data, z.ctypes.data,
            C.cast(C.byref(n), C.c_void_p), C.cast(C.byref(m), C.c_void_p), C.cast(C.byref(iters), C.c_void_p),
            C.cast(C.byref(status), C.c_void_p), mu.ctypes.data))
        return dict(alpha=alpha.value, x=x[:n.value], s=s[:m.value], z=z[:m.value], iters=iters.value,
                    status=status.value, mu=mu, n=n.value, m=m.value)


def pinned_empty(shape, dtype=np.float64):
    """A page-locked NumPy array (``dcol_host_alloc``): host buffers handed to ``solve_host`` copy
    asynchronously at PCIe rate when they live in such memory."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    ptr = C.c_void_p()
    _lib.check(_lib.lib().dcol_host_alloc(n, C.byref(ptr)))
    buf = (C.c_char * max(n, 1)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.ctypes.data] = ptr
    return arr


def pinned_free(arr):
    ptr = _PINNED.pop(arr.ctypes.data, None)
    if ptr is not None:
        _lib.lib().dcol_host_free(ptr)


_PINNED: dict = {}


def release_cached():
    """Hand the device scratch that destroyed engines / plans left in the library's cache back to the driver
    (``dcol_release_cached``)."""
    _lib.lib().dcol_release_cached()


def measure_fp64_peak(device: int = 0) -> float:
    """Measured FP64 FMA throughput of the device in FLOP/s (the roofline denominator)."""
    v = C.c_double()
    _lib.check(_lib.lib().dcol_measure_fp64_peak(int(device), C.byref(v)))
    return v.value


def proximity_batch(prims1, prims2, pdip_tol: float = 1e-6, want_grad: bool = True, device: int = 0) -> BatchResult:
    """Evaluate ``proximity_gradient(prims1[k], prims2[k])`` for all k in one call.

    ``prims1``/``prims2`` are equally long sequences of primitive objects (poses are read from
    ``.r`` / ``.p`` now).  Objects that are the same Python object share one shape record.
    Returns NumPy arrays; failures are reported in ``status`` (use :func:`raise_for_status`)."""
    if len(prims1) != len(prims2):
        raise ValueError("prims1 and prims2 must have the same length")
    uniq, index = [], {}
    idx = np.empty((2, len(prims1)), dtype=np.int32)
    poses = np.empty((2, len(prims1), 6))
    for side, prims in enumerate((prims1, prims2)):
        for k, prim in enumerate(prims):
            j = index.get(id(prim))
            if j is None:
                j = index[id(prim)] = len(uniq)
                uniq.append(prim)
            idx[side, k] = j
            poses[side, k] = pose_of(prim)
    eng = ProximityEngine(uniq, device=device)
    try:
        return eng.solve_host(idx[0], idx[1], poses[0], poses[1], tol=pdip_tol, want_grad=want_grad)
    finally:
        eng.close()

"""Multi-GPU plumbing: one process per GPU, the batch sharded by contiguous ranges, results
returned with ONE all-gather (BASELINE.json north_star).

Every (candidate x knot x obstacle) pair is an independent solve, so there is no exchange during
the solve.  The solve writes its outputs straight into one flat buffer per rank,

    [ alpha (B) | grad (12 B) | iters, status (B int32 each) ]   14 float64 words = 112 bytes per pair

so the gather is a single ``all_gather_into_tensor`` over NVLink with no packing pass (the contact
point, which no caller of the reference consumes across ranks, stays on the rank that computed it).
:class:`GatherPipeline` cuts a rank's batch into chunks and gathers chunk *c* on a side stream while
chunk *c+1* is being solved, so that only the last chunk's transfer is exposed.
Works with any ``torch.distributed`` backend (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import torch

from .engine import BatchResult

WORDS_PER_PAIR = 1 + 12 + 1   # float64 words of one pair's record (112 bytes)


def shard_bounds(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of ``n_items`` (first ``n_items % world`` ranks get one more)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def packed_views(flat: torch.Tensor, B: int) -> BatchResult:
    """Views of one rank's flat record buffer (``WORDS_PER_PAIR * B`` float64 words)."""
    if flat.dtype != torch.float64 or flat.numel() != WORDS_PER_PAIR * B or not flat.is_contiguous():
        raise ValueError("flat must be a contiguous float64 tensor of WORDS_PER_PAIR * B words")
    alpha = flat[:B]
    grad = flat[B:13 * B].view(B, 12)
    ints = flat[13 * B:14 * B].view(torch.int32)          # 2 B int32 words
    return BatchResult(alpha=alpha, contact=None, grad=grad, iters=ints[:B], status=ints[B:])


def alloc_packed(B: int, device, with_contact: bool = False) -> tuple[torch.Tensor, BatchResult]:
    """One rank's record buffer and its views; ``with_contact`` adds a rank-local ``[B, 3]`` contact buffer."""
    flat = torch.empty(WORDS_PER_PAIR * B, dtype=torch.float64, device=device)
    views = packed_views(flat, B)
    if with_contact:
        views.contact = torch.empty((B, 3), dtype=torch.float64, device=device)
    return flat, views


def all_gather_packed(flat: torch.Tensor, B: int, world: int, out: torch.Tensor | None = None, group=None,
                      async_op: bool = False):
    """Gather every rank's flat record buffer (equal ``B`` on all ranks).  Returns ``(gathered, work)``
    with ``gathered`` of shape ``[world, WORDS_PER_PAIR * B]``; ``split_gathered`` gives per-rank views."""
    import torch.distributed as dist
    if out is None:
        out = torch.empty((world, flat.numel()), dtype=flat.dtype, device=flat.device)
    if flat.is_cuda:
        work = dist.all_gather_into_tensor(out.view(-1), flat, group=group, async_op=async_op)
    else:   # gloo has no all_gather_into_tensor for CPU tensors on every version: use the list form
        work = dist.all_gather(list(out.unbind(0)), flat, group=group, async_op=async_op)
    return out, work


def split_gathered(gathered: torch.Tensor, B: int) -> list[BatchResult]:
    return [packed_views(gathered[r], B) for r in range(gathered.shape[0])]


def sharded_solve(solve_local, n_pairs: int, rank: int, world: int, device, group=None):
    """Solve ``n_pairs`` pairs split over ``world`` ranks and return every rank's results.

    ``solve_local(lo, hi, out: BatchResult)`` must fill ``out`` for global pairs ``[lo, hi)``.  Ranks
    are padded to the largest shard so one fixed-size all-gather suffices; returns a list of
    ``(lo, hi, BatchResult)`` in rank order (views into the gathered buffer)."""
    bounds = [shard_bounds(n_pairs, r, world) for r in range(world)]
    Bmax = max(hi - lo for lo, hi in bounds)
    flat, views = alloc_packed(Bmax, device)
    flat.zero_()
    lo, hi = bounds[rank]
    n = hi - lo
    local = BatchResult(alpha=views.alpha[:n], contact=None, grad=views.grad[:n],
                        iters=views.iters[:n], status=views.status[:n])
    solve_local(lo, hi, local)
    if world == 1:
        return [(lo, hi, local)]
    gathered, _ = all_gather_packed(flat, Bmax, world, group=group)
    outs = []
    for r, (rlo, rhi) in enumerate(bounds):
        v = packed_views(gathered[r], Bmax)
        k = rhi - rlo
        outs.append((rlo, rhi, BatchResult(alpha=v.alpha[:k], contact=None, grad=v.grad[:k],
                                           iters=v.iters[:k], status=v.status[:k])))
    return outs


class GatherPipeline:
    """Chunked solve with the all-gather of chunk c overlapped with the solve of chunk c+1.

    ``bounds``: the chunk boundaries of this rank's batch (equal on all ranks).  ``launch(c, out)`` must
    enqueue the solve of chunk ``c`` on the current stream, writing ``out`` (views of the chunk's record
    buffer).  ``step()`` enqueues one whole pass; it returns with the main stream waiting on the last
    gather, so the caller's event/synchronize brackets see the complete step."""

    def __init__(self, bounds, world: int, device, with_contact: bool = True, group=None):
        self.bounds, self.world, self.group = list(bounds), world, group
        self.flat, self.out, self.gathered = [], [], []
        for lo, hi in self.bounds:
            f, v = alloc_packed(hi - lo, device, with_contact=with_contact)
            self.flat.append(f)
            self.out.append(v)
            self.gathered.append(torch.empty((world, f.numel()), dtype=torch.float64, device=device)
                                 if world > 1 else None)
        self.comm = torch.cuda.Stream(device=device) if world > 1 else None
        self.events = [torch.cuda.Event() for _ in self.bounds] if world > 1 else []

    def step(self, launch):
        main = torch.cuda.current_stream()
        for c, (lo, hi) in enumerate(self.bounds):
            launch(c, self.out[c])
            if self.world > 1:
                self.events[c].record(main)
                self.comm.wait_event(self.events[c])
                with torch.cuda.stream(self.comm):
                    all_gather_packed(self.flat[c], hi - lo, self.world, out=self.gathered[c], group=self.group)
        if self.world > 1:
            main.wait_stream(self.comm)

    def rank_results(self, rank: int) -> list[BatchResult]:
        """Per-chunk views of ``rank``'s records after a step (the local buffers if ``world == 1``)."""
        if self.world == 1:
            return self.out
        return [packed_views(g[rank], hi - lo) for g, (lo, hi) in zip(self.gathered, self.bounds)]


class _RecordGather:
    """Shared behaviour of the two fused-gather fabrics.

    DOUBLE BUFFERED: the gathered allocation holds two halves ``[2, world, B, 14]`` and consecutive steps alternate
    between them.  With a single buffer, rank A's kernel of step n+1 (free to start as soon as A's own handshake n
    has completed) could store into slot A of rank B's buffer while B is still reading step n.  With two halves the
    buffer that step n+2 overwrites is the one step n was read from, and A cannot start step n+2 before handshake
    n+1 has completed on A, which needs B's stream to have reached its handshake n+1 — stream-ordered after B's reads
    of step n.  So: results of a step stay valid until the step after next is enqueued, provided every consumer reads
    them on the stream the steps are enqueued on (as :class:`FusedShardedSolver` does), and no extra synchronisation
    is needed.

    Use: ``ptrs = g.begin_step()`` -> ``engine.solve_records(..., ptrs, multicast=g.multicast)`` -> ``g.handshake()``
    -> read ``g.gathered`` (``[world, B, 14]`` float64, plan order per source rank)."""

    multicast = False

    def _init_common(self, B, rank, world, local_device, group):
        self.B, self.rank, self.world, self.dev, self.group = B, rank, world, local_device, group
        self._half = 1          # begin_step() flips first: step 0 uses half 0
        self._flag = torch.zeros(1, dtype=torch.float32, device=torch.device("cuda", local_device))

    @property
    def half_words(self) -> int:
        from . import _lib
        return self.world * max(self.B, 1) * _lib.RECORD_WORDS

    def begin_step(self):
        """Switch to the other half and return the destination addresses the solve of this step must write to."""
        self._half ^= 1
        return self.dest_ptrs

    @property
    def dest_ptrs(self):
        from . import _lib
        slot = (self._half * self.half_words + self.rank * max(self.B, 1) * _lib.RECORD_WORDS) * 8
        return [base + slot for base in self._dest_bases]

    @property
    def gathered(self):
        """``[world, B, 14]`` view of the half the current step writes / wrote."""
        return self._both[self._half][:, :self.B]

    def handshake(self):
        """Stream-ordered completion barrier (4-byte all-reduce): when it has completed on this rank's stream every
        peer's kernel of this step, and with it every peer's stores into this rank's buffer, has completed."""
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self._flag, group=self.group)


class PeerRecordGather(_RecordGather):
    """The all-gather fused into the solve: every rank's kernel writes its 112-byte records straight into
    slot ``rank`` of EVERY rank's gathered buffer, locally and over NVLink peer mappings (CUDA IPC), from
    the kernel's epilogue.  Nothing is packed, copied or sent afterwards; a scalar all-reduce on the same
    stream is the completion handshake.  All ranks must use the same ``B``.  See :class:`_RecordGather`."""

    multicast = False

    def __init__(self, B: int, rank: int, world: int, local_device: int, group=None):
        import ctypes as C
        import torch.distributed as dist
        from . import _lib
        from .engine import device_view
        if world > _lib.MAX_DEST:
            raise ValueError(f"at most {_lib.MAX_DEST} ranks per node")
        L = _lib.lib()
        self._init_common(B, rank, world, local_device, group)
        nbytes = 2 * self.half_words * 8
        ptr = C.c_void_p()
        _lib.check(L.dcol_device_alloc(local_device, nbytes, C.byref(ptr)))
        self._ptr = ptr.value
        self._both = device_view(self._ptr, (2, world, max(B, 1), _lib.RECORD_WORDS), torch.device("cuda", local_device))
        handle = (C.c_ubyte * 64)()
        self._peers = {}
        if world > 1:
            _lib.check(L.dcol_ipc_export(local_device, self._ptr, C.cast(handle, C.c_void_p)))
            mine = (local_device, bytes(handle))
            everyone = [None] * world
            dist.all_gather_object(everyone, mine, group=group)
            for r, (_, h) in enumerate(everyone):
                if r == rank:
                    continue
                p = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                _lib.check(L.dcol_ipc_import(local_device, C.cast(buf, C.c_void_p), C.byref(p)))
                self._peers[r] = p.value
        self._dest_bases = [(self._ptr if r == rank else self._peers[r]) for r in range(world)]

    def close(self):
        from . import _lib
        L = _lib.lib()
        for p in self._peers.values():
            L.dcol_ipc_close(self.dev, p)
        self._peers = {}
        if self._ptr:
            L.dcol_device_free(self.dev, self._ptr)
            self._ptr = None


class MulticastRecordGather(_RecordGather):
    """The fused all-gather over NVLink SHARP: the gathered buffers of all ranks are one symmetric allocation
    (``torch.distributed._symmetric_memory``) with a MULTICAST address, and the solve kernel's epilogue issues one
    ``multimem.st`` per 16 bytes of a record; the NVSwitch replicates it into slot ``rank`` of every rank's
    buffer.  Egress per GPU is 112 B/pair, independent of the world size.  Same interface as
    :class:`PeerRecordGather`; raises if the fabric has no multicast support (callers then fall back to it)."""

    multicast = True

    def __init__(self, B: int, rank: int, world: int, local_device: int, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._init_common(B, rank, world, local_device, group)
        dev = torch.device("cuda", local_device)
        flat = symm_mem.empty(2 * self.half_words, dtype=torch.float64, device=dev)
        self._hdl = symm_mem.rendezvous(flat, group if group is not None else dist.group.WORLD)
        mc = int(getattr(self._hdl, "multicast_ptr", 0) or 0)
        if mc == 0:
            raise RuntimeError("symmetric memory has no multicast address on this fabric")
        self._flat = flat
        self._both = flat.view(2, world, max(B, 1), _lib.RECORD_WORDS)
        self._dest_bases = [mc]

    def close(self):
        self._hdl = None
        self._flat = None
        self._both = None


class FusedShardedSolver:
    """Multi-GPU entry point: every rank passes the SAME global batch description, solves its contiguous shard and
    ends up with the results of ALL pairs (pair order), the all-gather being fused into the solve kernel.

    Set up once per (engine, idx1, idx2): shard bounds, the local plan, the gathered buffers (NVLink multicast if the
    fabric has it, else unicast CUDA-IPC peer stores) and every rank's plan permutation (exchanged once, it is static).
    ``solve(pose1, pose2)`` then takes the global ``[B, 6]`` pose tensors (CUDA, this rank's device; only the local
    shard is read) and returns a :class:`BatchResult` over all ``B`` pairs."""

    def __init__(self, engine, idx1, idx2, rank: int, world: int, group=None, fabric: str = "auto"):
        import torch.distributed as dist
        self.engine, self.rank, self.world, self.group = engine, rank, world, group
        idx1, idx2 = torch.as_tensor(idx1), torch.as_tensor(idx2)
        self.B = int(idx1.shape[0])
        if self.B < world:      # the same on every rank, so every rank raises (no rank is left waiting in a collective)
            raise ValueError(f"FusedShardedSolver needs at least one pair per rank (B = {self.B}, world = {world})")
        self.bounds = [shard_bounds(self.B, r, world) for r in range(world)]
        self.Bmax = max(hi - lo for lo, hi in self.bounds)
        lo, hi = self.bounds[rank]
        self.lo, self.n = lo, hi - lo
        # every rank's shard is padded to Bmax pairs by repeating its last pair, so that all slots have one size
        sel = torch.arange(lo, lo + self.Bmax).clamp(max=max(hi - 1, lo))
        self._sel = sel.to(engine.device)
        self.plan = engine.plan(idx1[sel], idx2[sel])
        dev = engine.device.index
        self.gather = None
        if world > 1:
            def agree(ok):
                flag = torch.tensor([1.0 if ok else 0.0], device=engine.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                return float(flag) == 1.0
            if fabric in ("auto", "multicast"):
                try:
                    self.gather = MulticastRecordGather(self.Bmax, rank, world, dev, group)
                except Exception:
                    self.gather = None
                if not agree(self.gather is not None):
                    self.gather = None
            if self.gather is None:
                self.gather = PeerRecordGather(self.Bmax, rank, world, dev, group)
            perms = [torch.empty(self.Bmax, dtype=torch.int32, device=engine.device) for _ in range(world)]
            dist.all_gather(perms, self.plan.perm(), group=group)
        else:
            self._local = torch.empty((1, self.Bmax, WORDS_PER_PAIR), dtype=torch.float64, device=engine.device)
            perms = [self.plan.perm()]
        # scatter map: record i of rank r belongs to global pair bounds[r].lo + perm_r[i] (padding rows dropped)
        dst, src = [], []
        for r, (rlo, rhi) in enumerate(self.bounds):
            pr = perms[r].long()
            keep = pr < (rhi - rlo)
            dst.append(rlo + pr[keep])
            src.append(r * self.Bmax + torch.nonzero(keep).squeeze(1))
        self._dst, self._src = torch.cat(dst), torch.cat(src)

    def solve(self, pose1, pose2, tol: float = 1e-6, max_iter: int = 50) -> BatchResult:
        from .engine import records_to_result
        p1 = pose1.index_select(0, self._sel).contiguous()
        p2 = pose2.index_select(0, self._sel).contiguous()
        if self.world > 1:
            # double-buffered gathered slots: see _RecordGather (back-to-back solves may overlap across ranks)
            self.engine.solve_records(self.plan, p1, p2, self.gather.begin_step(), tol=tol, max_iter=max_iter,
                                      multicast=self.gather.multicast)
            self.gather.handshake()
            rec = self.gather.gathered.reshape(self.world * self.Bmax, WORDS_PER_PAIR)
        else:
            self.engine.solve_records(self.plan, p1, p2, [self._local.data_ptr()], tol=tol, max_iter=max_iter)
            rec = self._local.reshape(self.Bmax, WORDS_PER_PAIR)
        out = torch.empty((self.B, WORDS_PER_PAIR), dtype=torch.float64, device=rec.device)
        out[self._dst] = rec[self._src]
        return records_to_result(out)

    def close(self):
        if self.gather is not None:
            self.gather.close()
        self.plan.close()

"""Table-wise sharded embeddings across the GPUs of one box (new functionality: DLRM.jl is
single-process; BASELINE.json's north_star asks for table-wise model parallelism with the dense
MLPs and the interaction data-parallel).

One process per GPU.  Every table lives whole on one rank.  Per step:

  forward   each rank sends the index columns of its local batch to the tables' owners
            (all-to-all, a few hundred KB), owners pool their tables for the GLOBAL batch in one
            lookup launch, and a second all-to-all delivers to every rank the pooled rows of its
            local samples, which land in the interaction input T [B_local][1 + ntab][D];
  backward  the mirror all-to-all returns dT slices to the owners, which run the fused sparse SGD
            locally over the global batch; dense gradients are summed with one all-reduce.

The routing logic here is pure tensor plumbing over ``torch.distributed`` (NCCL on GPUs, gloo in
the CPU tests); the compute is injected (`lookup_fn`, `update_fn`) and in the product is always the
CUDA library.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


@dataclass
class TableSharding:
    """Which rank owns which table, balanced by LOOKUP COUNT first (every table receives
    B_global * P lookups per step whatever its size), then by bytes: tables are dealt in
    descending row count, snake order, so the big tables land on different ranks."""
    rows: List[int]
    world: int
    owner: List[int]
    local: List[List[int]]      # local[r] = global table ids owned by rank r, ascending
    _slot_cache: Optional[dict] = None

    def slot_index(self, r: int, device) -> torch.Tensor:
        """Device tensor of the interaction slots (1 + table id) of rank r's tables (cached: no
        host-to-device copy inside the step, which also keeps the step CUDA-graph capturable)."""
        if self._slot_cache is None:
            self._slot_cache = {}
        key = (r, str(device))
        if key not in self._slot_cache:
            self._slot_cache[key] = torch.tensor([1 + k for k in self.local[r]], dtype=torch.int64, device=device)
        return self._slot_cache[key]

    def owner_order(self, device) -> torch.Tensor:
        """Table ids grouped by owner rank (the send order of the index all-to-all), cached."""
        if self._slot_cache is None:
            self._slot_cache = {}
        key = ("order", str(device))
        if key not in self._slot_cache:
            order = [k for r in range(self.world) for k in self.local[r]]
            self._slot_cache[key] = torch.tensor(order, dtype=torch.int64, device=device)
        return self._slot_cache[key]

    @classmethod
    def build(cls, rows: Sequence[int], world: int) -> "TableSharding":
        rows = [int(r) for r in rows]
        order = sorted(range(len(rows)), key=lambda k: (-rows[k], k))
        owner = [0] * len(rows)
        for i, k in enumerate(order):
            rnd, pos = divmod(i, world)
            owner[k] = pos if rnd % 2 == 0 else world - 1 - pos
        local = [[k for k in range(len(rows)) if owner[k] == r] for r in range(world)]
        return cls(rows, world, owner, local)

    def counts(self) -> List[int]:
        return [len(l) for l in self.local]

    def bytes_per_rank(self, D: int) -> List[int]:
        return [sum(self.rows[k] for k in l) * D * 4 for l in self.local]


def _a2a(out: torch.Tensor, inp: torch.Tensor, out_splits, in_splits, group) -> None:
    dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)


def exchange_indices(idx_local: torch.Tensor, sh: TableSharding, rank: int, group=None) -> torch.Tensor:
    """idx_local [ntab][B_local][P] (this rank's samples, all tables) ->
    [t_mine][B_global][P] (all samples, this rank's tables; sample order = rank-major)."""
    ntab, Bl, P = idx_local.shape
    W = sh.world
    if W == 1:
        return idx_local
    send = idx_local.index_select(0, sh.owner_order(idx_local.device))                 # grouped by owner
    t_mine = len(sh.local[rank])
    recv = torch.empty((W, t_mine, Bl, P), dtype=idx_local.dtype, device=idx_local.device)
    in_splits = [len(sh.local[r]) * Bl * P for r in range(W)]
    out_splits = [t_mine * Bl * P] * W
    _a2a(recv.view(-1), send.view(-1), out_splits, in_splits, group)
    # [W][t][Bl][P] -> [t][W*Bl][P]
    return recv.permute(1, 0, 2, 3).reshape(t_mine, W * Bl, P).contiguous()


def exchange_pooled(pooled: torch.Tensor, T: torch.Tensor, sh: TableSharding, rank: int, group=None) -> None:
    """pooled [B_global][t_mine][D] (owner side) -> T[:, 1 + k, :] for every table k, on the rank
    that owns the samples.  T is [B_local][1 + ntab][D]; slot 0 is left for x."""
    W = sh.world
    Bg, t_mine, D = pooled.shape
    Bl = Bg // W
    if W == 1:
        T[:, 1:, :] = pooled
        return
    counts = sh.counts()
    recv = torch.empty((Bl * sum(counts) * D,), dtype=pooled.dtype, device=pooled.device)
    in_splits = [Bl * t_mine * D] * W
    out_splits = [Bl * counts[r] * D for r in range(W)]
    _a2a(recv, pooled.reshape(-1), out_splits, in_splits, group)
    off = 0
    for r in range(W):
        if counts[r] == 0:
            continue
        chunk = recv[off:off + Bl * counts[r] * D].view(Bl, counts[r], D)
        T.index_copy_(1, sh.slot_index(r, T.device), chunk)
        off += Bl * counts[r] * D


def exchange_grads(dT: torch.Tensor, sh: TableSharding, rank: int, group=None) -> torch.Tensor:
    """dT [B_local][1 + ntab][D] -> [B_global][t_mine][D] on each owner (mirror of exchange_pooled)."""
    W = sh.world
    Bl, S, D = dT.shape
    if W == 1:
        return dT[:, 1:, :].contiguous()
    counts = sh.counts()
    t_mine = counts[rank]
    send = torch.cat([dT.index_select(1, sh.slot_index(r, dT.device)).reshape(-1) for r in range(W)])
    recv = torch.empty((W * Bl, t_mine, D), dtype=dT.dtype, device=dT.device)
    in_splits = [Bl * counts[r] * D for r in range(W)]
    out_splits = [Bl * t_mine * D] * W
    _a2a(recv.view(-1), send, out_splits, in_splits, group)
    return recv


class PeerExchange:
    """The interaction input buffer T [B_local][1 + ntab][D] of every rank, allocated by the
    library and mapped into every other rank's address space through CUDA IPC, so the owner of a
    table can store pooled rows straight into the buffer of the rank that owns the sample
    (`EmbeddingTables.lookup_p2p`).  `barrier` is the stream-ordered collective that orders the
    peer stores against the consumers (a one-element NCCL all-reduce by default)."""

    def __init__(self, B_local: int, ntab: int, D: int, rank: int, world: int, device, group=None,
                 barrier: Optional[Callable] = None, flag_barrier: bool = True):
        import ctypes as C

        from . import _lib
        from .embedding import _DevicePtrView
        self.lib = _lib.load()
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.device(device)
        self.shape = (B_local, 1 + ntab, D)
        nbytes = B_local * (1 + ntab) * D * 4
        h = C.c_void_p()
        _lib.check(self.lib.dlrmb_xbuf_create(self.device.index or 0, nbytes, C.byref(h)))
        self._xbuf = h
        p = C.c_void_p()
        _lib.check(self.lib.dlrmb_xbuf_ptr(h, C.byref(p)))
        handle = (C.c_uint8 * 64)()
        _lib.check(self.lib.dlrmb_xbuf_ipc_handle(h, handle))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.peer_ptrs = []
        self._owned, self._mapped = [], [self.peer_ptrs]      # exchange buffers / peer mappings to release in close()
        for r in range(world):
            if r == rank:
                self.peer_ptrs.append(p.value)
                continue
            buf = (C.c_uint8 * 64).from_buffer_copy(handles[r])
            q = C.c_void_p()
            _lib.check(self.lib.dlrmb_xbuf_open(self.device.index or 0, buf, C.byref(q)))
            self.peer_ptrs.append(q.value)
        self.T = torch.as_tensor(_DevicePtrView(p.value, self.shape, self), device=self.device)
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._barrier = barrier
        self.G = None
        self.peer_G_ptrs = None
        self._gbuf = None
        self._closed = False
        self._fbuf = None
        self.peer_flag_ptrs = None
        self._ibuf = None
        self.idx_owned = None
        self._idx_dests = None
        if barrier is None and flag_barrier:
            # ordering without NCCL: a flag array per rank, mapped by every peer (dlrmb_peer_barrier)
            self._fbuf, _own, self.peer_flag_ptrs = self._share(int(self.lib.dlrmb_peer_barrier_flag_bytes()))
            self._bstate = torch.zeros(int(self.lib.dlrmb_peer_barrier_state_bytes()) // 4, dtype=torch.int32, device=self.device)
            self._flag_arr = (C.c_void_p * world)(*[int(q) for q in self.peer_flag_ptrs])
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)           # every rank's flag array exists and is zero before anyone signals

    def enable_index_exchange(self, sharding: "TableSharding", B_local: int, P: int, idx_bytes: int = 4):
        """Index exchange by peer stores: an index buffer [t_mine][B_global][P] on every rank, mapped
        everywhere; each rank stores its samples' index columns straight into the owners' buffers."""
        from .embedding import _DevicePtrView
        counts = sharding.counts()
        t_mine = counts[self.rank]
        Bg = B_local * self.world
        per_table = Bg * P * idx_bytes
        self._ibuf, own, peer_idx = self._share(max(1, t_mine) * per_table)
        dt, ts = (torch.int32, "<i4") if idx_bytes == 4 else (torch.int64, "<i8")
        self.idx_owned = torch.as_tensor(_DevicePtrView(own, (max(1, t_mine), Bg, P), self, ts), device=self.device)
        assert self.idx_owned.dtype == dt
        dests = [peer_idx[sharding.owner[k]] + sharding.local[sharding.owner[k]].index(k) * per_table
                 for k in range(len(sharding.rows))]
        self._idx_dests = torch.tensor(dests, dtype=torch.int64, device=self.device)
        self._idx_geom = (len(sharding.rows), B_local, P, idx_bytes)

    def enable_small_allreduce(self, n: int) -> None:
        """Exchange buffer [2][world][n] floats for `allreduce_small` (dlrmb_peer_allreduce_f32)."""
        import ctypes as C
        assert self.peer_flag_ptrs is not None, "the one-shot all-reduce is ordered by the flag barrier"
        self._ar_n = int(n)
        self._arbuf, _own, ptrs = self._share(2 * self.world * self._ar_n * 4)
        self._ar_arr = (C.c_void_p * self.world)(*[int(q) for q in ptrs])

    def allreduce_small(self, t: torch.Tensor) -> None:
        """In-place sum of `t` (contiguous f32, numel == the enabled n, a multiple of 4) over all ranks: every
        rank pushes its copy into every rank's buffer over NVLink, flag barrier, rank-ordered local sum."""
        from . import _lib, _prof
        assert t.is_contiguous() and t.dtype == torch.float32 and t.numel() == self._ar_n
        with _prof.range("peer_allreduce"):
            _lib.check(self.lib.dlrmb_peer_allreduce_f32(
                self.device.index or 0, self._ar_arr, self._flag_arr, self.world, self.rank, 3, self._bstate.data_ptr(),
                t.data_ptr(), t.numel(), int(torch.cuda.current_stream(self.device).cuda_stream)))

    def scatter_indices(self, idx_local: torch.Tensor) -> torch.Tensor:
        """idx_local [ntab][B_local][P] -> the owners' buffers (peer stores) + barrier; returns this rank's
        idx_owned [t_mine][B_global][P]."""
        from . import _lib
        ntab, Bl, P, ib = self._idx_geom
        assert tuple(idx_local.shape) == (ntab, Bl, P) and idx_local.element_size() == ib and idx_local.is_contiguous()
        from . import _prof
        with _prof.range("indices_scatter"):
            _lib.check(self.lib.dlrmb_indices_scatter_p2p(
                self.device.index or 0, idx_local.data_ptr(), ib, ntab, Bl, P, self._idx_dests.data_ptr(), self.rank,
                int(torch.cuda.current_stream(self.device).cuda_stream)))
        self.barrier(0)
        return self.idx_owned

    def close(self) -> None:
        """Unmap the peers' buffers (cudaIpcCloseMemHandle) and free this rank's exchange buffers.  Every
        rank must have finished using the mappings (call after a barrier); views of T / G are dead after."""
        if getattr(self, "_closed", True):
            return
        self._closed = True
        dev = self.device.index or 0
        for ptrs in self._mapped:
            for r, q in enumerate(ptrs):
                if r != self.rank and q:
                    self.lib.dlrmb_xbuf_close(dev, q)
        self.T = self.G = self.idx_owned = None
        for h in [self._xbuf] + self._owned:
            if h is not None:
                self.lib.dlrmb_xbuf_destroy(h)
        self._xbuf = self._gbuf = self._fbuf = self._ibuf = None
        self._owned, self._mapped = [], []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _share(self, nbytes: int):
        """Allocate an exchange buffer and map every rank's copy: returns (own ptr, [ptr per rank])."""
        import ctypes as C

        from . import _lib
        h = C.c_void_p()
        _lib.check(self.lib.dlrmb_xbuf_create(self.device.index or 0, max(256, nbytes), C.byref(h)))
        p = C.c_void_p()
        _lib.check(self.lib.dlrmb_xbuf_ptr(h, C.byref(p)))
        handle = (C.c_uint8 * 64)()
        _lib.check(self.lib.dlrmb_xbuf_ipc_handle(h, handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        ptrs = []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(p.value)
                continue
            q = C.c_void_p()
            _lib.check(self.lib.dlrmb_xbuf_open(self.device.index or 0, (C.c_uint8 * 64).from_buffer_copy(handles[r]), C.byref(q)))
            ptrs.append(q.value)
        self._owned.append(h)
        self._mapped.append(ptrs)
        return h, p.value, ptrs

    def enable_gradient_exchange(self, sharding: "TableSharding", B_local: int, D: int, split_dx: bool = False):
        """Gradient buffer G [B_global][tables of this rank][D] on every rank, mapped everywhere, and
        the per-slot destination table the scattered interaction backward reads."""
        from .embedding import _DevicePtrView
        from .interact import ScatterPlan
        counts = sharding.counts()
        t_mine = counts[self.rank]
        Bg = B_local * self.world
        self._gbuf, own, self.peer_G_ptrs = self._share(Bg * max(1, t_mine) * D * 4)
        self.G = torch.as_tensor(_DevicePtrView(own, (Bg, max(1, t_mine), D), self), device=self.device)
        rows = [[0, 0, 0]]                                    # slot 0 (x) has no destination
        for k in range(len(sharding.rows)):
            o = sharding.owner[k]
            rows.append([self.peer_G_ptrs[o], counts[o] * D, sharding.local[o].index(k) * D])
        dests = torch.tensor(rows, dtype=torch.int64, device=self.device)
        return ScatterPlan(dests, self.rank * B_local, torch.cuda.Stream(self.device) if split_dx else None)

    def barrier(self, channel: int = 0) -> None:
        """Order this rank's earlier peer stores before every rank's later reads (stream-ordered).
        Flag barrier over the mapped buffers (one tiny kernel, dlrmb_peer_barrier) when available; a
        caller-supplied callable (tests: host-side barrier) or a one-element NCCL all-reduce otherwise."""
        if self._barrier is not None:
            self._barrier()
        elif self.peer_flag_ptrs is not None:
            from . import _lib, _prof
            with _prof.range(f"peer_barrier_{channel}"):
                _lib.check(self.lib.dlrmb_peer_barrier(
                    self.device.index or 0, self._flag_arr, self.world, self.rank, channel, self._bstate.data_ptr(),
                    int(torch.cuda.current_stream(self.device).cuda_stream)))
        else:
            dist.all_reduce(self._flag, group=self.group)

    def barrier_timeouts(self) -> int:
        """Number of flag-barrier waits that gave up (a peer died); 0 in a healthy run."""
        if self.peer_flag_ptrs is None:
            return 0
        return int(self._bstate[-1].item())


class _ShardedLookupFn(torch.autograd.Function):
    """lookup (owner side) + pooled all-to-all forward; gradient all-to-all backward.  The
    owner-side gradient [B_global][t_mine][D] is parked on ``ctx_obj.owned_grad`` for the sparse
    update (the tables are not torch parameters, as in the reference where the lookup pullback
    yields SparseEmbeddingUpdate objects rather than dense gradients)."""

    @staticmethod
    def forward(ctx, anchor: torch.Tensor, se: "ShardedEmbedding", idx_local: torch.Tensor):
        ctx.se = se
        D = se.D
        Bl = idx_local.shape[1]
        if se.world == 1:
            # no exchange: pool straight into the interaction input (slot 0 is filled with x by
            # the interaction kernel's fused fast_vcat); the same launch sorts the indices for the update
            se.idx_owned = idx_local
            T = torch.empty((Bl, 1 + se.ntab, D), dtype=torch.float32, device=idx_local.device)
            se.lookup_fn(idx_local, T, 1)
            se.presorted = se.lookup_sorts
            return T
        idx_owned = se._exchange_indices(idx_local)
        se.idx_owned = idx_owned
        se.presorted = False
        Bg = idx_owned.shape[1]
        if se.peer is not None:
            # fused path: pooled rows go straight into the owners' T over NVLink; the barrier
            # orders every rank's stores before anyone reads its own T
            if len(se.local_ids):
                fuse = idx_owned.shape[1] * idx_owned.shape[2] <= se.tables.FUSED_SORT_MAX
                se.tables.lookup_p2p(idx_owned, se.peer.peer_ptrs, Bl, 1 + se.ntab, se.idx_base, sort=fuse)
                se.presorted = fuse
            se.peer.barrier(1)
            # a fresh alias every step: the persistent buffer tensor itself must never pick up
            # autograd history (a stale grad_fn from an earlier step would chain the graphs)
            return se.peer.T.detach()
        pooled = torch.empty((Bg, len(se.local_ids), D), dtype=torch.float32, device=idx_local.device)
        if len(se.local_ids):
            se.lookup_fn(idx_owned, pooled, 0)
        # slot 0 needs no clearing: the interaction kernel writes x there (fused fast_vcat)
        T = torch.empty((Bl, 1 + se.ntab, D), dtype=torch.float32, device=idx_local.device)
        exchange_pooled(pooled, T, se.sharding, se.rank, se.group)
        return T

    @staticmethod
    def backward(ctx, dT: torch.Tensor):
        se = ctx.se
        if se.world == 1:
            se.owned_grad = dT.contiguous()      # [B][1 + ntab][D], consumed with slot0 = 1
        else:
            se.owned_grad = exchange_grads(dT.contiguous(), se.sharding, se.rank, se.group)
        se._update_from_backward(fused=False)
        return None, None, None


class ShardedEmbedding:
    """Table-wise sharded drop-in for ``maplookup`` + ``update!`` at world size W.

    ``lookup_fn(idx_owned [t][Bg][P], out [Bg][slot0 + t][D], slot0)`` and
    ``update_fn(idx_owned, grad [Bg][slot0 + t][D], lr, slot0, presorted)`` do the local compute;
    :meth:`create` wires them to an :class:`~dlrm_jl_b200.embedding.EmbeddingTables` holding this
    rank's tables.  ``slot0`` is 1 at world size 1 (the buffers ARE the interaction input) else 0.
    """

    def __init__(self, rows: Sequence[int], D: int, rank: int, world: int, lookup_fn: Callable,
                 update_fn: Callable, group=None, sort_fn: Optional[Callable] = None, idx_base: int = 0,
                 lookup_sorts: bool = False):
        """``idx_base``: 1 for reference-format (1-based) ids, 0 for PyTorch-style ids; it is handed to
        every kernel that reads the indices.  ``lookup_sorts``: `lookup_fn` also prepares the sort /
        dedup of the step's update (the fused launch), so `sort_async` has nothing left to do."""
        self.sharding = TableSharding.build(rows, world)
        self.rank, self.world, self.group = rank, world, group
        self.idx_base = int(idx_base)
        self.lookup_sorts = bool(lookup_sorts)
        self.presorted = False
        self.D = D
        self.ntab = len(rows)
        self.local_ids = self.sharding.local[rank]
        self.lookup_fn, self.update_fn, self.sort_fn = lookup_fn, update_fn, sort_fn
        self.idx_owned: Optional[torch.Tensor] = None
        self.owned_grad: Optional[torch.Tensor] = None
        self.slot0 = 1 if world == 1 else 0
        self.peer: Optional[PeerExchange] = None
        self.scatter_plan = None
        self._auto_update = None

    def update_inside_backward(self, lr: float, stream: "torch.cuda.Stream") -> None:
        """Launch the sparse update from INSIDE the backward pass, on ``stream``, as soon as the pooled-embedding
        gradient exists (the lookup's pullback, or the scattering interaction backward on the fused path)
        instead of after ``loss.backward()`` has returned.  The update touches the tables only, so it then runs
        beside the bottom MLP's backward (autograd joins every backward stream before ``backward()`` returns, so
        a call placed after it starts a whole bottom-MLP backward later).  The caller joins ``stream`` at the
        end of the step; do not call :meth:`update` / :meth:`finish_backward` yourself in this mode."""
        self._auto_update = (float(lr), stream)
        if self.scatter_plan is not None:
            self.scatter_plan.on_launched = lambda: self._update_from_backward(fused=True)

    def _update_from_backward(self, fused: bool) -> None:
        if self._auto_update is None:
            return
        lr, stream = self._auto_update
        cur = torch.cuda.current_stream(self.tables.device)
        stream.wait_stream(cur)
        with torch.cuda.stream(stream):
            if fused:
                self.finish_backward()
            self.update(lr)

    def enable_fused_backward(self, B_local: int, split_dx: bool = False) -> None:
        """Also fuse the backward exchange: the interaction backward stores dT rows into the owners'
        gradient buffers (pass `self.scatter_plan` to DotInteraction), `finish_backward()` orders the
        stores before the sparse update.  Needs enable_peer_exchange first.  ``split_dx``: dx is computed
        by a small kernel of its own ahead of the scattering pullback, which then runs on a side stream
        beside the bottom MLP's backward."""
        assert self.peer is not None
        self.scatter_plan = self.peer.enable_gradient_exchange(self.sharding, B_local, self.D, split_dx)
        if self._auto_update is not None:
            self.scatter_plan.on_launched = lambda: self._update_from_backward(fused=True)

    def lookup_fused(self, idx_local: torch.Tensor) -> torch.Tensor:
        """Forward of the fully fused path: no autograd node (the gradient never comes back through
        T; it is scattered to the owners by the interaction backward)."""
        with torch.no_grad():
            idx_owned = self._exchange_indices(idx_local)
            self.idx_owned = idx_owned
            self.presorted = False
            if len(self.local_ids):
                # up to 4096 lookups per table the sort of the coming update rides in the lookup launch
                fuse = idx_owned.shape[1] * idx_owned.shape[2] <= self.tables.FUSED_SORT_MAX
                self.tables.lookup_p2p(idx_owned, self.peer.peer_ptrs, idx_local.shape[1], 1 + self.ntab, self.idx_base,
                                       sort=fuse)
                self.presorted = fuse
            self.peer.barrier(1)
            return self.peer.T.detach()

    def _exchange_indices(self, idx_local: torch.Tensor) -> torch.Tensor:
        """Index columns to the tables' owners: peer stores + flag barrier when the index buffers are
        mapped (`enable_index_exchange`), NCCL all-to-all otherwise."""
        if self.peer is not None and self.peer._idx_dests is not None:
            owned = self.peer.scatter_indices(idx_local)
            return owned[:len(self.local_ids)] if len(self.local_ids) else owned[:0]
        return exchange_indices(idx_local, self.sharding, self.rank, self.group)

    def finish_backward(self) -> None:
        """After loss.backward() on the fused path: every rank's gradient rows have landed.  The index
        sort of this step (side stream) is joined first: once a rank passes this barrier its peers may
        start the next step and overwrite the index buffer the sort reads."""
        if len(self.local_ids) and getattr(self.tables, "_pending_side", False):
            torch.cuda.current_stream(self.tables.device).wait_event(self.tables._sorted_event)
            self.tables._pending_side = False
        plan = self.scatter_plan
        if plan is not None and plan.stream is not None and plan._keep is not None:
            torch.cuda.current_stream(self.tables.device).wait_event(plan.done)     # this rank's peer stores are issued
            plan._keep = None
        self.peer.barrier(2)
        self.owned_grad = self.peer.G

    def enable_peer_exchange(self, B_local: int, barrier: Optional[Callable] = None, flag_barrier: bool = True,
                             P: Optional[int] = None, idx_bytes: int = 4) -> None:
        """Switch the forward exchange to the fused lookup + NVLink peer-store kernel.  Safe with
        one buffer per rank when every step also runs the backward exchange (a barrier all ranks
        reach only after they have finished reading T).  With ``P`` (lookups per sample) given, the
        index exchange also goes over peer stores instead of an NCCL all-to-all."""
        if self.world == 1:
            return
        self.tables.set_slot_map([1 + k for k in self.local_ids] or [1])
        self.peer = PeerExchange(B_local, self.ntab, self.D, self.rank, self.world, self.tables.device,
                                 self.group, barrier, flag_barrier)
        if P is not None:
            self.peer.enable_index_exchange(self.sharding, B_local, P, idx_bytes)

    @classmethod
    def create(cls, rows: Sequence[int], D: int, B_local: int, P: int, rank: int, world: int, device,
               group=None, seed: int = 51234, idx_base: int = 0, dtype=None) -> "ShardedEmbedding":
        from .embedding import EmbeddingTables
        sh = TableSharding.build(rows, world)
        mine = sh.local[rank]
        kw = {} if dtype is None else {"dtype": dtype}
        tables = EmbeddingTables([rows[k] for k in mine] or [1], D, B_local * world * P, device, **kw)
        tables.init_uniform(seed + 7919 * rank)
        fused_sort = world == 1      # single GPU: the sort rides in the lookup launch
        se = cls(rows, D, rank, world,
                 lookup_fn=lambda idx, out, slot0: tables.lookup(idx, out, slot0, idx_base, sort=fused_sort),
                 update_fn=lambda idx, g, lr, slot0, presorted=False: (
                     tables.update_sorted(g, slot0, lr) if presorted else tables.bwd_sgd(idx, g, slot0, lr, idx_base)),
                 group=group,
                 sort_fn=lambda idx: tables.sort(idx, idx_base, side_stream=True),
                 idx_base=idx_base, lookup_sorts=fused_sort)
        se.tables = tables
        return se

    def lookup(self, idx_local: torch.Tensor, anchor: torch.Tensor) -> torch.Tensor:
        """idx_local [ntab][B_local][P] -> T [B_local][1 + ntab][D] (requires grad through `anchor`,
        any tensor that requires grad, so autograd calls the gradient exchange)."""
        return _ShardedLookupFn.apply(anchor, self, idx_local)

    def sort_async(self) -> None:
        """Start the index sort/dedup for this step's update on the side stream (nothing to do when
        the lookup launch already produced it)."""
        if self.presorted:
            return
        if self.sort_fn is not None and len(self.local_ids):
            self.sort_fn(self.idx_owned)
            self.presorted = True

    def update(self, lr: float, presorted: Optional[bool] = None) -> None:
        """Owner-side sparse SGD.  ``presorted`` defaults to whether this step's sort already ran
        (fused lookup launch or `sort_async`)."""
        if len(self.local_ids) == 0:
            return
        if presorted is None:
            presorted = self.presorted
        self.update_fn(self.idx_owned, self.owned_grad, lr, self.slot0, presorted)
        self.presorted = False

    def close(self) -> None:
        if self.peer is not None:
            self.peer.close()
            self.peer = None
            self.scatter_plan = None


class FlatGrads:
    """Dense (MLP) gradients as views of ONE flat buffer, so the data-parallel reduction is a
    single in-place all-reduce with no pack/unpack copies (NVSwitch: bucket sized for launch
    latency, not link count).  `zero()` before backward, `allreduce()` after; the 1/world mean
    factor is left to the caller's learning rate (`scale`)."""

    def __init__(self, params: Sequence[torch.nn.Parameter], world: int, group=None):
        self.params = list(params)
        self.world, self.group = world, group
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=self.params[0].dtype, device=self.params[0].device)
        off = 0
        self.views = []
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.scale = 1.0 / world

    def flatten_params(self) -> torch.Tensor:
        """Re-home the parameters as views of ONE flat buffer laid out like the gradient bucket, so
        the dense SGD step is a single `pflat.add_(flat, alpha=-lr)` instead of a multi-tensor pass."""
        pflat = torch.empty_like(self.flat)
        off = 0
        with torch.no_grad():
            for p in self.params:
                view = pflat[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                off += p.numel()
        self.pflat = pflat
        return pflat

    def zero(self) -> None:
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def allreduce(self) -> None:
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)


def allreduce_dense_grads(params: Sequence[torch.nn.Parameter], world: int, group=None) -> None:
    """Sum the data-parallel MLP gradients with ONE all-reduce over a flat bucket (a few MB: sized
    for launch latency, NVSwitch gives every pair full bandwidth).  The loss is the mean over the
    local batch, so the global-batch mean needs a 1/world factor."""
    if world == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.mul_(1.0 / world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n

"""Expert-parallel (EP) execution of the MoE layer over the GPUs of one NVLink box.

Sharding is the reference's (utils/common_config.py:179-185, utils/moe_utils.py:191-198):
tokens stay data-parallel, the router is replicated over all E_tot = W * E_loc experts,
rank r owns global experts [r*E_loc, (r+1)*E_loc).  What changes is the exchange.
FastMoE's global_scatter / global_gather (grouped ncclSend/ncclRecv sized by HOST
counts, i.e. one D2H sync per layer, reached from MOEScatter/MOEGather when
world_size > 1) is replaced by peer-mapped expert queues:

  * every rank owns one CUDA-IPC "arena"; all ranks sub-allocate it in lockstep, so a
    queue lives at the same offset on every rank and `base[p] + offset` is rank p's queue;
  * forward : all-gather of the [E_tot] count vector by OUR 1-warp kernel over peer memory
    (m3_ep_barrier; stream-ordered, NO host read-back, no NCCL on the data path) -> m3_ep_plan
    (every slot's owner rank + row) -> the dispatch kernel STORES token rows straight into the
    owners' queues over NVLink -> barrier -> grouped expert FFN on the local queue -> barrier ->
    the combine kernel LOADS result rows from the owners (and keeps a local copy for backward);
  * backward: combine_bwd pushes dyq (dscore from the local copy), dispatch_bwd pulls dxq.
    Expert weight gradients need no reduction (each expert's rows are all on its owner).
  * bf16 (the tcgen05 path), default: the RETURN half of both exchanges is fused into the expert GEMMs
    ("return store", m3_ep_ffn_fwd / m3_ep_ffn_bwd): the push also records every row's origin
    (source rank, slot) at the owner, and the epilogue of fc2 (forward) / of dxq = dz W1 (backward) stores each
    result row straight into the source rank's slot-ordered return buffer over NVLink, tile by tile, while the
    tensor cores work on the next tile.  No result queue exists on the owner, nothing is pulled; after the
    rendezvous the source combines its own return buffer with the local kernels.  M3_EP_RETURN=0 (or fp32
    queues) selects the pull protocol above.
    The push half stays a kernel of its own: running it on 16-64 CTAs of the first GEMM's launch behind arrival
    counters was built and measured (cp.async row pipelines, TMA bulk copies; bit-identical) and lost, 72.8 M against
    83.6 M tokens/s at 2 GPUs - a row mover needs all 148 SMs to keep NVLink busy.

The protocol is written as explicit phases so that it runs either over torch.distributed
(one process per GPU, NCCL) or as a single-process multi-rank simulation (tests; one GPU
holds every rank's arena, the "peers" are just other device pointers).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import ops
from ._lib import PAD_ROWS, check, dtype_code, load, ptr, stream_ptr


# ----------------------------------------------------------------------------- arena
class _CAI:
    def __init__(self, p, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (p, False), "version": 2}


class Arena:
    """A cudaMalloc'ed, IPC-exportable slab with a deterministic size-class allocator.
    Every rank performs the same alloc/free sequence, hence gets the same offsets."""

    ALIGN = 1024

    def __init__(self, nbytes: int, device: torch.device):
        self.nbytes = int(nbytes)
        self.device = device
        p = C.c_void_p()
        self.handle = (C.c_ubyte * 64)()
        with torch.cuda.device(device):
            check(load().m3_ipc_alloc(self.nbytes, C.byref(p), self.handle), "m3_ipc_alloc")
        self.base = p.value
        self._bytes = torch.as_tensor(_CAI(self.base, self.nbytes), device=device)
        self._top = 0
        self._free = {}       # size -> [offsets]

    def alloc(self, nbytes: int) -> int:
        n = (int(nbytes) + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        lst = self._free.get(n)
        if lst:
            return lst.pop()
        off = self._top
        if off + n > self.nbytes:
            raise MemoryError(f"EP arena exhausted ({self.nbytes} B): raise arena_bytes or lower capacity_factor")
        self._top = off + n
        return off

    def free(self, off: int, nbytes: int) -> None:
        n = (int(nbytes) + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self._free.setdefault(n, []).append(off)

    def view(self, off: int, rows: int, cols: int, dtype: torch.dtype) -> torch.Tensor:
        nb = rows * cols * torch.empty((), dtype=dtype).element_size()
        return self._bytes[off:off + nb].view(dtype).view(rows, cols)

    def handle_bytes(self) -> bytes:
        return bytes(self.handle)

    def close(self):
        if self.base:
            self._bytes = None
            load().m3_ipc_free(C.c_void_p(self.base))
            self.base = 0


# ----------------------------------------------------------------------------- groups
class TorchDistGroup:
    """torch.distributed (NCCL on GPUs; gloo in the CPU tests of the count exchange)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._flag = None

    def all_gather_counts(self, counts: torch.Tensor) -> torch.Tensor:
        out = torch.empty(self.world * counts.numel(), dtype=counts.dtype, device=counts.device)
        self.dist.all_gather_into_tensor(out, counts.contiguous().view(-1), group=self.group)
        return out.view(self.world, counts.numel())

    def barrier(self, device) -> None:
        """Stream-ordered rendezvous: a 1-element all-reduce.  No host synchronisation."""
        if self._flag is None or self._flag.device != device:
            self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        self.dist.all_reduce(self._flag, group=self.group)

    def exchange_bytes(self, payload: bytes) -> List[bytes]:
        out = [None] * self.world
        self.dist.all_gather_object(out, payload, group=self.group)
        return out


class PeerFlagGroup:
    """Rendezvous + count all-gather done by OUR kernel over peer memory (m3_ep_barrier) instead of
    NCCL: ~1 launch of a 1-warp kernel per synchronisation point, no host involvement.  Uses the
    first HEADER bytes of every rank's arena (flags [W] + gather buffer)."""
    HEADER = 64 * 1024

    def __init__(self, rank, world, arena: "Arena", bases: torch.Tensor, bootstrap):
        self.rank, self.world = rank, world
        self.bootstrap = bootstrap                         # TorchDistGroup: only for setup / teardown
        self.arena, self.bases = arena, bases
        self.epoch = 0
        off_flags = arena.alloc(4096)
        self.off_gather = arena.alloc(self.HEADER - 4096)
        assert off_flags == 0
        arena._bytes[: self.HEADER].zero_()
        self.flag_ptrs = bases + off_flags
        self._gp_cache = {}
        self.gather_ptrs = bases + self.off_gather
        torch.cuda.synchronize(arena.device)
        bootstrap.barrier(arena.device)                    # flags are zeroed everywhere before first use
        torch.cuda.synchronize(arena.device)

    def _sync(self, payload, n):
        self.epoch += 1
        check(load().m3_ep_barrier(ptr(self.flag_ptrs), ptr(self.gather_ptrs) if payload is not None else None,
                                   ptr(payload), n, self.rank, self.world, self.epoch, stream_ptr()), "m3_ep_barrier")
        ops.launch_count += 1

    def barrier(self, device) -> None:
        self._sync(None, 0)

    def all_gather_counts(self, counts: torch.Tensor) -> torch.Tensor:
        n = counts.numel()
        assert counts.dtype == torch.int32 and 2 * self.world * n * 4 <= self.HEADER - 4096
        # double-buffered rows: a fast peer may already push the NEXT layer's counts while I still read these
        half = (self.epoch + 1) & 1
        goff = self.off_gather + half * (self.HEADER - 4096) // 2
        self.epoch += 1
        gp = self._gp_cache.get(goff)
        if gp is None:
            gp = self._gp_cache[goff] = self.bases + goff
        check(load().m3_ep_barrier(ptr(self.flag_ptrs), ptr(gp), ptr(counts), n, self.rank, self.world, self.epoch,
                                   stream_ptr()), "m3_ep_barrier")
        ops.launch_count += 1
        return self.arena.view(goff, self.world, n, torch.int32).clone()


@dataclass
class EPContext:
    """Per-rank EP state: arena + peer-mapped bases of every rank's arena."""
    rank: int
    world: int
    group: object
    arena: Arena
    bases: torch.Tensor                 # [W] int64 device tensor: arena base pointer of every rank
    capacity_factor: Optional[float]    # receive-queue rows = factor * T*K (None: worst case W - nothing can be dropped)
    overflow: torch.Tensor              # [1] int32 device flag set by m3_ep_plan

    def peer_ptrs(self, off: int) -> torch.Tensor:
        """[W] device array of `arena base + off` of every rank.  Cached: the lockstep allocator hands out the same
        few offsets call after call, and `bases + off` would otherwise launch a kernel each time."""
        cache = self.__dict__.setdefault("_peer_cache", {})
        t = cache.get(off)
        if t is None:
            t = cache[off] = self.bases + off
        return t

    def cap_rows(self, T: int, K: int, E_loc: int) -> int:
        f = float(self.world) if self.capacity_factor is None else min(float(self.world), self.capacity_factor)
        rows = int(f * T * K) + E_loc * (PAD_ROWS - 1)
        return (rows + PAD_ROWS - 1) // PAD_ROWS * PAD_ROWS

    def check_overflow(self) -> None:
        """Host-synchronising check (call outside the hot loop)."""
        if int(self.overflow.item()) != 0:
            raise RuntimeError("EP receive queue overflow: tokens were dropped; raise capacity_factor")

    def fused_return(self, cdt, T: int, K: int) -> bool:
        """Return store (module docstring): bf16 queues, row origins packed into 32 bits.  Default for world sizes up to 4:
        measured on one 8 x B200 box (T = 38 432 per rank) the epilogue's 128-byte row segments, scattered over 7 peers,
        reach only ~300 GB/s per GPU (fc2 340 us instead of 70) where the pulling combine reads 570 GB/s - 203 M against
        248 M tokens/s at 8 GPUs - while at 2 GPUs they run at ~500 GB/s and the fusion wins (83.9 M against 79.6 M).
        M3_EP_RETURN = 1 / 0 forces it on / off."""
        import os
        env = os.environ.get("M3_EP_RETURN", "")
        if cdt != torch.bfloat16 or env == "0" or self.world >= 128 or T * K > (1 << 24):
            return False
        return env == "1" or self.world <= 4

    def poll_overflow(self) -> None:
        """Asynchronous check, run by every layer call: the flag of an EARLIER call is copied to pinned host memory on the
        stream (no synchronisation) and read once that copy has completed - a dropped slot raises at most a few calls
        late instead of silently corrupting the step.  With capacity_factor=None queues cannot overflow and this is free."""
        if self.capacity_factor is None:
            return
        st = self.__dict__.setdefault("_ovf", {})
        if "host" not in st:
            st["host"] = torch.zeros(1, dtype=torch.int32).pin_memory()
            st["event"] = None
        ev = st["event"]
        if ev is not None and ev.query():
            if int(st["host"][0]) != 0:
                raise RuntimeError("EP receive queue overflow: a rank received more than capacity_factor * T * K rows and "
                                   "slots were dropped; raise capacity_factor (None = worst case, nothing can be dropped)")
            ev = None
        if ev is None:
            st["host"].copy_(self.overflow, non_blocking=True)
            st["event"] = torch.cuda.Event()
            st["event"].record()

    def check_same_tokens(self, T: int) -> None:
        """The lockstep arena gives a queue the same offset on every rank only if every rank sizes it the same, i.e. runs the
        same number of tokens per call.  Checked on the host once per distinct T (a bootstrap-group all-gather)."""
        seen = self.__dict__.setdefault("_seen_T", set())
        if T in seen:
            return
        boot = getattr(self.group, "bootstrap", self.group)
        if hasattr(boot, "exchange_bytes"):
            all_T = [int(b.decode()) for b in boot.exchange_bytes(str(int(T)).encode())]
            if any(t != T for t in all_T):
                raise RuntimeError(f"expert parallelism needs the same number of tokens per call on every rank (the peer queues "
                                   f"live at lockstep arena offsets); got {all_T}.  Pad the local batch to a common size.")
        seen.add(T)

    # Arena blocks of a finished call are handed back only after the NEXT rendezvous (the count all-gather that opens every
    # forward): a peer that has passed it has finished every pull of the previous call, so no separate closing barrier is
    # needed (one rendezvous kernel fewer per forward-only call and per backward).
    def defer_free(self, off: int, nbytes: int) -> None:
        self.__dict__.setdefault("_pending", []).append((off, nbytes))

    def apply_deferred_frees(self) -> None:
        for off, nb in self.__dict__.get("_pending", []):
            self.arena.free(off, nb)
        self.__dict__["_pending"] = []


def make_context(group, device, arena_bytes: int, capacity_factor: Optional[float] = None,
                 device_barrier: bool = True) -> EPContext:
    """Collective: allocates the arena, exchanges IPC handles, maps the peers.  With
    device_barrier=True (default) the per-layer rendezvous / count exchange run as our own
    peer-memory kernel; otherwise through the torch.distributed group (NCCL).

    capacity_factor=None (default) sizes every receive queue for the worst case (all W ranks route everything to one
    rank): like the reference's FastMoE path, nothing can ever be dropped.  A float caps the queue at factor * T * K
    rows (less arena memory); an overflow then drops slots on the device and RAISES on the host within a few calls
    (EPContext.poll_overflow, run by every layer call).

    Evaluation must run under torch.no_grad(): a forward with grad enabled keeps its queues until its backward runs, and
    frees must happen in the same order on every rank (garbage-collection order is not), so they are never freed by GC."""
    arena = Arena(arena_bytes + PeerFlagGroup.HEADER, device)
    handles = group.exchange_bytes(arena.handle_bytes())
    bases = []
    lib = load()
    with torch.cuda.device(device):
        for r, h in enumerate(handles):
            if r == group.rank:
                bases.append(arena.base)
            else:
                p = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                check(lib.m3_ipc_open(buf, C.byref(p)), "m3_ipc_open")
                bases.append(p.value)
    bases_t = torch.tensor(bases, dtype=torch.int64, device=device)
    sync_group = PeerFlagGroup(group.rank, group.world, arena, bases_t, group) if device_barrier else group
    return EPContext(group.rank, group.world, sync_group, arena, bases_t, capacity_factor,
                     torch.zeros(1, dtype=torch.int32, device=device))


# ----------------------------------------------------------------------------- phases
@dataclass
class EPFwdState:
    g: ops.GateOut
    plan_local: ops.Plan
    dst_rank: Optional[torch.Tensor] = None
    dst_row: Optional[torch.Tensor] = None
    recv: Optional[ops.Plan] = None
    cap: int = 0
    off_xq: int = -1
    off_yq: int = -1
    xq: Optional[torch.Tensor] = None
    yq: Optional[torch.Tensor] = None
    hpre: Optional[torch.Tensor] = None
    ysave: Optional[torch.Tensor] = None      # local copy of the pulled result rows [T*K, D]
    nbytes_q: int = 0
    # return-store protocol (bf16): row origins of my receive queue, my slot-ordered return buffer (= ysave), identity plan
    ret: bool = False
    off_inv: int = -1
    inv: Optional[torch.Tensor] = None        # [T*K] sorted position -> slot, in the arena (peers read it)
    off_meta: int = -1
    off_yret: int = -1
    meta: Optional[torch.Tensor] = None
    pos_id: Optional[torch.Tensor] = None
    nbytes_ret: int = 0


def phase_a_gate(gx, w_gate, top_k, task_feat, noise, noise_stddev, want_gates, E_tot, ctx: Optional[EPContext] = None,
                 cdt=None) -> EPFwdState:
    """local: router over all E_tot experts + local stable plan (pad 1) -> this rank's count vector.  With the fused
    exchange (ctx / cdt given, bf16) the plan's inverse - my slots sorted by expert, the send order - goes into the arena,
    where the owners read the origins of their rows from it."""
    g = ops.gate_fwd(gx, w_gate, top_k, task_feat, noise, noise_stddev, want_gates)
    T = gx.shape[0]
    st = EPFwdState(g, None)
    if ctx is not None and ctx.fused_return(cdt, T, top_k) and T > 0:
        st.ret = True
        st.off_inv = ctx.arena.alloc(T * top_k * 4)
        st.inv = ctx.arena.view(st.off_inv, T * top_k, 1, torch.int32)
    st.plan_local = ops.route_plan(g.idx, E_tot, 1, g.imp_partial, g.load_partial, inv_pos=st.inv)
    return st


def phase_b_dispatch(ctx: EPContext, st: EPFwdState, x, cnt_all, E_loc, top_k, cdt) -> None:
    """m3_ep_plan, then PUSH token rows into the owners' queues (NVLink stores)."""
    lib = load()
    T, D = x.shape
    dev = x.device
    W = ctx.world
    st.cap = ctx.cap_rows(T, top_k, E_loc)
    R = T * top_k
    st.dst_rank = torch.empty(R, dtype=torch.int32, device=dev)
    st.dst_row = torch.empty(R, dtype=torch.int32, device=dev)
    rc = torch.empty(E_loc, dtype=torch.int32, device=dev)
    ro = torch.empty(E_loc + 1, dtype=torch.int32, device=dev)
    rt = torch.empty(st.cap // PAD_ROWS, dtype=torch.int32, device=dev)
    st.ret = st.ret and st.inv is not None
    el = 2 if cdt == torch.bfloat16 else 4
    peers_inv = None
    if st.ret:
        st.pos_id = torch.empty(R, dtype=torch.int32, device=dev)
        st.nbytes_ret = R * D * el
        st.off_meta = ctx.arena.alloc(st.cap * 4)
        st.meta = ctx.arena.view(st.off_meta, st.cap, 1, torch.int32)
        peers_inv = ctx.peer_ptrs(st.off_inv)
    check(lib.m3_ep_plan(ptr(st.g.idx), ptr(st.plan_local.pos), ptr(cnt_all), ctx.rank, W, E_loc, T, top_k, PAD_ROWS,
                         st.cap, ptr(st.dst_rank), ptr(st.dst_row), ptr(rc), ptr(ro), ptr(rt), ptr(ctx.overflow),
                         ptr(st.pos_id), ptr(peers_inv), ptr(st.meta), stream_ptr()), "m3_ep_plan")
    st.recv = ops.Plan(rc, ro, None, rt, st.cap, PAD_ROWS)
    st.nbytes_q = st.cap * D * el
    st.off_xq = ctx.arena.alloc(st.nbytes_q)
    st.xq = ctx.arena.view(st.off_xq, st.cap, D, cdt)
    check(lib.m3_ep_dispatch_fwd(ptr(x), dtype_code(x), ptr(st.dst_rank), ptr(st.dst_row), T, top_k, D,
                                 ptr(ctx.peer_ptrs(st.off_xq)), dtype_code(st.xq), stream_ptr()), "m3_ep_dispatch_fwd")
    check(lib.m3_zero_pad_rows(ptr(st.xq), dtype_code(st.xq), ptr(rc), ptr(ro), E_loc, D, ptr(st.meta), stream_ptr()),
          "m3_zero_pad_rows")
    ops.launch_count += 3


def phase_c_ffn(ctx: EPContext, st: EPFwdState, w1c, b1, w2c, b2, save_hpre: bool, drop=None) -> None:
    """local: grouped expert FFN over this rank's receive queue -> yq (in the arena, peers read it)"""
    lib = load()
    cap, D = st.xq.shape
    E_loc, H, _ = w1c.shape
    dt = dtype_code(st.xq)
    st.hpre = (torch.empty(max(int(lib.m3_ffn_saved_bytes(dt, cap, D, H)), 16), dtype=torch.uint8, device=st.xq.device)
               if save_hpre else None)        # opaque activation state for phase F
    ws = torch.empty(max(lib.m3_ffn_workspace_bytes(dt, cap, D, H, E_loc, 0), 16), dtype=torch.uint8, device=st.xq.device)
    p_drop, rng = (float(drop[0]), ptr(drop[1])) if (drop is not None and drop[0] > 0 and save_hpre) else (0.0, None)
    if st.ret:      # fc2's epilogue stores every result row into its SOURCE rank's return buffer (slot order)
        st.off_yret = ctx.arena.alloc(st.nbytes_ret)
        st.ysave = ctx.arena.view(st.off_yret, st.nbytes_ret // (D * 2), D, st.xq.dtype)
        check(lib.m3_ep_ffn_fwd(dt, ptr(st.xq), ptr(st.recv.offsets), ptr(st.recv.tile_expert), cap, E_loc, D, H, ptr(w1c),
                                ptr(b1), ptr(w2c), ptr(b2), ptr(st.hpre), ptr(st.meta), ptr(ctx.peer_ptrs(st.off_yret)),
                                ptr(ws), ws.numel(), p_drop, rng, stream_ptr()), "m3_ep_ffn_fwd")
        ops.launch_count += 2
        return
    st.off_yq = ctx.arena.alloc(st.nbytes_q)
    st.yq = ctx.arena.view(st.off_yq, cap, D, st.xq.dtype)
    if drop is not None and drop[0] > 0 and save_hpre:      # expert dropout: the mask is a function of the OWNER's queue row
        check(lib.m3_ffn_fwd_dropout(dt, ptr(st.xq), ptr(st.recv.offsets), ptr(st.recv.tile_expert), cap, E_loc, D, H,
                                     ptr(w1c), ptr(b1), ptr(w2c), ptr(b2), ptr(st.hpre), ptr(st.yq), ptr(ws), ws.numel(),
                                     float(drop[0]), ptr(drop[1]), stream_ptr()), "m3_ffn_fwd_dropout")
    else:
        check(lib.m3_ffn_fwd(dt, ptr(st.xq), ptr(st.recv.offsets), ptr(st.recv.tile_expert), cap, E_loc, D, H, ptr(w1c),
                             ptr(b1), ptr(w2c), ptr(b2), ptr(st.hpre), ptr(st.yq), ptr(ws), ws.numel(), stream_ptr()),
              "m3_ffn_fwd")
    ops.launch_count += 2


def phase_d_combine(ctx: EPContext, st: EPFwdState, T, D, top_k, out_dtype, keep_rows: bool = True,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """PULL result rows from the owners' yq queues and combine with the gate scores.  With `keep_rows` the
    pulled rows are also kept locally (slot order) so that the backward pass does not pull them again."""
    if out is None:
        out = torch.empty(T, D, dtype=out_dtype, device=st.g.score.device)
    assert out.is_contiguous() and out.shape == (T, D)
    if st.ret:      # every owner has stored my rows into my return buffer: an ordinary local combine over it
        check(load().m3_combine_fwd(ptr(st.ysave), dtype_code(st.ysave), ptr(st.pos_id), ptr(st.g.score), T, top_k, D,
                                    ptr(out), dtype_code(out), stream_ptr()), "m3_combine_fwd")
        ops.launch_count += 1
        return out
    peers = ctx.peer_ptrs(st.off_yq)
    st.ysave = torch.empty(T * top_k, D, dtype=st.yq.dtype, device=out.device) if keep_rows else None
    check(load().m3_ep_combine_fwd(ptr(peers), dtype_code(st.yq), ptr(st.dst_rank), ptr(st.dst_row), ptr(st.g.score),
                                   T, top_k, D, ptr(out), dtype_code(out), ptr(st.ysave), stream_ptr()),
          "m3_ep_combine_fwd")
    ops.launch_count += 1
    return out


@dataclass
class EPBwdState:
    off_dyq: int = -1
    off_dxq: int = -1
    dyq: Optional[torch.Tensor] = None
    dxq: Optional[torch.Tensor] = None
    dscore: Optional[torch.Tensor] = None
    grads: Optional[tuple] = None
    ws: Optional[torch.Tensor] = None
    off_dxret: int = -1
    dxret: Optional[torch.Tensor] = None      # return store: my slot-ordered [T*K, D] buffer the owners fill


def phase_e_combine_bwd(ctx: EPContext, st: EPFwdState, g_out, top_k) -> EPBwdState:
    """dscore from the owners' yq rows (pull); dyq = score * g pushed into the owners' dyq queues."""
    lib = load()
    bs = EPBwdState()
    T, D = g_out.shape
    E_loc = st.recv.counts.numel()
    bs.off_dyq = ctx.arena.alloc(st.nbytes_q)
    bs.dyq = ctx.arena.view(bs.off_dyq, st.cap, D, st.xq.dtype)
    bs.dscore = torch.empty(T, top_k, dtype=torch.float32, device=g_out.device)
    py = ctx.peer_ptrs(st.off_yq) if st.off_yq >= 0 else None      # (not needed: dscore comes from the local copy)
    pd = ctx.peer_ptrs(bs.off_dyq)
    check(lib.m3_ep_combine_bwd(ptr(g_out), dtype_code(g_out), ptr(py), ptr(pd), dtype_code(bs.dyq), ptr(st.dst_rank),
                                ptr(st.dst_row), ptr(st.g.score), T, top_k, D, ptr(bs.dscore), ptr(st.ysave),
                                stream_ptr()), "m3_ep_combine_bwd")
    check(lib.m3_zero_pad_rows(ptr(bs.dyq), dtype_code(bs.dyq), ptr(st.recv.counts), ptr(st.recv.offsets), E_loc, D,
                               None, stream_ptr()), "m3_zero_pad_rows")
    ops.launch_count += 2
    return bs


def phase_f_ffn_bwd(ctx: EPContext, st: EPFwdState, bs: EPBwdState, w1c, w2c, w1t, w2t, drop=None, parts: int = 3) -> None:
    """local: expert FFN backward over this rank's receive queue.  parts = 1: data gradients (dxq, which the peers pull
    next - or, with the return store, which the epilogue sends straight back to them), 2: weight gradients, 3: both."""
    lib = load()
    cap, D = st.xq.shape
    E_loc, H, _ = w1c.shape
    dt = dtype_code(st.xq)
    dev = st.xq.device
    if parts & 1:
        if st.ret:
            bs.off_dxret = ctx.arena.alloc(st.nbytes_ret)
            bs.dxret = ctx.arena.view(bs.off_dxret, st.nbytes_ret // (D * 2), D, st.xq.dtype)
        else:
            bs.off_dxq = ctx.arena.alloc(st.nbytes_q)
            bs.dxq = ctx.arena.view(bs.off_dxq, cap, D, st.xq.dtype)
        bs.grads = (torch.empty(E_loc, H, D, dtype=torch.float32, device=dev), torch.empty(E_loc, H, dtype=torch.float32, device=dev),
                    torch.empty(E_loc, D, H, dtype=torch.float32, device=dev), torch.empty(E_loc, D, dtype=torch.float32, device=dev))
        bs.ws = torch.empty(max(lib.m3_ffn_workspace_bytes(dt, cap, D, H, E_loc, 1), 16), dtype=torch.uint8, device=dev)
    dw1, db1, dw2, db2 = bs.grads
    p, rng = (float(drop[0]), ptr(drop[1])) if (drop is not None and drop[0] > 0) else (0.0, None)
    if st.ret:
        check(lib.m3_ep_ffn_bwd(dt, ptr(st.xq), ptr(st.hpre), ptr(bs.dyq), ptr(st.recv.counts), ptr(st.recv.offsets),
                                ptr(st.recv.tile_expert), cap, E_loc, D, H, ptr(w1c), ptr(w2c), ptr(w1t), ptr(w2t),
                                ptr(st.meta), ptr(ctx.peer_ptrs(bs.off_dxret)), ptr(dw1), ptr(db1), ptr(dw2), ptr(db2),
                                ptr(bs.ws), bs.ws.numel(), p, rng, int(parts), stream_ptr()), "m3_ep_ffn_bwd")
    else:
        check(lib.m3_ffn_bwd_parts(dt, ptr(st.xq), ptr(st.hpre), ptr(bs.dyq), ptr(st.recv.counts), ptr(st.recv.offsets),
                                   ptr(st.recv.tile_expert), cap, E_loc, D, H, ptr(w1c), ptr(w2c), ptr(w1t), ptr(w2t),
                                   ptr(bs.dxq), ptr(dw1), ptr(db1), ptr(dw2), ptr(db2), ptr(bs.ws), bs.ws.numel(), p, rng,
                                   int(parts), stream_ptr()), "m3_ffn_bwd_parts")
    if parts & 1:
        ops.launch_count += 2 if dt == 1 else 2
    if parts & 2:
        ops.launch_count += 2 if dt == 1 else 4


def phase_g_dispatch_bwd(ctx: EPContext, st: EPFwdState, bs: EPBwdState, T, D, top_k, dz, w_gate, out_dtype,
                         out: Optional[torch.Tensor] = None):
    """PULL dxq rows from the owners and sum them per token (+ the router's dx)."""
    dx = out if out is not None else torch.empty(T, D, dtype=out_dtype, device=st.g.score.device)
    assert dx.is_contiguous() and dx.shape == (T, D)
    E = w_gate.shape[1] if dz is not None else 0
    if st.ret:      # the owners' dgrad epilogues have filled my return buffer: local gather-sum (+ the router's dx)
        check(load().m3_dispatch_bwd(ptr(bs.dxret), dtype_code(bs.dxret), ptr(st.pos_id), T, top_k, D, ptr(dz),
                                     ptr(w_gate) if dz is not None else None, E, ptr(dx), dtype_code(dx), stream_ptr()),
              "m3_dispatch_bwd")
        ops.launch_count += 1
        return dx
    peers = ctx.peer_ptrs(bs.off_dxq)
    check(load().m3_ep_dispatch_bwd(ptr(peers), dtype_code(bs.dxq), ptr(st.dst_rank), ptr(st.dst_row), T, top_k, D,
                                    ptr(dz), ptr(w_gate) if dz is not None else None, E, ptr(dx), dtype_code(dx),
                                    stream_ptr()), "m3_ep_dispatch_bwd")
    ops.launch_count += 1
    return dx


def release_fwd(ctx: EPContext, st: EPFwdState) -> None:
    for off in (st.off_yq, st.off_xq):
        if off >= 0:
            ctx.defer_free(off, st.nbytes_q)
    if st.off_meta >= 0:
        ctx.defer_free(st.off_meta, st.cap * 4)
    if st.off_inv >= 0:
        ctx.defer_free(st.off_inv, st.inv.numel() * 4)
        st.off_inv, st.inv = -1, None
    if st.off_yret >= 0:
        ctx.defer_free(st.off_yret, st.nbytes_ret)
        st.ysave = None
    st.off_xq = st.off_yq = st.off_meta = st.off_yret = -1
    st.xq = st.yq = st.meta = None


def release_bwd(ctx: EPContext, st: EPFwdState, bs: EPBwdState) -> None:
    for off in (bs.off_dxq, bs.off_dyq):
        if off >= 0:
            ctx.defer_free(off, st.nbytes_q)
    if bs.off_dxret >= 0:
        ctx.defer_free(bs.off_dxret, st.nbytes_ret)
    bs.off_dxq = bs.off_dyq = bs.off_dxret = -1
    bs.dxq = bs.dyq = bs.dxret = None


# ----------------------------------------------------------------------------- autograd over torch.distributed
class EPMoEFunction(torch.autograd.Function):
    """Same contract as functions.MoEFunction, experts sharded over `ctx_ep.world` ranks.
    w1/b1/w2/b2 are this rank's LOCAL experts [E_loc, ...]; w_gate is [Dg, W*E_loc]."""

    @staticmethod
    def forward(ctx, x, gate_x, w_gate, task_feat, w1, b1, w2, b2, noise, top_k, noise_stddev, compute_dtype,
                want_gates, wcache, ep: EPContext, drop=None):
        ctx.set_materialize_grads(False)      # undefined output grads stay None (no zero fills)
        T, D = x.shape
        E_loc = w1.shape[0]
        E_tot = w_gate.shape[1]
        assert E_tot == E_loc * ep.world, "w_gate must have world_size * num_expert columns"
        x = x.contiguous()
        gx = x if gate_x is None else gate_x.contiguous()
        grp = ep.group
        ep.check_same_tokens(T)
        ep.poll_overflow()
        st = phase_a_gate(gx, w_gate, top_k, task_feat, noise, noise_stddev, want_gates, E_tot, ep, compute_dtype)
        cnt_all = grp.all_gather_counts(st.plan_local.counts)         # also: "all queues are free" rendezvous
        ep.apply_deferred_frees()                                     # ... so the previous call's blocks can be reused
        phase_b_dispatch(ep, st, x, cnt_all, E_loc, top_k, compute_dtype)
        if compute_dtype == torch.bfloat16:
            w1c, w2c, w1t, w2t = wcache.get_bf16(w1, w2)
        else:
            w1c, w2c, w1t, w2t = w1, w2, None, None
        needs_grad = any(ctx.needs_input_grad)
        grp.barrier(x.device)                                         # every push has landed
        phase_c_ffn(ep, st, w1c, b1, w2c, b2, needs_grad, drop)
        grp.barrier(x.device)                                         # every owner's yq is complete
        out = phase_d_combine(ep, st, T, D, top_k, x.dtype, keep_rows=needs_grad)
        g, pl = st.g, st.plan_local
        if needs_grad:
            ctx.st, ctx.ep = st, ep
            ctx.save_for_backward(x, gate_x, w_gate, task_feat, w1c, w2c, w1t, w2t, pl.importance)
            ctx.cfg = (top_k, gate_x is not None)
            ctx.drop = drop
        else:
            release_fwd(ep, st)                                       # handed back after the next rendezvous
        gates = g.gates if g.gates is not None else x.new_empty(0)
        ctx.mark_non_differentiable(g.idx, pl.load, pl.counts)
        noisy = g.clean_logits.view_as(g.clean_logits) if noise is None else g.noisy_logits
        return out, g.score, g.top_vals, g.clean_logits, noisy, gates, pl.importance, pl.load, g.idx, pl.counts, pl.cv_loss

    @staticmethod
    def backward(ctx, d_out, d_score, d_top, d_clean, d_noisy, d_gates, d_imp, _dl, _di, _dc, d_cv):
        x, gate_x, w_gate, task_feat, w1c, w2c, w1t, w2t, importance = ctx.saved_tensors
        st, ep = ctx.st, ctx.ep
        top_k, separate = ctx.cfg
        T, D = x.shape
        grp = ep.group
        if d_out is None:
            d_out = torch.zeros_like(x)
        bs = phase_e_combine_bwd(ep, st, d_out.contiguous(), top_k)     # (fresh dyq block: nobody reads or writes it yet)
        grp.barrier(x.device)                                         # every dy push has landed
        # Data gradients first: they are what the peers wait for (with the return store the dgrad epilogue sends every dxq
        # row home over NVLink while the tensor cores run the next tile).  The weight gradients and the router backward
        # need nothing from the peers and run before the rendezvous, where they also absorb rank skew.  (Running the weight
        # gradients on a side stream beside the pull was measured at 2 GPUs and gains nothing: a GEMM CTA holds ~60 k of
        # an SM's 64 k registers, so the movers cannot become resident next to it.)
        phase_f_ffn_bwd(ep, st, bs, w1c, w2c, w1t, w2t, ctx.drop, parts=1)
        dscore = bs.dscore if d_score is None else bs.dscore + d_score
        if d_gates is not None and d_gates.numel() == 0:
            d_gates = None
        gx = x if gate_x is None else gate_x
        dz, dwg, dtf, dxg = ops.gate_bwd(gx, w_gate, st.g.noisy_logits, st.g.idx_full, top_k, task_feat, dscore, d_top,
                                         d_gates, d_imp, d_clean, d_noisy, want_dx_gate=separate,
                                         importance=importance, dcv_loss=d_cv)
        phase_f_ffn_bwd(ep, st, bs, w1c, w2c, w1t, w2t, ctx.drop, parts=2)
        grp.barrier(x.device)                                         # every owner's dxq is complete / has come home
        if separate:
            dx = phase_g_dispatch_bwd(ep, st, bs, T, D, top_k, None, w_gate, x.dtype)
            dgx = dxg.to(gate_x.dtype)
        else:
            dx = phase_g_dispatch_bwd(ep, st, bs, T, D, top_k, dz, w_gate, x.dtype)
            dgx = None
        dw1, db1, dw2, db2 = bs.grads
        release_bwd(ep, st, bs)                                       # handed back after the next rendezvous
        release_fwd(ep, st)
        ctx.st = None
        if dtf is not None and task_feat is not None:
            dtf = dtf.view_as(task_feat).to(task_feat.dtype)
        return dx, dgx, dwg, dtf, dw1, db1, dw2, db2, None, None, None, None, None, None, None, None


class _EPRunner:
    def __init__(self, ep: EPContext):
        self.ep = ep

    def forward(self, layer, gate, x, gx, tf, noise, nstd, cdt, drop=None):
        return EPMoEFunction.apply(
            x, gx, gate.w_gate, tf, layer.experts.htoh4.weight, layer.experts.htoh4.bias,
            layer.experts.h4toh.weight, layer.experts.h4toh.bias, noise, layer.top_k, nstd, cdt,
            layer.RETURN_SUMMARIES, layer._wcache, self.ep, drop)


def attach(layer, ep) -> None:
    """Enable expert parallelism on an FMoETransformerMLP built with world_size = ep.world
    (num_expert = experts per rank, as the reference's get_backbone does)."""
    if layer.world_size != ep.world:
        raise ValueError(f"layer.world_size={layer.world_size} but the EP group has {ep.world} ranks")
    layer._ep = _EPRunner(ep)

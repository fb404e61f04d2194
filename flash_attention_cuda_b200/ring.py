"""Multi-GPU drivers for the forward path: one process per GPU over torch.distributed.

* batch x heads sharding (config 3): every (b, h) is an independent problem (reference
  flash_attention.cu:120-122), so ranks take contiguous slices of B*H and never communicate.
* context parallelism (config 5): the sequence is cut into 2P chunks, rank r owns chunks r and
  2P-1-r (zig-zag: every hop costs every rank exactly two unmasked chunk pairs under a causal mask);
  Q and the partial state stay.  Two ways of getting at the other ranks' K/V:
    - exchange="pull" (default on CUDA): nothing travels around a ring.  Every rank keeps its K/V chunk
      pair in peer-readable memory (`PeerKV`, flash_attn_peer_* in include/flash_attn.h); at hop h a rank
      pulls the chunks of rank r-h it needs -- only the unmasked ones -- straight out of the owner's HBM
      with a copy engine over NVLink/NVSwitch, into one of two local landing buffers, while the kernels
      of hop h-1 run.  No communication kernel competes with the persistent attention grid for SMs, and
      the ranks are not chained to each other: two tiny all-reduces per step (everyone's K/V in place /
      everyone done reading) are the only collectives.
    - exchange="gather" (causal; the fastest): the pulls land in ONE buffer per head that holds the rank's whole visible key
      sequence -- chunks entirely in the past in arrival order, the rank's own high chunk (the diagonal) last -- and a
      rank runs just two kernels, one per Q chunk, over that gathered sequence (`flash_attn_fwd_gathered`): the
      accumulators stay in tensor memory across every hop, no partial state is written, nothing is merged.  The
      kernels are launched at once; their TMA producer waits on per-chunk flags that the copy stream sets behind each
      pull (`GatheredKV`).
    - exchange="sendrecv": K/V chunk pairs rotate with NCCL send/recv on a side stream (also the path
      the gloo CPU test runs); needs `flash_attn_set_sm_margin` so that NCCL's kernel finds a free SM.
  Per-hop math is `flash_attn_fwd_ex`: every chunk
  pair writes its own partial state (O un-normalised fp32, m, l) in the format of the reference's
  split-K code (flash_attention.cu:460-496), and `flash_attn_merge` -- the reference's
  flash_attention_splitk_merge, flash_attention.cu:559-598 -- combines them into O at the end.  (A
  read-modify-write of one running state per pair costs +25 % per kernel; write-only costs +5 %.)

The compute callables are injectable so the schedule and the send/recv plumbing can be exercised on CPU
with gloo (tests/test_ring_cpu.py); the defaults are the CUDA library and nothing else.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple


def bh_shard(total_bh: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [start, start+count) of the B*H independent heads for `rank`."""
    base, rem = divmod(total_bh, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def zigzag_chunks(rank: int, world: int) -> Tuple[int, int]:
    """Chunk ids (of 2*world equal chunks) owned by `rank`: (r, 2P-1-r)."""
    return rank, 2 * world - 1 - rank


def hop_pairs(rank: int, src: int, world: int, causal: bool) -> List[Tuple[int, int, bool]]:
    """(q chunk slot, kv chunk slot, needs_diagonal_mask) pairs rank must compute while it holds
    src's K/V.  Slot 0 = the rank's low chunk, slot 1 = its high chunk.  Fully masked pairs are
    dropped here, on the host, so they cost nothing."""
    qa = zigzag_chunks(rank, world)
    kb = zigzag_chunks(src, world)
    out = []
    for qi, a in enumerate(qa):
        for ki, b in enumerate(kb):
            if not causal:
                out.append((qi, ki, False))
            elif b < a:
                out.append((qi, ki, False))      # K/V chunk entirely in the past: no mask
            elif b == a:
                out.append((qi, ki, True))       # diagonal chunk: causal mask inside
    return out


# SMs left to NCCL while a ring is running (see flash_attn_set_sm_margin); FLASH_ATTN_RING_SM_MARGIN overrides
DEFAULT_RING_SM_MARGIN = 16


_workspace: dict = {}     # see ring_attention_forward


def _cuda_partial(q, k, v, o_partial, ml, causal, q_offset, kv_offset, accumulate):
    from . import flash_attn_fwd_partial
    flash_attn_fwd_partial(q, k, v, o_partial, ml, causal, q_offset, kv_offset, accumulate)


def _cuda_finalize(o_partials, mls, out):
    from . import flash_attn_merge
    flash_attn_merge(o_partials, mls, out)


def pull_plan(rank: int, world: int, causal: bool) -> List[Tuple[int, List[int], List[Tuple[int, int, bool]]]]:
    """Hop by hop: (owner rank of the K/V used, K/V chunk slots that must be fetched from it, chunk pairs to
    compute).  Hop h uses rank (r - h) mod P, so at every hop the ranks read from P distinct owners: each
    NVSwitch port carries one outgoing and one incoming block.  Under a causal mask an owner ahead of us
    contributes only its low chunk (its high chunk is entirely in our future) -- half the bytes."""
    plan = []
    for hop in range(world):
        src = (rank - hop) % world
        pairs = hop_pairs(rank, src, world, causal)
        plan.append((src, sorted({ki for _, ki, _ in pairs}), pairs))
    return plan


class PeerKV:
    """This rank's K/V chunk pair in peer-readable device memory, the mapped blocks of all other ranks, and
    two local landing buffers.  Block layout: [k_lo, k_hi, v_lo, v_hi], each [B, H, C, D] fp16.

    Collective constructor (exchanges the 64-byte handles with all_gather_object).  `self.k` / `self.v` are
    ordinary torch views of the block: a caller that produces K/V directly into them pays no staging copy.
    """

    def __init__(self, B: int, H: int, C: int, D: int, device, group=None):
        import torch
        import torch.distributed as dist
        from . import peer_alloc, peer_open
        self.group, self.dev = group, torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.shape = (B, H, C, D)
        self.chunk_bytes = B * H * C * D * 2
        self.block, handle, self.ptr = peer_alloc(4 * self.chunk_bytes, self.dev)
        t = self.block.view(torch.float16).view(4, B, H, C, D)
        self.k, self.v = [t[0], t[1]], [t[2], t[3]]
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        self.peer_ptr = [self.ptr if r == self.rank else peer_open(handles[r]) for r in range(self.world)]
        self.land = [torch.empty((4, B, H, C, D), dtype=torch.float16, device=self.dev) for _ in range(2)]
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.comm = torch.cuda.Stream(device=self.dev)
        self.ready = {}       # hop -> event: landing buffer filled
        self.consumed = {}    # hop -> event: the kernels reading that hop's buffer have been enqueued

    # -- the five steps the driver calls (tests inject an object with the same methods) --
    def begin(self, k, v):
        """Make this rank's K/V readable (staging copy unless the caller already wrote into self.k / self.v),
        then: everyone's block is in place."""
        import torch
        import torch.distributed as dist
        for mine, given in zip(self.k + self.v, list(k) + list(v)):
            if mine.data_ptr() != given.data_ptr():
                mine.copy_(given)
        dist.all_reduce(self.flag, group=self.group)          # stream-ordered behind the copies above
        self.comm.wait_stream(torch.cuda.current_stream(self.dev))
        self.ready.clear()
        self.consumed.clear()
        return self.k + self.v

    def prefetch(self, hop: int, src: int, slots):
        """Start pulling the chunks `slots` of rank `src` for hop `hop` (>= 1) into landing buffer (hop-1)&1."""
        import torch
        from . import peer_copy
        land = self.land[(hop - 1) & 1]
        if hop - 2 in self.consumed:                          # that buffer's previous readers (hop - 2)
            self.comm.wait_event(self.consumed[hop - 2])
        cb = self.chunk_bytes
        if list(slots) == [0, 1]:
            spans = [(0, 4)]                                  # the whole block in one transfer
        else:
            spans = [(s, 1) for s in slots] + [(2 + s, 1) for s in slots]
        for first, n in spans:
            peer_copy(land.data_ptr() + first * cb, self.peer_ptr[src] + first * cb, n * cb, self.comm)
        ev = torch.cuda.Event()
        ev.record(self.comm)
        self.ready[hop] = ev

    def wait(self, hop: int):
        import torch
        torch.cuda.current_stream(self.dev).wait_event(self.ready[hop])
        land = self.land[(hop - 1) & 1]
        return [land[0], land[1], land[2], land[3]]

    def done(self, hop: int):
        import torch
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.consumed[hop] = ev

    def end(self):
        """Everyone is done reading everyone's block: K/V may be rewritten after this (stream-ordered)."""
        import torch
        import torch.distributed as dist
        torch.cuda.current_stream(self.dev).wait_stream(self.comm)
        dist.all_reduce(self.flag, group=self.group)

    def close(self):
        """Collective: unmap the peers' blocks, then free ours."""
        import torch
        import torch.distributed as dist
        from . import peer_close, peer_free
        torch.cuda.synchronize(self.dev)
        for r, ptr in enumerate(self.peer_ptr):
            if r != self.rank:
                peer_close(ptr)
        dist.barrier(group=self.group)                        # nobody still maps our block
        self.k = self.v = self.block = None
        peer_free(self.ptr)


_peer_kv: dict = {}


def peer_kv(B: int, H: int, C: int, D: int, device, group=None) -> PeerKV:
    """The cached PeerKV for this shape (collective on first use; one shape at a time)."""
    import torch
    key = (str(torch.device(device)), B, H, C, D, id(group))
    px = _peer_kv.get(key)
    if px is None:
        release_peer_kv()
        px = _peer_kv[key] = PeerKV(B, H, C, D, device, group)
    return px


def release_peer_kv() -> None:
    """Collective: drop the cached PeerKV / GatheredKV (call before destroying the process group)."""
    for px in _peer_kv.values():
        px.close()
    _peer_kv.clear()
    for gk in _gathered_kv.values():
        gk.close()
    _gathered_kv.clear()


# ---------------------------------------------------------------------------------------------------------------------
# exchange="gather": the rank's visible keys in one buffer, two kernels, no partial states
# ---------------------------------------------------------------------------------------------------------------------
def gather_layout(rank: int, world: int) -> List[Tuple[int, int]]:
    """Slots of rank `rank`'s gathered K/V buffer under a causal mask, as (owner rank, 0 = its low chunk / 1 = its high
    chunk).  Order = order of arrival with the staggered schedule (hop h pulls from rank r - h, so at every hop the P ranks
    read from P distinct owners), own chunks where the mask needs them:
        low chunks of ranks r-1, ..., 0      entirely in the past of both Q chunks          (hops 1 .. r)
        own low chunk                         diagonal of the low Q chunk, in the past of the high one
        low + high chunk of ranks P-1 .. r+1  in the past of the high Q chunk only           (hops r+1 .. P-1)
        own high chunk                        diagonal of the high Q chunk
    The low Q chunk attends to the first r + 1 slots, the high one to all 2P - r; in both cases every slot but the last is
    entirely visible, so the kernel's single causal offset describes the mask although the keys are not in sequence order."""
    lay = [(s, 0) for s in range(rank - 1, -1, -1)] + [(rank, 0)]
    for s in range(world - 1, rank, -1):
        lay += [(s, 0), (s, 1)]
    return lay + [(rank, 1)]


def gather_slot(buffer_rank: int, world: int, owner: int, which: int) -> int:
    """Slot of chunk (owner, which) in the gathered buffer of rank `buffer_rank`."""
    return gather_layout(buffer_rank, world).index((owner, which))


class GatheredKV:
    """Peer-readable gathered K/V of one rank: [K | V], each B*H heads of nslots*C rows, slot order = gather_layout.
    The rank's own chunks live in slots r and nslots-1 (`self.k` / `self.v` are strided views of them: a caller that produces
    K/V there pays no staging copy); the other ranks read them from there.  Collective constructor."""

    def __init__(self, B: int, H: int, C: int, D: int, device, group=None):
        import torch
        import torch.distributed as dist
        from . import peer_alloc, peer_open
        self.group, self.dev = group, torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.B, self.H, self.C, self.D = B, H, C, D
        self.layout = gather_layout(self.rank, self.world)
        self.nslots = len(self.layout)
        self.chunk_row_bytes = C * D * 2                         # one head of one chunk
        self.head_bytes = self.nslots * self.chunk_row_bytes     # one head of the gathered sequence
        self.tensor_bytes = B * H * self.head_bytes              # K (or V) region
        self.block, handle, self.ptr = peer_alloc(2 * self.tensor_bytes, self.dev)
        kv = self.block.view(torch.float16).view(2, B, H, self.nslots, C, D)
        own = (self.rank, self.nslots - 1)
        self.k = [kv[0, :, :, s] for s in own]                   # [B, H, C, D] views, head stride nslots*C*D
        self.v = [kv[1, :, :, s] for s in own]
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        self.peer_ptr = [self.ptr if r == self.rank else peer_open(handles[r]) for r in range(self.world)]
        # A chunk may land in GATHER_PARTS row blocks with a ready flag each (FLASH_ATTN_GATHER_PARTS), and odd / even hops
        # may pull on two copy streams side by side (FLASH_ATTN_GATHER_STREAMS=2).  Measured at 8 GPUs
        # (profiles/r02_c12_gather_parts.log): K and V rows of a chunk on two streams 16.4 ms against 15.7 ms on one, and
        # 16.6 / 16.8 ms with 2 / 4 blocks per chunk -- one block per chunk is the default.
        self.parts = GATHER_PARTS if C % (128 * GATHER_PARTS) == 0 else 1
        self.flags = torch.zeros(self.nslots * self.parts, dtype=torch.int32, device=self.dev)
        self.sync = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.comm = torch.cuda.Stream(device=self.dev)
        self.comm2 = torch.cuda.Stream(device=self.dev) if GATHER_STREAMS == 2 else self.comm
        self.side = torch.cuda.Stream(device=self.dev)           # second kernel stream (GATHER_OVERLAP)

    @property
    def k_ptr(self) -> int:
        return self.ptr

    @property
    def v_ptr(self) -> int:
        return self.ptr + self.tensor_bytes

    def begin(self, k, v):
        """Own chunks into their slots (unless the caller wrote them there), flags reset, 'every block is in place', then
        the pulls of all hops are queued on the copy stream, each followed by its chunks' ready flags."""
        import torch
        import torch.distributed as dist
        from . import peer_copy_2d, stream_write_flag
        cur = torch.cuda.current_stream(self.dev)
        for mine, given in zip(self.k + self.v, list(k) + list(v)):
            if mine.data_ptr() != given.data_ptr():
                mine.copy_(given)
        self.flags.zero_()
        nparts = self.parts
        self.flags.view(self.nslots, nparts)[self.rank] = 1
        self.flags.view(self.nslots, nparts)[self.nslots - 1] = 1
        dist.all_reduce(self.sync, group=self.group)             # stream-ordered behind the copies above
        self.comm.wait_stream(cur)
        if self.comm2 is not self.comm:
            self.comm2.wait_stream(cur)
        P, r = self.world, self.rank
        heads = self.B * self.H
        part_bytes = self.chunk_row_bytes // nparts
        for hop in range(1, P):
            src = (r - hop) % P
            src_head_bytes = (2 * P - src) * self.chunk_row_bytes
            src_tensor_bytes = heads * src_head_bytes
            stream = self.comm if hop % 2 == 1 else self.comm2    # two streams: odd and even hops pull side by side
            for which in ((0,) if src < r else (0, 1)):           # an owner behind us contributes only its low chunk
                dst_slot = self.layout.index((src, which))
                src_slot = src if which == 0 else 2 * P - src - 1
                for part in range(nparts):
                    for t in range(2):                             # K region, V region
                        peer_copy_2d(self.ptr + t * self.tensor_bytes + dst_slot * self.chunk_row_bytes + part * part_bytes,
                                     self.head_bytes,
                                     self.peer_ptr[src] + t * src_tensor_bytes + src_slot * self.chunk_row_bytes + part * part_bytes,
                                     src_head_bytes, part_bytes, heads, stream)
                    stream_write_flag(self.flags.data_ptr() + 4 * (dst_slot * nparts + part), 1, stream)

    def end(self):
        """Everyone is done reading everyone's block: K/V may be rewritten after this (stream-ordered)."""
        import torch
        import torch.distributed as dist
        torch.cuda.current_stream(self.dev).wait_stream(self.comm)
        if self.comm2 is not self.comm:
            torch.cuda.current_stream(self.dev).wait_stream(self.comm2)
        dist.all_reduce(self.sync, group=self.group)

    def close(self):
        import torch
        import torch.distributed as dist
        from . import peer_close, peer_free
        torch.cuda.synchronize(self.dev)
        for r, ptr in enumerate(self.peer_ptr):
            if r != self.rank:
                peer_close(ptr)
        dist.barrier(group=self.group)
        self.k = self.v = self.block = None
        peer_free(self.ptr)


# row blocks per chunk / copy streams of the gathered form (FLASH_ATTN_GATHER_PARTS, FLASH_ATTN_GATHER_STREAMS override)
import os as _os
GATHER_PARTS = int(_os.environ.get("FLASH_ATTN_GATHER_PARTS", "1"))
GATHER_STREAMS = 2 if _os.environ.get("FLASH_ATTN_GATHER_STREAMS") == "2" else 1
# the two kernels of a step on two streams (FLASH_ATTN_GATHER_OVERLAP=0 puts them back on one): the second grid fills SMs as the
# first one drains -- 15.2-15.4 ms against 15.7-15.9 ms per step at 8 GPUs (profiles/r02_final_8gpu_gather_overlap.log)
GATHER_OVERLAP = _os.environ.get("FLASH_ATTN_GATHER_OVERLAP", "1") == "1"

_gathered_kv: dict = {}


def gathered_kv(B: int, H: int, C: int, D: int, device, group=None) -> GatheredKV:
    """The cached GatheredKV for this shape (collective on first use; one shape at a time)."""
    import torch
    key = (str(torch.device(device)), B, H, C, D, id(group))
    gk = _gathered_kv.get(key)
    if gk is None:
        for old in _gathered_kv.values():
            old.close()
        _gathered_kv.clear()
        gk = _gathered_kv[key] = GatheredKV(B, H, C, D, device, group)
    return gk


def gather_attention_forward(q: Sequence, k: Sequence, v: Sequence, causal: bool, group=None, *, gathered=None):
    """Context-parallel forward over a gathered K/V buffer (module docstring); causal only.  Same arguments and result as
    `ring_attention_forward`; k / v may be the views `gathered_kv(...).k / .v` themselves."""
    import torch
    from . import flash_attn_fwd_gathered
    if not causal:
        raise ValueError("exchange='gather' serves the causal mask (every chunk is either entirely visible or the diagonal); "
                         "use exchange='pull' without one")
    B, H, C, D = q[0].shape
    for t in q:
        if t.dtype != torch.float16 or not t.is_contiguous() or t.shape != q[0].shape:
            raise TypeError("q chunks must be contiguous float16 [B, H, C, D] tensors of one shape")
    gk = gathered or gathered_kv(B, H, C, D, q[0].device, group)
    gk.begin(k, v)
    n = gk.nslots
    out = [torch.empty_like(q[0]), torch.empty_like(q[1])]
    # the low Q chunk first: it needs the slots that land first (the low chunks of the ranks behind us), so the pulls of the
    # later hops run under it; the high chunk's launch then finds most of its slots in place
    rr = C // gk.parts                 # rows per ready flag
    if GATHER_OVERLAP:
        # The two kernels are independent (different Q chunk, different output): the high chunk's launch goes to a second
        # stream so that its CTAs take over SM by SM as the low chunk's persistent CTAs run out of work, instead of waiting
        # for the whole grid to drain (work items here are 0.1 ms per slot: the tail of a launch is not small)
        cur = torch.cuda.current_stream(q[0].device)
        gk.side.wait_stream(cur)
        flash_attn_fwd_gathered(q[0], gk.k_ptr, gk.v_ptr, out[0], (gk.rank + 1) * C, n * C, True, gk.rank * C, gk.flags, rr)
        with torch.cuda.stream(gk.side):
            flash_attn_fwd_gathered(q[1], gk.k_ptr, gk.v_ptr, out[1], n * C, n * C, True, (n - 1) * C, gk.flags, rr)
        cur.wait_stream(gk.side)
    else:
        flash_attn_fwd_gathered(q[0], gk.k_ptr, gk.v_ptr, out[0], (gk.rank + 1) * C, n * C, True, gk.rank * C, gk.flags, rr)
        flash_attn_fwd_gathered(q[1], gk.k_ptr, gk.v_ptr, out[1], n * C, n * C, True, (n - 1) * C, gk.flags, rr)
    gk.end()
    return out


def _check_chunks(q, k, v):
    """The CUDA partial / merge path is FP16 only and reads plain contiguous [B, H, C, D] chunks (a slice of a
    [B, H, N, D] tensor along N is not one): refuse anything else instead of reading it with the wrong strides or format."""
    import torch
    if len(q) != 2 or len(k) != 2 or len(v) != 2:
        raise ValueError("q, k, v must each be a pair (low chunk, high chunk)")
    ref = q[0]
    for name, pair in (("q", q), ("k", k), ("v", v)):
        for t in pair:
            if t.shape != ref.shape or t.dim() != 4:
                raise ValueError(f"{name}: every chunk must be [B, H, C, D] of one shape")
            if not t.is_contiguous():
                raise ValueError(f"{name}: chunks must be contiguous [B, H, C, D] tensors")
            if t.device != ref.device or t.dtype != ref.dtype:
                raise ValueError(f"{name}: all chunks must share one device and dtype")
    if ref.is_cuda and ref.dtype != torch.float16:
        raise TypeError("context parallelism runs the FP16 partial/merge path; got " + str(ref.dtype))


def _partial_workspace(q, world, causal, schedule):
    """Partial states (one per chunk pair, write-only), cached per (device, shape): a step allocates nothing
    and zero-fills nothing."""
    import torch
    B, H, C, D = q[0].shape
    dev = q[0].device
    n_split = [max(1, sum(1 for hp in schedule for (qi, _, _) in hp if qi == i)) for i in range(2)]
    key = (str(dev), B, H, C, D, q[0].dtype, world, bool(causal), tuple(n_split))
    ws = _workspace.get(key)
    if ws is None:
        ws = {
            "o_part": [torch.empty((n_split[i], B * H * C, D), dtype=torch.float32, device=dev) for i in range(2)],
            "ml": [torch.empty((n_split[i], B * H * C, 2), dtype=torch.float32, device=dev) for i in range(2)],
        }
        _workspace.clear()          # one shape at a time: the buffers are large
        _workspace[key] = ws
    return ws


def _merge_partials(q, o_part, ml, used, finalize):
    import torch
    out = [torch.empty_like(q[0]), torch.empty_like(q[1])]
    for i in range(2):
        if used[i] == 0:           # a Q chunk that saw no key at all (cannot happen with the zig-zag layout)
            out[i].zero_()
        else:
            finalize(o_part[i][:used[i]], ml[i][:used[i]], out[i])
    return out


def pull_attention_forward(q: Sequence, k: Sequence, v: Sequence, causal: bool, group=None, *,
                           partial: Optional[Callable] = None, finalize: Optional[Callable] = None,
                           peer=None, rank: Optional[int] = None, world: Optional[int] = None):
    """Context-parallel forward with peer pulls (module docstring).  Same arguments and result as
    `ring_attention_forward`.  `peer` is a PeerKV (default: the cached one for this shape); tests inject an
    object with the same begin / prefetch / wait / done / end methods together with rank and world."""
    _check_chunks(q, k, v)
    partial = partial or _cuda_partial
    finalize = finalize or _cuda_finalize
    if rank is None or world is None:
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    B, H, C, D = q[0].shape
    if peer is None:
        peer = peer_kv(B, H, C, D, q[0].device, group)
    plan = pull_plan(rank, world, causal)
    ws = _partial_workspace(q, world, causal, [pairs for _, _, pairs in plan])
    o_part, ml = ws["o_part"], ws["ml"]
    used = [0, 0]
    qa = zigzag_chunks(rank, world)

    cur = peer.begin(k, v)
    for hop, (src, _, pairs) in enumerate(plan):
        if hop + 1 < world:
            nsrc, nslots, _ = plan[hop + 1]
            peer.prefetch(hop + 1, nsrc, nslots)       # lands while this hop's kernels run
        if hop > 0:
            cur = peer.wait(hop)
        kb = zigzag_chunks(src, world)
        for qi, ki, diag in pairs:
            partial(q[qi], cur[ki], cur[2 + ki], o_part[qi][used[qi]], ml[qi][used[qi]], bool(diag),
                    qa[qi] * C, kb[ki] * C, False)
            used[qi] += 1
        peer.done(hop)
    peer.end()
    return _merge_partials(q, o_part, ml, used, finalize)


def ring_attention_forward(q: Sequence, k: Sequence, v: Sequence, causal: bool, group=None, *,
                           partial: Optional[Callable] = None, finalize: Optional[Callable] = None,
                           comm_stream=None, exchange: Optional[str] = None):
    """Context-parallel forward.  q, k, v: pairs (low chunk, high chunk) of contiguous [B, H, C, D] tensors in
    the zig-zag layout of `zigzag_chunks`.  Returns the pair of output chunks, same layout as q.

    exchange = "gather" (CUDA default under a causal mask), "pull" (CUDA default without one) or "sendrecv" (module
    docstring); FLASH_ATTN_RING_EXCHANGE overrides the default.
    sendrecv, each hop: post isend/irecv of the K/V pair for the next hop, run the (at most four, causal: two)
    chunk-pair kernels of this hop, wait for the transfer, swap buffers."""
    import os
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if exchange is None:
        # CUDA default: the gathered form under a causal mask (no partial states, no merge), peer pulls with one partial
        # state per chunk pair without one (the gathered layout relies on every slot but the last being entirely visible)
        on_gpus = q[0].is_cuda and world > 1
        plain = partial is None and finalize is None          # caller-supplied partial / finalize hooks go with pull / sendrecv
        exchange = os.environ.get("FLASH_ATTN_RING_EXCHANGE") or \
            ("sendrecv" if not on_gpus else "gather" if causal and plain else "pull")
    if exchange == "gather":               # K/V may be the strided views of the gathered buffer itself
        return gather_attention_forward(q, k, v, causal, group)
    _check_chunks(q, k, v)
    if exchange not in ("pull", "sendrecv"):
        raise ValueError(f"exchange must be 'pull', 'gather' or 'sendrecv', not {exchange!r}")
    if exchange == "pull":
        return pull_attention_forward(q, k, v, causal, group, partial=partial, finalize=finalize)
    partial = partial or _cuda_partial
    finalize = finalize or _cuda_finalize
    B, H, C, D = q[0].shape
    dev = q[0].device
    on_cuda = dev.type == "cuda"

    # Which chunk pairs does this rank compute, hop by hop?  (deterministic: every rank can enumerate it)
    schedule = [hop_pairs(rank, (rank - hop) % world, world, causal) for hop in range(world)]
    ws = _partial_workspace(q, world, causal, schedule)
    if world > 1 and "recv" not in ws:     # two receive sets, cached with the partial states
        ws["recv"] = [[torch.empty_like(k[0]) for _ in range(4)] for _ in range(2)]
    o_part, ml = ws["o_part"], ws["ml"]
    used = [0, 0]          # partial states written so far, per Q chunk

    cur = [k[0], k[1], v[0], v[1]]
    # two receive sets: hop h receives into set h & 1 while set (h - 1) & 1 (or the caller's k, v) is read
    nxt = ws["recv"][0] if world > 1 else None
    send_to = (rank + 1) % world
    recv_from = (rank - 1) % world
    if on_cuda and comm_stream is None and world > 1:
        comm_stream = torch.cuda.Stream(device=dev)
    old_margin = None
    if on_cuda and world > 1 and partial is _cuda_partial:
        import os
        from . import set_sm_margin
        old_margin = set_sm_margin(int(os.environ.get("FLASH_ATTN_RING_SM_MARGIN", DEFAULT_RING_SM_MARGIN)))

    qa = zigzag_chunks(rank, world)
    try:
        for hop in range(world):
            src = (rank - hop) % world
            reqs = []
            if hop + 1 < world:
                ops = []
                for t_send, t_recv in zip(cur, nxt):
                    ops.append(dist.P2POp(dist.isend, t_send, send_to, group))
                    ops.append(dist.P2POp(dist.irecv, t_recv, recv_from, group))
                if on_cuda:
                    comm_stream.wait_stream(torch.cuda.current_stream(dev))   # cur is ready to be read
                    with torch.cuda.stream(comm_stream):
                        reqs = dist.batch_isend_irecv(ops)
                else:
                    reqs = dist.batch_isend_irecv(ops)
            kb = zigzag_chunks(src, world)
            for qi, ki, diag in schedule[hop]:
                partial(q[qi], cur[ki], cur[2 + ki], o_part[qi][used[qi]], ml[qi][used[qi]], bool(diag),
                        qa[qi] * C, kb[ki] * C, False)
                used[qi] += 1
            if hop + 1 < world:
                for r in reqs:
                    r.wait()
                if on_cuda:
                    torch.cuda.current_stream(dev).wait_stream(comm_stream)
                    comm_stream.wait_stream(torch.cuda.current_stream(dev))   # kernels that read cur are ordered first
                cur, nxt = nxt, ws["recv"][(hop + 1) & 1]
    finally:
        # the margin is process-global: an exception in a hop must not leave every later launch on fewer SMs
        if old_margin is not None:
            from . import set_sm_margin
            set_sm_margin(old_margin)
    return _merge_partials(q, o_part, ml, used, finalize)

"""Multi-GPU drivers for the forward path: one process per GPU over torch.distributed.

* batch x heads sharding (config 3): every (b, h) is an independent problem (reference
  flash_attention.cu:120-122), so ranks take contiguous slices of B*H and never communicate.
* ring context parallelism (config 5): the sequence is cut into 2P chunks, rank r owns chunks r and
  2P-1-r (zig-zag: every hop costs every rank exactly two unmasked chunk pairs under a causal mask);
  Q and the partial state stay, K/V chunk pairs travel around the ring with send/recv on a side
  stream while the current pair is being consumed.  Per-hop math is `flash_attn_fwd_ex`: every chunk
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


def ring_attention_forward(q: Sequence, k: Sequence, v: Sequence, causal: bool, group=None, *,
                           partial: Optional[Callable] = None, finalize: Optional[Callable] = None,
                           comm_stream=None):
    """Ring-CP forward.  q, k, v: pairs (low chunk, high chunk) of contiguous [B, H, C, D] tensors in the
    zig-zag layout of `zigzag_chunks`.  Returns the pair of output chunks, same layout as q.

    Each hop: post isend/irecv of the K/V pair for the next hop, run the (at most four, causal: two)
    chunk-pair kernels of this hop, wait for the transfer, swap buffers."""
    import torch
    import torch.distributed as dist

    partial = partial or _cuda_partial
    finalize = finalize or _cuda_finalize
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B, H, C, D = q[0].shape
    dev = q[0].device
    on_cuda = dev.type == "cuda"

    # Which chunk pairs does this rank compute, hop by hop?  (deterministic: every rank can enumerate it)
    schedule = [hop_pairs(rank, (rank - hop) % world, world, causal) for hop in range(world)]
    n_split = [max(1, sum(1 for hp in schedule for (qi, _, _) in hp if qi == i)) for i in range(2)]
    # partial states (one per chunk pair, write-only) and receive buffers are cached per (device, shape):
    # a ring step allocates nothing and zero-fills nothing
    key = (str(dev), B, H, C, D, q[0].dtype, world, bool(causal))
    ws = _workspace.get(key)
    if ws is None:
        ws = {
            "o_part": [torch.empty((n_split[i], B * H * C, D), dtype=torch.float32, device=dev) for i in range(2)],
            "ml": [torch.empty((n_split[i], B * H * C, 2), dtype=torch.float32, device=dev) for i in range(2)],
            "recv": [[torch.empty_like(k[0]) for _ in range(4)] for _ in range(2)] if world > 1 else None,
        }
        _workspace.clear()          # one shape at a time: the buffers are large
        _workspace[key] = ws
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
    if old_margin is not None:
        from . import set_sm_margin
        set_sm_margin(old_margin)
    out = [torch.empty_like(q[0]), torch.empty_like(q[1])]
    for i in range(2):
        if used[i] == 0:           # a Q chunk that saw no key at all (cannot happen with the zig-zag layout)
            out[i].zero_()
        else:
            finalize(o_part[i][:used[i]], ml[i][:used[i]], out[i])
    return out

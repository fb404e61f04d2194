"""CPU-only, world_size 2 over gloo: the multi-GPU host logic of flash_attention_cuda_b200/ring.py.

The ring driver is run for real (zig-zag ownership, hop schedule, double-buffered send/recv of K/V
chunk pairs, one partial state per chunk pair, final merge); only the per-hop math is injected as a
numpy stand-in with the semantics of flash_attn_fwd_ex, so this needs no GPU.  The result is gated
against the monolithic CPU oracle."""
import os
import socket
import sys

import numpy as np
import pytest

import _oracle

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from flash_attention_cuda_b200 import ring  # noqa: E402

FLT_MAX = np.finfo(np.float32).max


def np_partial(q, k, v, o_part, ml, causal, q_off, kv_off, accumulate):
    """numpy stand-in for flash_attn_fwd_ex: partial state (O un-normalised, m, l), FA.cu:460-496 format,
    merged in place with the algebra of FA.cu:575-597 when `accumulate`."""
    B, H, C, D = q.shape
    Ck = k.shape[2]
    qf, kf, vf = (t.float().numpy().reshape(B * H, -1, D) for t in (q, k, v))
    scale = np.float32(1.0 / np.sqrt(D))
    rows = np.arange(C)[:, None] + q_off
    cols = np.arange(Ck)[None, :] + kv_off
    for bh in range(B * H):
        s = (qf[bh] @ kf[bh].T) * scale
        if causal:
            s = np.where(cols <= rows, s, -np.inf)
        m = s.max(axis=1)
        m_safe = np.where(np.isinf(m), 0.0, m)
        p = np.exp(s - m_safe[:, None])
        l_new = p.sum(axis=1).astype(np.float32)
        o_new = (p.astype(np.float16).astype(np.float32) @ vf[bh]).astype(np.float32)
        m_new = np.where(np.isinf(m), -FLT_MAX, m).astype(np.float32)
        sl = slice(bh * C, (bh + 1) * C)
        op = o_part[sl].numpy()
        mlv = ml[sl].numpy()
        if accumulate:
            m_old, l_old = mlv[:, 0].copy(), mlv[:, 1].copy()
            m_max = np.maximum(m_old, m_new)
            w_old = np.where(m_old <= -FLT_MAX, 0.0, np.exp(m_old - m_max)).astype(np.float32)
            w_new = np.where(m_new <= -FLT_MAX, 0.0, np.exp(m_new - m_max)).astype(np.float32)
            op[:] = op * w_old[:, None] + o_new * w_new[:, None]
            mlv[:, 0] = m_max
            mlv[:, 1] = l_old * w_old + l_new * w_new
        else:
            op[:] = o_new
            mlv[:, 0] = m_new
            mlv[:, 1] = l_new


def np_finalize(o_parts, mls, out):
    """numpy stand-in for flash_attn_merge (FA.cu:559-598): o_parts [S, rows, D], mls [S, rows, 2]."""
    op, m, l = o_parts.numpy(), mls[:, :, 0].numpy(), mls[:, :, 1].numpy()
    m_max = m.max(axis=0)
    w = np.where(m <= -FLT_MAX, 0.0, np.exp(m - m_max[None, :])).astype(np.float32)
    lsum = (w * l).sum(axis=0)
    acc = (w[:, :, None] * op).sum(axis=0)
    res = np.where(lsum[:, None] > 0, acc / np.maximum(lsum[:, None], 1e-30), 0.0).astype(np.float32)
    out.copy_(torch.from_numpy(res).reshape(out.shape).half())


def _worker(rank, world, port, causal, N, D, H, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)     # same global tensors on every rank
        q, k, v = (rng.standard_normal((1, H, N, D), dtype=np.float32).astype(np.float16) for _ in range(3))
        v = (v.astype(np.float32) * 0.5).astype(np.float16)
        C = N // (2 * world)
        lo, hi = ring.zigzag_chunks(rank, world)

        def chunks(x):
            t = torch.from_numpy(x)
            return [t[:, :, c * C:(c + 1) * C].contiguous() for c in (lo, hi)]

        out = ring.ring_attention_forward(chunks(q), chunks(k), chunks(v), causal,
                                          partial=np_partial, finalize=np_finalize)
        ref = _oracle.attention(q, k, v, causal)
        worst = (0.0, 0.0)
        for o, c in zip(out, (lo, hi)):
            mx, mean = _oracle.diff(o.numpy(), ref[:, :, c * C:(c + 1) * C])
            worst = (max(worst[0], mx), max(worst[1], mean))
        ret[rank] = worst
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("causal", [True, False])
def test_ring_cp_two_ranks_gloo(causal):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), causal, 128, 64, 2, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        mx, mean = ret[r]
        assert mx <= 2e-3 and mean <= 2e-4, (r, mx, mean)


def test_zigzag_hop_schedule_is_balanced_and_complete():
    for world in (2, 4, 8):
        covered = set()
        for rank in range(world):
            qa = ring.zigzag_chunks(rank, world)
            for hop in range(world):
                src = (rank - hop) % world
                kb = ring.zigzag_chunks(src, world)
                pairs = ring.hop_pairs(rank, src, world, causal=True)
                # every hop costs every rank two chunk pairs (diagonal pairs count half each)
                cost = sum(0.5 if d else 1.0 for _, _, d in pairs)
                assert cost == (2.0 if hop else 2.0), (world, rank, hop, pairs)
                for qi, ki, d in pairs:
                    covered.add((qa[qi], kb[ki]))
                    assert (qa[qi] == kb[ki]) == d
        n = 2 * world
        assert covered == {(a, b) for a in range(n) for b in range(n) if b <= a}


def test_bh_shard_partitions_heads():
    for total, world in [(512, 8), (512, 4), (32, 8), (10, 4), (3, 8)]:
        spans = [ring.bh_shard(total, r, world) for r in range(world)]
        assert sum(c for _, c in spans) == total
        pos = 0
        for s, c in spans:
            assert s == pos
            pos += c
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


class FakePeer:
    """Stand-in for ring.PeerKV: `blocks` plays the peer-readable memory of all ranks.  A pull copies ONLY the
    requested chunks; the rest of the landing buffer is NaN, so a kernel reading a chunk the plan did not
    fetch poisons the output.  Also checks the call protocol (prefetch one hop ahead, double buffer)."""

    def __init__(self, blocks, rank):
        self.blocks, self.rank = blocks, rank
        self.land = [None, None]
        self.land_hop = [None, None]
        self.finished = set()
        self.log = []

    def begin(self, k, v):
        assert [t.data_ptr() for t in self.blocks[self.rank]] == [t.data_ptr() for t in list(k) + list(v)]
        self.log.append("begin")
        return list(k) + list(v)

    def prefetch(self, hop, src, slots):
        b = (hop - 1) & 1
        # the buffer's previous readers (hop - 2) must have been enqueued before it is overwritten
        assert self.land_hop[b] is None or self.land_hop[b] in self.finished
        full = [torch.full_like(t, float("nan")) for t in self.blocks[src]]
        for s in slots:
            full[s] = self.blocks[src][s].clone()
            full[2 + s] = self.blocks[src][2 + s].clone()
        self.land[b], self.land_hop[b] = full, hop
        self.log.append(("prefetch", hop, src, tuple(slots)))

    def wait(self, hop):
        b = (hop - 1) & 1
        assert self.land_hop[b] == hop
        return self.land[b]

    def done(self, hop):
        self.finished.add(hop)

    def end(self):
        self.log.append("end")


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("causal", [True, False])
def test_pull_cp_all_ranks_simulated(world, causal):
    """The pull driver has no rank-to-rank dependency inside a step, so all ranks can be run one after the
    other in one process against a shared table of K/V blocks."""
    N, D, H = 64 * world, 64, 2
    rng = np.random.default_rng(11)
    q, k, v = (rng.standard_normal((1, H, N, D), dtype=np.float32).astype(np.float16) for _ in range(3))
    v = (v.astype(np.float32) * 0.5).astype(np.float16)
    C = N // (2 * world)

    def chunks(x, rank):
        t = torch.from_numpy(x)
        return [t[:, :, c * C:(c + 1) * C].contiguous() for c in ring.zigzag_chunks(rank, world)]

    blocks = {r: chunks(k, r) + chunks(v, r) for r in range(world)}
    ref = _oracle.attention(q, k, v, causal)
    for rank in range(world):
        peer = FakePeer(blocks, rank)
        out = ring.pull_attention_forward(chunks(q, rank), blocks[rank][:2], blocks[rank][2:], causal,
                                          partial=np_partial, finalize=np_finalize, peer=peer, rank=rank, world=world)
        assert peer.log[0] == "begin" and peer.log[-1] == "end"
        assert [e[1] for e in peer.log[1:-1]] == list(range(1, world))
        for o, c in zip(out, ring.zigzag_chunks(rank, world)):
            mx, mean = _oracle.diff(o.numpy(), ref[:, :, c * C:(c + 1) * C])
            assert mx <= 2e-3 and mean <= 2e-4, (rank, c, mx, mean)


def test_pull_plan_reads_distinct_owners_and_skips_masked_chunks():
    for world in (2, 4, 8):
        plans = [ring.pull_plan(r, world, True) for r in range(world)]
        for hop in range(world):
            # one reader per owner at every hop: no NVSwitch port is asked for two blocks at once
            assert sorted(plans[r][hop][0] for r in range(world)) == list(range(world))
        for r in range(world):
            fetched = 0
            for hop, (src, slots, pairs) in enumerate(plans[r]):
                assert src == (r - hop) % world
                assert set(slots) == {ki for _, ki, _ in pairs}
                if hop:
                    fetched += len(slots)
                    # an owner ahead of us in the sequence contributes its low chunk only
                    assert slots == ([0] if src < r else [0, 1])
            assert fetched == r + 2 * (world - 1 - r)
        full = ring.pull_plan(0, world, False)
        assert all(slots == [0, 1] and len(pairs) == 4 for _, slots, pairs in full)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_gathered_layout_gives_each_q_chunk_its_visible_keys_with_the_diagonal_last(world):
    """exchange="gather": slot order of every rank's gathered K/V buffer (ring.gather_layout)."""
    for r in range(world):
        lay = ring.gather_layout(r, world)
        ids = [o if w == 0 else 2 * world - 1 - o for o, w in lay]          # chunk ids in sequence order
        lo, hi = ring.zigzag_chunks(r, world)
        assert len(lay) == 2 * world - r and len(set(lay)) == len(lay)
        assert ids[r] == lo and ids[-1] == hi                               # the diagonals close the two visible prefixes
        assert sorted(ids[:r + 1]) == list(range(lo + 1))                   # low Q chunk: chunks 0 .. lo, nothing else
        assert sorted(ids) == list(range(hi + 1))                           # high Q chunk: chunks 0 .. hi
        # arrival order: hop h brings the chunks of rank r - h; P distinct owners are read at every hop
        arrivals = [o for o, w in lay if o != r]
        hops = [(r - o) % world for o in arrivals]
        assert hops == sorted(hops)
        for o, w in lay:
            assert ring.gather_slot(r, world, o, w) == lay.index((o, w))
    for hop in range(1, world):
        assert sorted((r - hop) % world for r in range(world)) == list(range(world))

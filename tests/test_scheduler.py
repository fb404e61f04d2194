"""CPU-only: the persistent kernel's work decomposition (host mirror of fa::decode_work, which
replaces the reference's blockIdx mapping / GRID_SWAP, flash_attention.cu:103-112): every
(head, q-tile) exactly once, heavy-first inside a head, fully masked KV tiles skipped."""
import pytest

import flash_attention_cuda_b200 as fa


@pytest.fixture(scope="module", autouse=True)
def _built():
    fa.build()


def items(B, H, Nq, Nkv, D, causal, shift=0):
    total = fa.work_item(0, B, H, Nq, Nkv, D, causal, shift)["total"]
    return [fa.work_item(w, B, H, Nq, Nkv, D, causal, shift) for w in range(total)]


@pytest.mark.parametrize("N", [1, 127, 128, 129, 256, 257, 1000, 1024, 4096])
@pytest.mark.parametrize("causal", [False, True])
def test_every_q_tile_exactly_once(N, causal):
    B, H = 2, 3
    its = items(B, H, N, N, 128, causal)
    nq_tiles = (N + 127) // 128
    seen = set()
    for it in its:
        for t in range(2):
            q_start = it["q0"] + 128 * t
            n = it["n1"] if t else it["n0"]
            if q_start < N:
                assert (it["bh"], q_start) not in seen
                seen.add((it["bh"], q_start))
                last_row = min(q_start + 127, N - 1)
                # trip counts are in 64-wide KV sub-tiles (fa::kSubN)
                want = (min(N, last_row + 1) + 63) // 64 if causal else (N + 63) // 64
                assert n == want, (it, t)
            else:
                assert n == 0
    assert len(seen) == B * H * nq_tiles


def test_heavy_first_within_head_and_heads_outermost():
    its = items(1, 4, 2048, 2048, 128, True)
    bhs = [it["bh"] for it in its]
    assert bhs == sorted(bhs)
    for bh in range(4):
        w = [it["n1"] + it["n0"] for it in its if it["bh"] == bh]
        assert w == sorted(w, reverse=True)


def test_masked_tiles_skipped_with_offsets():
    # ring hop where the K/V block lies entirely in the future of the queries: nothing to do
    its = items(1, 1, 512, 512, 128, True, shift=-512)
    assert all(it["n0"] == 0 and it["n1"] == 0 for it in its)
    # block entirely in the past: every tile needs all KV tiles, as in non-causal
    its = items(1, 1, 512, 512, 128, True, shift=512)
    assert all(it["n0"] == 8 and it["n1"] == 8 for it in its)


def test_total_causal_tiles_is_triangular():
    N = 8192
    its = items(1, 1, N, N, 128, True)
    tiles = sum(it["n0"] + it["n1"] for it in its)
    n = N // 128          # q tiles; q tile i needs 2*(i+1) sub-tiles of 64 keys
    assert tiles == n * (n + 1)

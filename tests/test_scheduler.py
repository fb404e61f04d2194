"""CPU-only: the persistent kernel's work decomposition (a work item = one head x tiles_per_item 128-row Q tiles) (host mirror of fa::decode_work, which
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
    cg = fa.tiles_per_item(128)
    seen = set()
    for it in its:
        if cg == 1:
            assert it["n1"] == 0
        for t in range(cg):
            q_start = it["q0"] + 128 * t
            n = it["n1"] if t else it["n0"]
            if q_start < N:
                assert (it["bh"], q_start) not in seen
                seen.add((it["bh"], q_start))
                last_row = min(q_start + 127, N - 1)
                want = (min(N, last_row + 1) + 127) // 128 if causal else (N + 127) // 128
                assert n == want, (it, t)
            else:
                assert n == 0
    assert len(seen) == B * H * nq_tiles


def test_heavy_first_within_l2_sized_head_groups():
    # N=8192 D=128: K+V of a head = 4 MB -> 8 heads per 32 MB group; inside a group heavy-first across heads
    its = items(1, 32, 8192, 8192, 128, True)
    nqp = 64 // fa.tiles_per_item(128)
    per_group = 8 * nqp
    assert len(its) == 32 * nqp
    for g in range(4):
        grp = its[g * per_group:(g + 1) * per_group]
        assert {it["bh"] for it in grp} == set(range(8 * g, 8 * g + 8))
        w = [it["n"] for it in grp]
        assert w == sorted(w, reverse=True)
    # the launch ends with the lightest items
    assert its[-1]["n"] == min(it["n"] for it in its)


def test_short_sequences_form_one_group():
    its = items(1, 4, 2048, 2048, 128, True)
    w = [it["n"] for it in its]
    assert w == sorted(w, reverse=True)          # pure heavy-first: all heads fit one group
    assert [it["bh"] for it in its[:4]] == [0, 1, 2, 3]


def test_last_group_may_be_smaller():
    # 5 heads of 16 MB K/V each (N=32768): 32 MB groups of 2 + 2 + 1, every (head, pair) still exactly once
    its = items(1, 5, 32768, 32768, 128, True)
    seen = {(it["bh"], it["q0"]) for it in its}
    units = 256 // fa.tiles_per_item(128)
    assert len(seen) == len(its) == 5 * units
    assert [it["bh"] for it in its[:4]] == [0, 1, 0, 1] and its[2 * units]["bh"] == 2 and its[4 * units]["bh"] == 4


def test_masked_tiles_skipped_with_offsets():
    # ring hop where the K/V block lies entirely in the future of the queries: nothing to do
    its = items(1, 1, 512, 512, 128, True, shift=-512)
    assert all(it["n0"] == 0 and it["n1"] == 0 for it in its)
    # block entirely in the past: every tile needs all KV tiles, as in non-causal
    its = items(1, 1, 512, 512, 128, True, shift=512)
    assert all(it["n"] == 4 and it["n0"] == 4 for it in its)


def test_total_causal_tiles_is_triangular():
    N = 8192
    its = items(1, 1, N, N, 128, True)
    tiles = sum(it["n0"] + it["n1"] for it in its)
    n = N // 128
    assert tiles == n * (n + 1) // 2


def test_head_groups_are_equal_sized():
    # 20 heads of 4 MB: 8 per 32 MB group would leave a last group of 4 -> three groups of 7, 7, 6 instead
    its = items(1, 20, 8192, 8192, 128, True)
    nqp = 64 // fa.tiles_per_item(128)
    assert len(its) == 20 * nqp
    bounds = [0, 7 * nqp, 14 * nqp, 20 * nqp]
    heads = [sorted({it["bh"] for it in its[a:b]}) for a, b in zip(bounds, bounds[1:])]
    assert heads == [list(range(0, 7)), list(range(7, 14)), list(range(14, 20))]


@pytest.mark.parametrize("N", [1, 127, 128, 129, 256, 257, 1000, 1024, 2048])
@pytest.mark.parametrize("causal", [False, True])
def test_split_mode_covers_every_q_tile_and_every_kv_tile_once(N, causal):
    """Split mode (short sequences): an item is ONE Q tile; slot 0 takes its KV tiles 0, 2, 4, ... and slot 1 the odd ones."""
    B, H = 2, 3
    total = fa.work_item(0, B, H, N, N, 128, causal, split=True)["total"]
    its = [fa.work_item(w, B, H, N, N, 128, causal, split=True) for w in range(total)]
    nq_tiles = (N + 127) // 128
    assert total == B * H * nq_tiles
    seen = set()
    for it in its:
        assert (it["bh"], it["q0"]) not in seen and it["q0"] % 128 == 0 and it["q0"] < N
        seen.add((it["bh"], it["q0"]))
        last_row = min(it["q0"] + 127, N - 1)
        want = (min(N, last_row + 1) + 127) // 128 if causal else (N + 127) // 128
        assert it["n0"] + it["n1"] == want and it["n0"] == (want + 1) // 2 and it["n1"] == want // 2
    assert len(seen) == total
    w = [it["n"] for it in its]
    assert w == sorted(w, reverse=True)      # heavy-first (one L2 group at these sizes)

"""CPU tests of the host-side logic of flash_attn_fwd_host (the reference harness's H2D -> dispatch -> D2H sequence,
flash_attention.cu:771-780, cut into pipelined head chunks): the chunk boundaries, through the library's test hook.
No GPU is touched."""
import ctypes

import pytest

import flash_attention_cuda_b200 as fa

MiB = 1 << 20


def bounds(BH, bytes_per_tensor, want=8, taper=1, first=0):
    L = fa.lib()
    L.flash_attn_debug_host_chunks.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.POINTER(ctypes.c_int)]
    L.flash_attn_debug_host_chunks.restype = ctypes.c_int
    out = (ctypes.c_int * 33)()
    n = L.flash_attn_debug_host_chunks(BH, bytes_per_tensor, want, taper, first, out)
    assert n >= 1, fa.lib().flash_attn_error_string(n)
    return list(out[:n + 1])


@pytest.mark.parametrize("taper", [0, 1])
@pytest.mark.parametrize("first", [0, 1, 3, 50])
@pytest.mark.parametrize("want", [1, 4, 8, 12, 32])
def test_every_head_in_exactly_one_chunk_and_no_empty_chunk(taper, first, want):
    for BH in (1, 2, 3, 7, 8, 9, 31, 32, 33, 100, 512, 4096):
        for per_head in (64 * 1024, 2 * MiB, 32 * MiB):       # N*D*2 bytes of one head
            b = bounds(BH, BH * per_head, want, taper, first)
            assert b[0] == 0 and b[-1] == BH
            assert all(b[i] < b[i + 1] for i in range(len(b) - 1)), (BH, per_head, b)
            n = len(b) - 1
            assert n <= want and n <= BH
            # a chunk carries >= 16 MiB of Q+K+V unless the whole call is one chunk
            assert n == 1 or 3 * BH * per_head >= n * 16 * MiB


def test_headline_shape_has_eight_chunks_that_shrink():
    # B1 H32 N8192 D128: 64 MiB per tensor
    b = bounds(32, 64 * MiB)
    sizes = [b[i + 1] - b[i] for i in range(len(b) - 1)]
    assert sizes == [6, 5, 5, 4, 4, 3, 3, 2]
    assert bounds(32, 64 * MiB, taper=0) == list(range(0, 33, 4))
    # a small first chunk (A/B switch)
    s3 = bounds(32, 64 * MiB, first=3)
    sizes3 = [s3[i + 1] - s3[i] for i in range(len(s3) - 1)]
    assert sizes3[0] < sizes[0] and sum(sizes3) == 32 and len(sizes3) == 8


def test_small_calls_are_one_chunk():
    assert bounds(2, 2 * 384 * 128 * 2) == [0, 2]          # the smoke shape
    assert bounds(32, 8 * MiB) == [0, 32]                  # BASELINE config 1: 24 MiB of input -> one chunk
    assert len(bounds(16, 16 * MiB)) - 1 == 3              # 48 MiB of input -> three chunks


def test_bad_arguments():
    L = fa.lib()
    out = (ctypes.c_int * 33)()
    L.flash_attn_debug_host_chunks.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.POINTER(ctypes.c_int)]
    L.flash_attn_debug_host_chunks.restype = ctypes.c_int
    assert L.flash_attn_debug_host_chunks(0, 1, 8, 1, 0, out) < 0
    assert L.flash_attn_debug_host_chunks(1, 0, 8, 1, 0, out) < 0
    assert L.flash_attn_debug_host_chunks(1, 1, 0, 1, 0, out) < 0
    assert L.flash_attn_debug_host_chunks(1, 1, 8, 1, 0, None) < 0

"""In-process A/B of flash_attn_fwd_host variants (the call's time is bimodal from process to process, so variants are
alternated inside ONE process): O copied back by a copy engine / stored by the kernel into the pinned buffer, and the
weight of the first head chunk.   python tests/harness/host_ab.py [rounds]"""
import ctypes
import os
import statistics
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import flash_attention_cuda_b200 as fa   # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B, H, N, D = 1, 32, 8192, 128
L = fa.lib()
L.flash_attn_debug_set_host_zerocopy.argtypes = [ctypes.c_int]
L.flash_attn_debug_set_host_first.argtypes = [ctypes.c_int]
hq, hk, hv = ((torch.rand((B, H, N, D)) - 0.5).half().pin_memory() for _ in range(3))
ho = torch.empty((B, H, N, D), dtype=torch.float16).pin_memory()
variants = [("staged first=0", 0, 0), ("zerocopy first=0", 1, 0), ("zerocopy first=3", 1, 3), ("zerocopy first=5", 1, 5),
            ("staged first=3", 0, 3)]
res = {v[0]: [] for v in variants}


def call():
    fa.check(L.flash_attn_fwd_host(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), ho.data_ptr(), B, H, N, D, 1))


for rnd in range(rounds + 1):
    for name, zc, first in variants:
        L.flash_attn_debug_set_host_zerocopy(zc)
        L.flash_attn_debug_set_host_first(first)
        call()
        t0 = time.perf_counter()
        for _ in range(8):
            call()
        if rnd:                                   # round 0 warms up descriptors, streams and clocks
            res[name].append((time.perf_counter() - t0) / 8 * 1e3)
for name, v in res.items():
    print(f"{name:18s}: median {statistics.median(v):.3f} ms  min {min(v):.3f}  max {max(v):.3f}   ({len(v)} rounds of 8 calls)")

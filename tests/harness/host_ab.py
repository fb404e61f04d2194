"""In-process A/B of flash_attn_fwd_host variants (the call's time is bimodal from process to process, so variants are
alternated inside ONE process): O copied back by a copy engine / stored by the kernel into the pinned buffer, the
weight of the first head chunk, and the width of the chunk kernels' grid under the direct store.
   python tests/harness/host_ab.py [rounds] [variant-set: first | ctas]"""
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
which = sys.argv[2] if len(sys.argv) > 2 else "ctas"
B, H, N, D = 1, 32, 8192, 128
L = fa.lib()
for name in ("flash_attn_debug_set_host_zerocopy", "flash_attn_debug_set_host_first", "flash_attn_debug_set_host_ctas"):
    getattr(L, name).argtypes = [ctypes.c_int]
    getattr(L, name).restype = None
hq, hk, hv = ((torch.rand((B, H, N, D)) - 0.5).half().pin_memory() for _ in range(3))
ho = torch.empty((B, H, N, D), dtype=torch.float16).pin_memory()
# (name, zerocopy, first-chunk weight, CTAs per chunk kernel)
if which == "first":
    variants = [("staged first=0", 0, 0, 0), ("zerocopy first=0", 1, 0, 0), ("zerocopy first=3", 1, 3, 0),
                ("zerocopy first=5", 1, 5, 0), ("staged first=3", 0, 3, 0)]
else:
    variants = [("staged", 0, 0, 0), ("zerocopy ctas=all", 1, 0, 0)] + [(f"zerocopy ctas={n}", 1, 0, n) for n in (64, 40, 28, 20, 14)]
res = {v[0]: [] for v in variants}


def call():
    fa.check(L.flash_attn_fwd_host(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), ho.data_ptr(), B, H, N, D, 1))


ref = None
for rnd in range(rounds + 1):
    for name, zc, first, ctas in variants:
        L.flash_attn_debug_set_host_zerocopy(zc)
        L.flash_attn_debug_set_host_first(first)
        L.flash_attn_debug_set_host_ctas(ctas)
        if rnd == 0:                              # round 0 warms up descriptors, streams and clocks, and checks the output
            ho.fill_(float("nan"))
        call()
        if rnd == 0:
            if ref is None:
                ref = ho.clone()
                assert not torch.isnan(ref).any()
            same = torch.equal(ho, ref)
            d = (ho.float() - ref.float()).abs().max().item()
            print(f"{name:20s}: output {'bit-identical to' if same else f'max-abs {d:.2e} from'} the first variant's", flush=True)
            assert d <= 1e-3
            continue
        t0 = time.perf_counter()
        for _ in range(8):
            call()
        res[name].append((time.perf_counter() - t0) / 8 * 1e3)
for name, v in res.items():
    print(f"{name:20s}: median {statistics.median(v):.3f} ms  min {min(v):.3f}  max {max(v):.3f}   ({len(v)} rounds of 8 calls)")

"""Burst-regime A/B of library builds in ONE process: short runs (30 launches) separated by idle gaps, so the part is not yet
power-capped (tests/harness/ab_quick.py measures the sustained regime).   python tests/harness/burst_ab.py lib1.so lib2.so ..."""
import ctypes
import sys
import time

import torch

libs = []
for path in sys.argv[1:]:
    L = ctypes.CDLL(path)
    L.flash_attn_fwd.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 5 + [ctypes.c_void_p]
    L.flash_attn_fwd.restype = ctypes.c_int
    libs.append((path.split("/")[-1], L))
B, H, N, D = 1, 32, 8192, 128
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = ((torch.rand((B, H, N, D), device="cuda", generator=g) - 0.5).half() for _ in range(3))
o = torch.empty_like(q)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
res = {(n, c): [] for n, _ in libs for c in (0, 1)}
for rnd in range(6):
    for causal in (1, 0):
        for name, L in libs:
            for _ in range(3):
                assert L.flash_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, D, causal, st) == 0
            torch.cuda.synchronize()
            time.sleep(0.15)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30):
                L.flash_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, D, causal, st)
            e1.record()
            torch.cuda.synchronize()
            res[(name, causal)].append(4.0 * B * H * N * N * D / (2 if causal else 1) / (e0.elapsed_time(e1) / 30) / 1e9)
            time.sleep(0.15)
for (name, causal), vals in res.items():
    vals = sorted(vals)
    print(f"{name:28s} {'causal' if causal else 'full  '} N=8192: median {vals[len(vals) // 2]:7.1f}  min {vals[0]:7.1f}  max {vals[-1]:7.1f} TFLOPS")

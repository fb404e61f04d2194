"""Per-tile softmax timing of a -DFA_TIMING build: cycles a softmax warp waits for S vs cycles from
S-ready to P-arrive.   FLASH_ATTN_B200_LIB=build/lib_timing.so python tests/harness/timing.py [N] [causal] [B] [H] [D]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import flash_attention_cuda_b200 as fa  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
causal = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
B, H, D = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((3, 1), (4, 32), (5, 128)))
L = fa.lib()
L.flash_attn_debug_timing.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
buf = (ctypes.c_ulonglong * 64)()
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = ((torch.rand((B, H, N, D), device="cuda", generator=g) - 0.5).half() for _ in range(3))
o = torch.empty_like(q)
for _ in range(3):
    fa.flash_attn_fwd(q, k, v, causal=causal, out=o)
L.flash_attn_debug_timing(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    fa.flash_attn_fwd(q, k, v, causal=causal, out=o)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
L.flash_attn_debug_timing(buf, 0)
fl = 4.0 * B * H * N * N * D / (2 if causal else 1)
print(f"{os.path.basename(fa.LIB_PATH)} B={B} H={H} N={N} D={D} causal={causal}: {fl / ms / 1e9:.1f} TFLOPS")
for t in range(2):
    w, b, n = (buf[t * 3 + i] for i in range(3))
    if n:
        print(f"  tile {t}: wait-for-S {w / n:7.1f} cyc  S->P {b / n:7.1f} cyc  period {(w + b) / n:7.1f}  ({n} sampled tiles)")
for t in range(2):
    if buf[9 + t * 2]:
        n_it = buf[9 + t * 2]
        print(f"  tile {t}: item start -> first S {buf[8 + t * 2] / n_it:8.0f} cyc   epilogue {buf[14 + t] / n_it:7.0f} cyc   whole item {buf[16 + t] / n_it:9.0f} cyc"
              f"   kernel start -> first S of the CTA's first item {buf[12 + t] / max(1, buf[22]):8.0f} cyc   ({n_it} items)")
c, ns, n = buf[20], buf[21], buf[22]
if n:
    print(f"  CTA lifetime: {c / n:.0f} cycles = {ns / n / 1e3:.1f} us -> SM clock {c / ns * 1e3:.0f} MHz; kernel {ms * 1e3:.1f} us")

"""Burst-regime A/B of library builds over several shapes in ONE process (short runs separated by idle gaps).
   python tests/harness/ab_shapes.py [--env K=V ...] lib1.so lib2.so ... -- B,H,N,D,causal ...
   A lib argument may carry environment settings for the work decomposition: path.so@SPLIT=1"""
import ctypes
import os
import sys
import time

import torch

args = sys.argv[1:]
split = args.index("--")
libs_arg, shapes_arg = args[:split], args[split + 1:]
libs = []
for a in libs_arg:
    path, _, tag = a.partition("@")
    L = ctypes.CDLL(os.path.abspath(path))
    L.flash_attn_fwd.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 5 + [ctypes.c_void_p]
    L.flash_attn_fwd.restype = ctypes.c_int
    if hasattr(L, "flash_attn_debug_set_split"):
        L.flash_attn_debug_set_split.argtypes = [ctypes.c_int]
    libs.append((os.path.basename(path) + ("@" + tag if tag else ""), L, tag))
shapes = [tuple(int(x) for x in s.split(",")) for s in shapes_arg]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print(f"{'shape':28s} " + " ".join(f"{n:>26s}" for n, _, _ in libs))
for (B, H, N, D, causal) in shapes:
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = ((torch.rand((B, H, N, D), device="cuda", generator=g) - 0.5).half() for _ in range(3))
    o = torch.empty_like(q)
    fl = 4.0 * B * H * N * N * D / (2 if causal else 1)
    res = {n: [] for n, _, _ in libs}
    cold = {n: [] for n, _, _ in libs}
    for rnd in range(5):
        for name, L, tag in libs:
            if tag.startswith("SPLIT") and hasattr(L, "flash_attn_debug_set_split"):
                L.flash_attn_debug_set_split(int(tag.split("=")[1]))      # @SPLIT=0 pair items, @SPLIT=1 split mode, @SPLIT=-1 automatic
            for _ in range(3):
                assert L.flash_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, D, causal, st) == 0
            torch.cuda.synchronize()
            time.sleep(0.1)
            iters = 30
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                L.flash_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, D, causal, st)
            e1.record()
            torch.cuda.synchronize()
            res[name].append(fl / (e0.elapsed_time(e1) / iters) / 1e9)
            # cold L2: flush between launches, each launch timed on its own
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
            for a, z in ev:
                flush.fill_(1)
                a.record()
                L.flash_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, H, N, D, causal, st)
                z.record()
            torch.cuda.synchronize()
            cold[name].append(fl / (sum(a.elapsed_time(z) for a, z in ev) / len(ev)) / 1e9)
            time.sleep(0.1)
    med = lambda xs: sorted(xs)[len(xs) // 2]
    print(f"B{B} H{H} N{N} D{D} c{causal}".ljust(28) + " " + " ".join(f"{med(res[n]):9.1f} hot {med(cold[n]):8.1f} cold" for n, _, _ in libs), flush=True)

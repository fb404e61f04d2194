"""TFLOPS sweep of the current library build (README-table shapes), with a quick parity check first.
   FLASH_ATTN_B200_LIB=build/libfa_x.so python tests/harness/sweep.py [quick]"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _oracle  # noqa: E402
import flash_attention_cuda_b200 as fa  # noqa: E402

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"


def check(B, H, N, D, causal, dist):
    rng = np.random.default_rng(N)
    if dist == "n":
        q, k = (rng.standard_normal((B, H, N, D), dtype=np.float32).astype(np.float16) for _ in range(2))
        v = (rng.standard_normal((B, H, N, D), dtype=np.float32) * 0.5).astype(np.float16)
    else:
        q, k, v = ((rng.random((B, H, N, D), dtype=np.float32) - 0.5).astype(np.float16) for _ in range(3))
    out = fa.flash_attn_fwd(*(torch.from_numpy(x).cuda() for x in (q, k, v)), causal=bool(causal))
    torch.cuda.synchronize()
    mx, mean = _oracle.diff(out.cpu().numpy(), _oracle.attention(q, k, v, causal))
    ok = mx <= 2e-3 and mean <= 2e-4
    return ok, mx, mean


import pynvml
pynvml.nvmlInit()
_h = pynvml.nvmlDeviceGetHandleByIndex(0)
CLOCKS = []


def tflops(B, H, N, D, causal, iters=None, warm=5, target_ms=150.0):
    """Times enough back-to-back launches to cover ~target_ms (steady clocks), returns TFLOPS and
    records the SM clock NVML reports right after the timed loop."""
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = ((torch.rand((B, H, N, D), device="cuda", generator=g) - 0.5).half() for _ in range(3))
    o = torch.empty_like(q)
    fl = 4.0 * B * H * N * N * D / (2 if causal else 1)
    for _ in range(warm):
        fa.flash_attn_fwd(q, k, v, causal=bool(causal), out=o)
    if iters is None:
        iters = int(max(10, min(2000, target_ms / (fl / 1.0e12 * 1e-3 * 1e3 / 1.0))))   # assume ~1 PFLOP/s
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fa.flash_attn_fwd(q, k, v, causal=bool(causal), out=o)
    e1.record()
    CLOCKS.append(pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM))
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return fl / ms / 1e9


print("lib:", fa.LIB_PATH)
bad = 0
for case in [(1, 2, 384, 128, 1, "n"), (1, 2, 1000, 128, 0, "n"), (1, 4, 2048, 128, 1, "u"), (2, 2, 777, 64, 1, "n")]:
    ok, mx, mean = check(*case)
    bad += not ok
    print(("PASS" if ok else "FAIL"), case, f"max={mx:.2e} mean={mean:.2e}")
wd = fa.watchdog_status()
print("watchdog:", wd)
seqs = (2048, 8192) if quick else (512, 1024, 2048, 4096, 8192, 16384)
for causal in (0, 1):
    print("causal" if causal else "full  ", " ".join(f"N{n}:{tflops(1, 32, n, 128, causal):7.1f}" for n in seqs))
print("SM clock (MHz) sampled during each timed loop:", CLOCKS)
print("d64 B32 H16 N2048 full:", f"{tflops(32, 16, 2048, 64, 0):7.1f}", "| cfg3-like B4 H32 N8192 causal:", f"{tflops(4, 32, 8192, 128, 1):7.1f}")
sys.exit(1 if bad or wd["aborted"] else 0)

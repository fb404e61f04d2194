"""PCIe floor of the end-to-end path: pinned H2D of Q,K,V (192 MiB), D2H of O (64 MiB), alone and concurrently,
then flash_attn_fwd_host with 4 / 8 / 16 / 32 head chunks (FLASH_ATTN_B200_HOST_CHUNKS, one process each)."""
import os
import subprocess
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "host":
    sys.path.insert(0, REPO)
    import flash_attention_cuda_b200 as fa
    B, H, N, D = 1, 32, 8192, 128
    hq, hk, hv = ((torch.rand((B, H, N, D)) - 0.5).half().pin_memory() for _ in range(3))
    ho = torch.empty((B, H, N, D), dtype=torch.float16).pin_memory()
    L = fa.lib()
    for _ in range(3):
        fa.check(L.flash_attn_fwd_host(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), ho.data_ptr(), B, H, N, D, 1))
    t0 = time.perf_counter()
    for _ in range(20):
        fa.check(L.flash_attn_fwd_host(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), ho.data_ptr(), B, H, N, D, 1))
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print(f"chunks={os.environ.get('FLASH_ATTN_B200_HOST_CHUNKS', 'default')}: {ms:.3f} ms per call, {4.0 * B * H * N * N * D / 2 / ms / 1e9:.1f} TFLOPS e2e")
    sys.exit(0)

MiB = 1 << 20
hin = torch.empty(192 * MiB, dtype=torch.uint8).pin_memory()
hout = torch.empty(64 * MiB, dtype=torch.uint8).pin_memory()
din = torch.empty(192 * MiB, dtype=torch.uint8, device="cuda")
dout = torch.empty(64 * MiB, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def h2d():
    with torch.cuda.stream(s1):
        din.copy_(hin, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        hout.copy_(dout, non_blocking=True)


a, b = timed(h2d), timed(d2h)
c = timed(lambda: (h2d(), d2h()))
print(f"H2D 192 MiB alone {a:.3f} ms ({192 * MiB / a / 1e6:.1f} GB/s); D2H 64 MiB alone {b:.3f} ms ({64 * MiB / b / 1e6:.1f} GB/s); "
      f"both at once {c:.3f} ms")
for ch in ("4", "8", "16", "32"):
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "host"], env=dict(os.environ, FLASH_ATTN_B200_HOST_CHUNKS=ch),
                       capture_output=True, text=True)
    print(r.stdout.strip(), r.stderr.strip()[-200:])

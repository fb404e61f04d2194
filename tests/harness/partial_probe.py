"""Cost of the partial-state epilogue (ring hop kernels) relative to the plain forward, one GPU.
   python tests/harness/partial_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import flash_attention_cuda_b200 as fa  # noqa: E402

B, H, C, D = 1, 32, 8192, 128
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = ((torch.rand((B, H, C, D), device="cuda", generator=g) - 0.5).half() for _ in range(3))
o = torch.empty_like(q)
op = torch.empty((B * H * C, D), dtype=torch.float32, device="cuda")
ml = torch.empty((B * H * C, 2), dtype=torch.float32, device="cuda")


def timed(fn, n=30):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


fl = 4.0 * B * H * C * C * D
for causal in (False, True):
    f = fl / (2 if causal else 1)
    t0 = timed(lambda: fa.flash_attn_fwd(q, k, v, causal=causal, out=o))
    t1 = timed(lambda: fa.flash_attn_fwd_partial(q, k, v, op, ml, causal, 0, 0, False))
    t2 = timed(lambda: fa.flash_attn_fwd_partial(q, k, v, op, ml, causal, 0, 0, True))
    print(f"causal={causal}: fp16 out {t0:.3f} ms ({f / t0 / 1e9:.0f} TFLOPS) | partial write {t1:.3f} ms ({f / t1 / 1e9:.0f}) | "
          f"partial accumulate {t2:.3f} ms ({f / t2 / 1e9:.0f})")
t3 = timed(lambda: fa.flash_attn_finalize(op, ml, o))
print(f"finalize: {t3:.3f} ms")

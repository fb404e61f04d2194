"""Runs the forward kernel a few times on one shape (target for ncu / compute-sanitizer).
   python tests/harness/profile_one.py B H N D causal [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import flash_attention_cuda_b200 as fa  # noqa: E402

B, H, N, D, causal = (int(x) for x in sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = ((torch.rand((B, H, N, D), device="cuda", generator=g) - 0.5).half() for _ in range(3))
o = torch.empty_like(q)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    if i == iters - 1:
        e0.record()
    fa.flash_attn_fwd(q, k, v, causal=bool(causal), out=o)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
fl = 4.0 * B * H * N * N * D / (2 if causal else 1)
print(f"B{B} H{H} N{N} D{D} causal={causal}: last launch {ms:.4f} ms = {fl / ms / 1e9:.1f} TFLOPS; checksum {o.float().sum().item():.4f}")

"""The reference README's comparison column ("PyTorch FA2", README.md:13-35) reproduced on this box:
torch's scaled_dot_product_attention (library code: cuDNN / flash backend, whatever torch picks) next to
libflashattn_b200 on the README table shapes (B=1, H=32, D=128).  A reference point only.
   python tests/harness/sdpa_compare.py"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import flash_attention_cuda_b200 as fa  # noqa: E402


def timed(fn, iters, reps=10):
    """GPU time per call: `reps` calls recorded into a CUDA graph and replayed, so that neither side pays (or hides behind)
    its host-side dispatch -- at N=512 a call is ~10 us of GPU time, less than the Python around it."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    n = max(2, iters // reps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps)


print(f"{'seq':>6} {'causal':>6} {'ours TFLOPS':>12} {'torch SDPA':>11} {'max|diff|':>10}")
for causal in (False, True):
    for n in (512, 1024, 2048, 4096, 8192, 16384):
        g = torch.Generator(device="cuda").manual_seed(n)
        q, k, v = ((torch.rand((1, 32, n, 128), device="cuda", generator=g) - 0.5).half() for _ in range(3))
        o = torch.empty_like(q)
        fl = 4.0 * 32 * n * n * 128 / (2 if causal else 1)
        iters = max(20, min(500, int(2e11 / fl * 50)))
        t_ours = timed(lambda: fa.flash_attn_fwd(q, k, v, causal=causal, out=o), iters)
        t_sdpa = timed(lambda: F.scaled_dot_product_attention(q, k, v, is_causal=causal), iters)
        d = (o.float() - F.scaled_dot_product_attention(q, k, v, is_causal=causal).float()).abs().max().item()
        print(f"{n:6d} {int(causal):6d} {fl / t_ours / 1e9:12.1f} {fl / t_sdpa / 1e9:11.1f} {d:10.2e}", flush=True)

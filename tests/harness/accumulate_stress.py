"""Repeats the in-place accumulate sequence of flash_attn_fwd_ex (4 K/V blocks, then finalize) and compares with the monolithic
kernel on the GPU:  python tests/harness/accumulate_stress.py [iters]   (FLASH_ATTN_B200_LIB selects the build)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import flash_attention_cuda_b200 as fa  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
bad = 0
worst = 0.0
for it in range(iters):
    for (B, H, N, D, P) in ((1, 2, 1024, 128, 4), (1, 8, 2048, 128, 4), (2, 4, 1024, 64, 2)):
        g = torch.Generator(device="cuda").manual_seed(31 + it)
        q, k = (torch.randn((B, H, N, D), device="cuda", generator=g).half() for _ in range(2))
        v = (torch.randn((B, H, N, D), device="cuda", generator=g) * 0.5).half()
        full = fa.flash_attn_fwd(q, k, v, causal=True)
        o_part = torch.empty((B * H * N, D), dtype=torch.float32, device="cuda")
        ml = torch.empty((B * H * N, 2), dtype=torch.float32, device="cuda")
        blk = N // P
        for s in range(P):
            ks = k[:, :, s * blk:(s + 1) * blk].contiguous()
            vs = v[:, :, s * blk:(s + 1) * blk].contiguous()
            fa.flash_attn_fwd_partial(q, ks, vs, o_part, ml, True, 0, s * blk, accumulate=(s > 0))
        out = torch.empty_like(q)
        fa.flash_attn_finalize(o_part, ml, out)
        torch.cuda.synchronize()
        d = (out.float() - full.float()).abs().max().item()
        worst = max(worst, d)
        if not d <= 2e-3:
            bad += 1
            rows = ((out.float() - full.float()).abs().amax(dim=-1) > 2e-3).nonzero()
            print(f"iter {it} shape {(B, H, N, D, P)}: max|diff| {d:.3e}, bad rows {rows[:6].tolist()} (+{max(0, len(rows) - 6)})", flush=True)
print(f"{os.path.basename(fa.LIB_PATH)}: {bad} bad of {iters * 3}, worst {worst:.3e}, watchdog {fa.watchdog_status()['aborted']}")

"""`torch.library` registration of the forward kernel: `torch.ops.flashattn_b200.fwd`.

The reference ships no binding and compares itself against "PyTorch FA2" in its README (README.md:13, 25);
this makes the same comparison reproducible on the box -- `F.scaled_dot_product_attention` as one more
reference point next to the CPU oracle, never as a dispatch target -- and lets the kernel sit inside
`torch.compile`d graphs (a fake/meta implementation describes the output).  The op is a thin wrapper
over the C ABI (`flash_attn_fwd` / `flash_attn_fwd_bf16`): FP16 or BF16 CUDA tensors `[B, H, N, D]`, D in {64, 128},
scale 1/sqrt(D).
There is no CPU implementation: calling it with CPU tensors raises.

    import flash_attention_cuda_b200.torch_op            # registers the op
    o = torch.ops.flashattn_b200.fwd(q, k, v, True)      # causal
"""
from __future__ import annotations

import torch

from . import flash_attn_fwd as _fwd

_lib = torch.library.Library("flashattn_b200", "DEF")
_lib.define("fwd(Tensor q, Tensor k, Tensor v, bool causal) -> Tensor")


def _fwd_cuda(q, k, v, causal):
    return _fwd(q.contiguous(), k.contiguous(), v.contiguous(), causal=bool(causal))


def _fwd_meta(q, k, v, causal):
    if q.dim() != 4 or q.shape != k.shape or q.shape != v.shape:
        raise ValueError("q, k, v must all be [B, H, N, D]")
    if q.dtype not in (torch.float16, torch.bfloat16) or k.dtype != q.dtype or v.dtype != q.dtype:
        raise TypeError("flashattn_b200::fwd takes float16 or bfloat16 tensors, all of one type")
    if q.shape[-1] not in (64, 128):
        raise ValueError("head_dim must be 64 or 128")
    return torch.empty_like(q, memory_format=torch.contiguous_format)


def _fwd_cpu(q, k, v, causal):
    raise RuntimeError("flashattn_b200::fwd has no CPU implementation (the product is the sm_100a kernel)")


_lib.impl("fwd", _fwd_cuda, "CUDA")
_lib.impl("fwd", _fwd_meta, "Meta")
_lib.impl("fwd", _fwd_cpu, "CPU")


def flash_attn(q, k, v, causal: bool = True):
    """Functional spelling of `torch.ops.flashattn_b200.fwd`."""
    return torch.ops.flashattn_b200.fwd(q, k, v, causal)

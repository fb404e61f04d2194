"""`torch.ops.flashattn_b200.fwd` (flash_attention_cuda_b200/torch_op.py): registration and the fake/meta
kernel on CPU; on a B200 the op against the CPU oracle and against `F.scaled_dot_product_attention`
(the reference README's "PyTorch FA2" comparison point, README.md:13, 25 -- a reference point, not a path)."""
import numpy as np
import pytest

import _oracle

torch = pytest.importorskip("torch")


def test_op_is_registered_and_meta_kernel_describes_the_output():
    import flash_attention_cuda_b200.torch_op  # noqa: F401
    q = torch.empty((2, 3, 100, 128), dtype=torch.float16, device="meta")
    o = torch.ops.flashattn_b200.fwd(q, q, q, True)
    assert o.shape == q.shape and o.dtype == torch.float16 and o.device.type == "meta"
    qb = q.to(torch.bfloat16)
    assert torch.ops.flashattn_b200.fwd(qb, qb, qb, False).dtype == torch.bfloat16
    with pytest.raises((TypeError, RuntimeError)):
        torch.ops.flashattn_b200.fwd(qb, q, q, True)          # mixed operand types
    with pytest.raises((TypeError, RuntimeError)):
        torch.ops.flashattn_b200.fwd(q.float(), q.float(), q.float(), True)
    with pytest.raises((ValueError, RuntimeError)):
        torch.ops.flashattn_b200.fwd(torch.empty((2, 3, 100, 96), dtype=torch.float16, device="meta"), q, q, True)


def test_no_cpu_implementation():
    import flash_attention_cuda_b200.torch_op  # noqa: F401
    q = torch.zeros((1, 1, 8, 64), dtype=torch.float16)
    with pytest.raises(RuntimeError):
        torch.ops.flashattn_b200.fwd(q, q, q, False)


@pytest.mark.gpu
@pytest.mark.parametrize("D,causal", [(128, True), (64, False)])
def test_op_matches_oracle_and_sdpa(D, causal):
    import flash_attention_cuda_b200.torch_op  # noqa: F401
    rng = np.random.default_rng(5)
    q, k, v = (rng.standard_normal((2, 4, 777, D), dtype=np.float32) for _ in range(3))
    q, k, v = q.astype(np.float16), k.astype(np.float16), (v * 0.5).astype(np.float16)
    tq, tk, tv = (torch.from_numpy(x).cuda() for x in (q, k, v))
    out = torch.ops.flashattn_b200.fwd(tq, tk, tv, causal)
    torch.cuda.synchronize()
    mx, mean = _oracle.diff(out.cpu().numpy(), _oracle.attention(q, k, v, int(causal)))
    assert mx <= _oracle.MAX_ABS_TOL and mean <= _oracle.MEAN_ABS_TOL
    sdpa = torch.nn.functional.scaled_dot_product_attention(tq, tk, tv, is_causal=causal)
    assert (out.float() - sdpa.float()).abs().max().item() <= 4e-3     # two fp16 kernels, each within 2e-3 of fp32

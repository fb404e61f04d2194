"""GPU tests of the boundary itself (pytest -m gpu): the `./flash_attention <seq> <causal>` harness binary
(reference main(), flash_attention.cu:702-974, with the argv contract of README.md:83-85) including the
reference-signature C++ shim it links, the watchdog report, CUDA-graph capture of the launcher, the descriptor
cache and the argument checks of the Python binding."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import _oracle

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(REPO, "flash_attention")


@pytest.fixture(scope="module")
def fa():
    import flash_attention_cuda_b200 as m
    m.lib()
    return m


def run_cli(*args, timeout=600):
    if not os.path.exists(CLI):
        subprocess.run(["make", "-C", REPO, "flash_attention"], check=True, capture_output=True)
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=timeout, cwd=REPO)


@pytest.mark.parametrize("seq,causal", [(1024, 1), (2048, 0)])
def test_cli_single_shape_checks_and_benchmarks(seq, causal):
    """README.md:83-85: `./flash_attention 1024 1`, `./flash_attention 2048 0`.  The correctness block goes through
    flash_attention_b200_dispatch, the shim with the reference's 12-argument signature (include/flash_attn.h)."""
    r = run_cli(str(seq), str(causal), "--quick")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    m = re.search(r"b200: max_diff=([0-9.eE+-]+) mean_diff=([0-9.eE+-]+) (PASS|FAIL)", r.stdout)
    assert m, r.stdout[-2000:]
    assert m.group(3) == "PASS"
    assert float(m.group(1)) <= _oracle.MAX_ABS_TOL and float(m.group(2)) <= _oracle.MEAN_ABS_TOL
    assert "ALL CHECKS PASS" in r.stdout
    assert ("CAUSAL" if causal else "NON-CAUSAL") in r.stdout
    # the benchmark row: seq, then a TFLOPS column for this library
    row = [ln for ln in r.stdout.splitlines() if re.match(rf"\s*{seq}\b", ln)]
    assert row, r.stdout[-2000:]
    assert max(float(x) for x in re.findall(r"\d+\.\d+", row[-1])) > 50.0, row[-1]


def test_cli_rejects_what_the_library_rejects():
    r = run_cli("256", "1", "--dim", "96", "--quick", "--no-v9")
    assert r.returncode != 0
    assert "head_dim" in (r.stdout + r.stderr)
    r = run_cli("256", "1", "extra", "positional")
    assert r.returncode == 2 and "usage" in r.stderr


def _rand(shape, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return tuple((torch.rand(shape, device="cuda", generator=g) - 0.5).half() for _ in range(3))


def test_watchdog_abort_is_reported_once_and_cleared(fa):
    """A kernel that gives up on a barrier leaves a record; the NEXT call returns FA_ERR_WATCHDOG instead of running
    on a poisoned device, clears the record, and the call after that works and is correct again."""
    L = fa.lib()
    q, k, v = _rand((1, 2, 300, 128))
    ref = fa.flash_attn_fwd(q, k, v, causal=True).clone()
    torch.cuda.synchronize()
    assert fa.watchdog_status()["aborted"] == 0
    assert L.flash_attn_debug_trip_watchdog(77, None) == 0        # what a timed-out mbarrier wait does
    torch.cuda.synchronize()
    st = fa.watchdog_status(sync=False)
    assert st["aborted"] == 1 and st["tag"] == 77
    with pytest.raises(fa.FlashAttnError) as ei:
        fa.flash_attn_fwd(q, k, v, causal=True)
    assert ei.value.code == -8
    st = fa.watchdog_status()
    assert st["aborted"] == 0 and st["tag"] == 77                 # reported and cleared; the last report stays readable
    out = fa.flash_attn_fwd(q, k, v, causal=True)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    assert fa.watchdog_status()["aborted"] == 0


def test_launches_recorded_into_cuda_graphs_replay_next_to_eager_ones(fa):
    """A captured launch owns its scheduler word (fa_api.cu: kCaptureSlots): replays of two graphs and eager launches
    on another stream interleave freely and every output equals the eager result."""
    qa, ka, va = _rand((1, 8, 1024, 128), 1)
    qb, kb, vb = _rand((2, 4, 640, 64), 2)
    ref_a = fa.flash_attn_fwd(qa, ka, va, causal=True).clone()
    ref_b = fa.flash_attn_fwd(qb, kb, vb, causal=False).clone()
    oa, ob = torch.empty_like(qa), torch.empty_like(qb)
    torch.cuda.synchronize()
    ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(ga):
        fa.flash_attn_fwd(qa, ka, va, causal=True, out=oa)
        fa.flash_attn_fwd(qa, ka, va, causal=True, out=oa)
    with torch.cuda.graph(gb):
        fa.flash_attn_fwd(qb, kb, vb, causal=False, out=ob)
    side = torch.cuda.Stream()
    eager = torch.empty_like(qa)
    for _ in range(5):
        oa.zero_(); ob.zero_()
        ga.replay()
        with torch.cuda.stream(side):
            for _ in range(3):
                fa.flash_attn_fwd(qa, ka, va, causal=True, out=eager)
        gb.replay()
        torch.cuda.synchronize()
        assert torch.equal(oa, ref_a) and torch.equal(ob, ref_b) and torch.equal(eager, ref_a)
    assert fa.watchdog_status()["aborted"] == 0


def test_descriptor_cache_follows_pointers_and_shapes(fa):
    """Repeated calls reuse cached TMA descriptors; a different shape on the SAME allocation must not."""
    buf = torch.empty(3, 1, 4, 512, 128, dtype=torch.float16, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(5)
    buf.copy_(torch.rand(buf.shape, device="cuda", generator=g) - 0.5)
    q, k, v = buf[0], buf[1], buf[2]
    o1 = fa.flash_attn_fwd(q, k, v, causal=True).clone()
    o2 = fa.flash_attn_fwd(q, k, v, causal=True).clone()          # cache hit
    assert torch.equal(o1, o2)
    # same base pointers, other geometry: [1, 8, 256, 128] over the same bytes
    q2, k2, v2 = (t.reshape(1, 8, 256, 128) for t in (q, k, v))
    o3 = fa.flash_attn_fwd(q2, k2, v2, causal=True)
    torch.cuda.synchronize()
    hq, hk, hv = (t.cpu().numpy() for t in (q2, k2, v2))
    mx, mean = _oracle.diff(o3.cpu().numpy(), _oracle.attention(hq, hk, hv, True))
    assert mx <= _oracle.MAX_ABS_TOL and mean <= _oracle.MEAN_ABS_TOL
    # and many distinct buffers (more than the cache holds) still give right answers
    outs = []
    for i in range(12):
        qi, ki, vi = _rand((1, 2, 130, 128), 100 + i)
        outs.append((qi, ki, vi, fa.flash_attn_fwd(qi, ki, vi, causal=False)))
    torch.cuda.synchronize()
    for qi, ki, vi, oi in outs[::5]:
        mx, mean = _oracle.diff(oi.cpu().numpy(), _oracle.attention(*(t.cpu().numpy() for t in (qi, ki, vi)), False))
        assert mx <= _oracle.MAX_ABS_TOL and mean <= _oracle.MEAN_ABS_TOL


def test_binding_refuses_wrong_formats_on_the_partial_and_merge_path(fa):
    """the partial/merge path is FP16-only and stride-blind; the binding has to say so."""
    from flash_attention_cuda_b200 import ring
    q, k, v = _rand((1, 2, 256, 128))
    o_part = torch.empty(1 * 2 * 256, 128, dtype=torch.float32, device="cuda")
    ml = torch.empty(1 * 2 * 256, 2, dtype=torch.float32, device="cuda")
    fa.flash_attn_fwd_partial(q, k, v, o_part, ml, True, 0, 0, False)                     # the good call
    with pytest.raises(TypeError):
        fa.flash_attn_fwd_partial(q.bfloat16(), k.bfloat16(), v.bfloat16(), o_part, ml, True, 0, 0, False)
    with pytest.raises(ValueError):
        fa.flash_attn_fwd_partial(q[:, :, :128], k, v, o_part[:256], ml[:256], True, 0, 0, False)   # slice along N
    with pytest.raises(TypeError):
        fa.flash_attn_fwd_partial(q, k, v, o_part.half(), ml, True, 0, 0, False)
    with pytest.raises(ValueError):
        fa.flash_attn_fwd_partial(q, k, v, o_part[:100], ml, True, 0, 0, False)
    out = torch.empty(1, 2, 256, 128, dtype=torch.float16, device="cuda")
    fa.flash_attn_merge(o_part[None], ml[None], out)
    with pytest.raises(TypeError):
        fa.flash_attn_merge(o_part[None], ml[None], out.bfloat16())
    with pytest.raises(ValueError):
        fa.flash_attn_merge(o_part[None], ml[None, :100], out)
    with pytest.raises(TypeError):
        fa.flash_attn_finalize(o_part, ml, out.float())
    with pytest.raises(ValueError):
        fa.flash_attn_fwd(q, k, v, causal=True, out=torch.empty(1, 2, 128, 128, dtype=torch.float16, device="cuda"))
    with pytest.raises(ValueError):
        fa.flash_attn_fwd(q, k, v, causal=True, out=out.transpose(1, 2))
    with pytest.raises(TypeError):
        ring._check_chunks([q.bfloat16()] * 2, [k.bfloat16()] * 2, [v.bfloat16()] * 2)
    with pytest.raises(ValueError):
        ring._check_chunks([q[:, :, :128], q[:, :, 128:]], [k, k], [v, v])
    torch.cuda.synchronize()
    mx, mean = _oracle.diff(out.cpu().numpy(), _oracle.attention(*(t.cpu().numpy() for t in (q, k, v)), True))
    assert mx <= _oracle.MAX_ABS_TOL and mean <= _oracle.MEAN_ABS_TOL


def test_host_entry_point_reports_errors_without_leaving_work_behind(fa):
    """flash_attn_fwd_host with an argument the kernel launcher rejects mid-pipeline drains its three streams."""
    L = fa.lib()
    h = torch.zeros(1, 2, 64, 128, dtype=torch.float16).pin_memory()
    assert L.flash_attn_fwd_host(h.data_ptr(), h.data_ptr(), h.data_ptr(), h.data_ptr(), 1, 2, 64, 96, 1) == -1
    assert L.flash_attn_fwd_host(h.data_ptr(), h.data_ptr(), h.data_ptr(), h.data_ptr(), 1, 2, 64, 128, 1) == 0
    torch.cuda.synchronize()
    assert not torch.isnan(h.float()).any()


def test_destroy_then_reuse(fa):
    q, k, v = _rand((1, 3, 200, 64), 9)
    ref = fa.flash_attn_fwd(q, k, v, causal=True).clone()
    torch.cuda.synchronize()
    fa.lib().flash_attn_destroy()
    out = fa.flash_attn_fwd(q, k, v, causal=True)                  # sets the device up again
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    assert fa.kernel_info(1, 32, 1024, 128, True)["ctas"] >= 1

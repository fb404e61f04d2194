# BF16 operand path: parity tests, then FP16 vs BF16 throughput on the same box
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "bf16" > gpurun_out/bf16_pytest.log 2>&1; echo pytest rc=$?
tail -n 15 gpurun_out/bf16_pytest.log
python - > gpurun_out/bf16_perf.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
import flash_attention_cuda_b200 as fa
def tf(dt, B, H, N, D, causal, iters):
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = ((torch.rand((B, H, N, D), device="cuda", generator=g) - 0.5).to(dt) for _ in range(3))
    o = torch.empty_like(q)
    for _ in range(5): fa.flash_attn_fwd(q, k, v, causal=bool(causal), out=o)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fa.flash_attn_fwd(q, k, v, causal=bool(causal), out=o)
    e1.record(); torch.cuda.synchronize()
    return 4.0 * B * H * N * N * D / (2 if causal else 1) / (e0.elapsed_time(e1) / iters) / 1e9
for dt in (torch.float16, torch.bfloat16, torch.float16, torch.bfloat16):
    print(dt, "full8192 %7.1f  causal8192 %7.1f  causal2048 %7.1f  d64 %7.1f" % (tf(dt,1,32,8192,128,0,150), tf(dt,1,32,8192,128,1,300), tf(dt,1,32,2048,128,1,1000), tf(dt,32,16,2048,64,0,200)), fa.watchdog_status()["aborted"], flush=True)
PY
cat gpurun_out/bf16_perf.log

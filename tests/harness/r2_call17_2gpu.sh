# Round 2, call 17 (2 GPUs): the new kernel build under context parallelism (tests + every bench leg), and the gathered
# form with its two kernels on two streams (FLASH_ATTN_GATHER_OVERLAP=1)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "gathered or context_parallel or pull" > gpurun_out/r02_c17_pytest_2gpu.log 2>&1; echo pytest rc=$?
tail -n 5 gpurun_out/r02_c17_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c17_bench_2gpu.json 2> gpurun_out/r02_c17_bench_2gpu.err; echo bench rc=$?
for ov in 0 1 0 1; do
  FLASH_ATTN_GATHER_OVERLAP=$ov timeout 600 $TR --master-port 2954$ov bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --workload cfg5_ring_n131072_causal --ring-exchange gather 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('overlap=$ov', d.get('value'), d.get('ms_per_step'))
" | tee -a gpurun_out/r02_c17_gather_overlap.log
done
python - <<'PY'
import json
for line in open('gpurun_out/r02_c17_bench_2gpu.json'):
    if line.startswith('{'):
        d = json.loads(line)
        print(d.get('value'), d.get('ms_per_step')); print(json.dumps(d.get('cp_cfg5'))); print(json.dumps(d.get('cp_parity'))); print(json.dumps(d.get('strong_cfg3')))
PY
tail -n 5 gpurun_out/r02_c17_bench_2gpu.err

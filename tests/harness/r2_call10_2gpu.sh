# Round 2, tenth GPU call (2 GPUs): gathered-K/V context parallelism
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "gathered or context_parallel" > gpurun_out/r02_c10_pytest.log 2>&1; echo pytest rc=$?
tail -n 25 gpurun_out/r02_c10_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c10_bench_2gpu.json 2> gpurun_out/r02_c10_bench_2gpu.err; echo bench rc=$?
python - <<'PY'
import json
for line in open('gpurun_out/r02_c10_bench_2gpu.json'):
    if line.startswith('{'):
        d = json.loads(line)
        print(json.dumps(d.get('cp_cfg5'), indent=1)); print(json.dumps(d.get('cp_parity'))); print(json.dumps(d.get('strong_cfg3')))
PY
tail -n 15 gpurun_out/r02_c10_bench_2gpu.err

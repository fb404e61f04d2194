# Round 2, eleventh GPU call (8 GPUs): the scaling run with gathered-K/V context parallelism
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_c11_bench_8gpu.json 2> gpurun_out/r02_c11_bench_8gpu.err; echo bench rc=$?
python - <<'PY'
import json
for line in open('gpurun_out/r02_c11_bench_8gpu.json'):
    if line.startswith('{'):
        d = json.loads(line)
        print(d['value'], d['ms_per_step'], d['e2e'])
        print(json.dumps(d.get('cp_cfg5'), indent=1)); print(json.dumps(d.get('cp_parity'))); print(json.dumps(d.get('strong_cfg3')))
PY
tail -n 8 gpurun_out/r02_c11_bench_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tests/harness/ring_check.py 8192 gather,pull 2>&1 | grep "rank\|Error\|error" | sort | tee gpurun_out/r02_c11_ring_check_8gpu.log | tail -n 30

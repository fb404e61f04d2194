# Round 2, call 46 (2 GPUs): the driver's multi-GPU bench line on the final tree (replicated cfg 2 + strong_cfg3 + cp_cfg5 x 3 exchanges + cp_parity)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_c46_bench_2gpu.json 2> gpurun_out/r02_c46_bench_2gpu.err; echo bench rc=$?
python - <<'PY'
import json
for line in open('gpurun_out/r02_c46_bench_2gpu.json'):
    if line.startswith('{'):
        d = json.loads(line)
        print(d.get('value'), d.get('ms_per_step'), json.dumps(d.get('e2e'))[:200]); print(json.dumps(d.get('cp_cfg5'))[:900]); print(json.dumps(d.get('cp_parity'))[:600]); print(json.dumps(d.get('strong_cfg3'))[:400])
PY
tail -n 3 gpurun_out/r02_c46_bench_2gpu.err

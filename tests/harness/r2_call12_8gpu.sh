# Round 2, twelfth GPU call (8 GPUs): gathered context parallelism -- row blocks per chunk, two copy streams
set -x
mkdir -p gpurun_out
for parts in 1 2 4; do
  FLASH_ATTN_GATHER_PARTS=$parts timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2955$parts bench.py --gpus 8 --workload cfg5_ring_n131072_causal --ring-exchange gather --steps 5 2>/dev/null | cut -c1-200
done 2>&1 | tee gpurun_out/r02_c12_gather_parts.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_c12_bench_8gpu.json 2> gpurun_out/r02_c12_bench_8gpu.err; echo bench rc=$?
python - <<'PY'
import json
for line in open('gpurun_out/r02_c12_bench_8gpu.json'):
    if line.startswith('{'):
        d = json.loads(line)
        print(d['value'], d['ms_per_step'])
        print(json.dumps(d.get('cp_cfg5'))); print(json.dumps(d.get('cp_parity'))); print(json.dumps(d.get('strong_cfg3')))
PY
tail -n 5 gpurun_out/r02_c12_bench_8gpu.err

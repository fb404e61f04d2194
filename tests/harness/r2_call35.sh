# Round 2, call 35: flash_attn_fwd_host, number of (tapered) head chunks
set -x
mkdir -p gpurun_out
for rep in 1 2; do for ch in 4 5 6 8 10; do
  FLASH_ATTN_B200_HOST_CHUNKS=$ch timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sustain-s 0 --e2e-steps 30 2>/dev/null > gpurun_out/tmp_e2e.json
  python - <<PY
import json
for l in open('gpurun_out/tmp_e2e.json'):
    if l.startswith('{'):
        e = json.loads(l)['e2e']; print("chunks=$ch rep=$rep", e['ms_per_step'], "ms", e['value'], "TFLOPS, copy floor", e['copy_only_ms'])
PY
done; done 2>&1 | grep chunks= | tee gpurun_out/r02_c35_host_chunks.log

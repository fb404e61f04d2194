# Round 2, call 34: flash_attn_fwd_host with Q, K, V on three input streams and tapered chunks vs round 1's pipeline
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "host" 2>&1 | tail -2
for rep in 1 2; do for cfg in "1 0" "3 0" "1 1" "3 1"; do set -- $cfg
  FLASH_ATTN_B200_HOST_STREAMS=$1 FLASH_ATTN_B200_HOST_TAPER=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sustain-s 0 --e2e-steps 30 2>/dev/null > gpurun_out/tmp_e2e.json
  python tests/harness/print_value.py "streams=$1 taper=$2 rep=$rep" gpurun_out/tmp_e2e.json e2e | grep -v "^streams" | cut -c1-230 | sed "s/^/streams=$1 taper=$2 rep=$rep /"
done; done 2>&1 | tee gpurun_out/r02_c34_host_pipeline.log

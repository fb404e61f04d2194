# Round 2, call 42: zero-copy O store: chunk counts and a small first chunk (one process each), after the parity of both modes
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "host" > gpurun_out/r02_c42_pytest_host.log 2>&1; echo pytest rc=$?
tail -n 3 gpurun_out/r02_c42_pytest_host.log
export FLASH_ATTN_B200_HOST_ZEROCOPY=1
for cfg in "8 0" "8 3" "12 0" "12 4" "6 0" "8 0"; do
  set -- $cfg
  FLASH_ATTN_B200_HOST_CHUNKS=$1 FLASH_ATTN_B200_HOST_FIRST=$2 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sustain-s 0 --e2e-steps 30 > gpurun_out/tmp_e2e.json 2>/dev/null
  python tests/harness/print_value.py "zerocopy chunks=$1 first=$2" gpurun_out/tmp_e2e.json e2e | grep -o "zerocopy.*\|\"ms_per_step\": [0-9.]*\|\"copy_only_ms\": [0-9.]*" | tr '\n' ' ' | tee -a gpurun_out/r02_c42_zerocopy_chunks.log
  echo | tee -a gpurun_out/r02_c42_zerocopy_chunks.log
done

# Round 2, call 40: where the time of flash_attn_fwd_host goes (FLASH_ATTN_B200_HOST_TRACE=1: per-chunk device timestamps)
set -x
mkdir -p gpurun_out
FLASH_ATTN_B200_HOST_TRACE=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sustain-s 0 --e2e-steps 6 > gpurun_out/r02_c40_trace.json 2> gpurun_out/r02_c40_host_trace.log; echo rc=$?
grep -c host_trace gpurun_out/r02_c40_host_trace.log
tail -n 27 gpurun_out/r02_c40_host_trace.log
python tests/harness/print_value.py trace gpurun_out/r02_c40_trace.json e2e | cut -c1-300
FLASH_ATTN_B200_HOST_TRACE=1 FLASH_ATTN_B200_HOST_CHUNKS=4 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sustain-s 0 --e2e-steps 6 > gpurun_out/r02_c40_trace4.json 2> gpurun_out/r02_c40_host_trace4.log; echo rc=$?
tail -n 10 gpurun_out/r02_c40_host_trace4.log
python tests/harness/print_value.py trace4 gpurun_out/r02_c40_trace4.json e2e | cut -c1-300

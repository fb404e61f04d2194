# Round 2, call 41: flash_attn_fwd_host with O stored by the kernel straight into the pinned host buffer (FLASH_ATTN_B200_HOST_ZEROCOPY=1)
# against the staged copy back: parity of both, then e2e A/B (one process each, alternating), then the trace of the new mode
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "host" > gpurun_out/r02_c41_pytest_host.log 2>&1; echo pytest rc=$?
tail -n 5 gpurun_out/r02_c41_pytest_host.log
for rep in 1 2; do for zc in 0 1; do
  FLASH_ATTN_B200_HOST_ZEROCOPY=$zc timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sustain-s 0 --e2e-steps 30 > gpurun_out/tmp_e2e.json 2>/dev/null
  python tests/harness/print_value.py "zerocopy=$zc rep=$rep" gpurun_out/tmp_e2e.json e2e | cut -c1-250 | tee -a gpurun_out/r02_c41_zerocopy_ab.log
done; done
FLASH_ATTN_B200_HOST_ZEROCOPY=1 FLASH_ATTN_B200_HOST_TRACE=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sustain-s 0 --e2e-steps 4 > gpurun_out/tmp_e2e.json 2> gpurun_out/r02_c41_host_trace_zerocopy.log
tail -n 9 gpurun_out/r02_c41_host_trace_zerocopy.log

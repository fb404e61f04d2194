# Round 2, fifth GPU call: where the time of a short-sequence launch goes; half-row softmax microbenchmark
set -x
mkdir -p gpurun_out
for sp in 0 1; do for n in 512 1024 2048; do
  FLASH_ATTN_B200_SPLIT=$sp FLASH_ATTN_B200_LIB=$PWD/build/lib_timing.so timeout 120 python tests/harness/timing.py $n 1
done; done 2>&1 | tee gpurun_out/r02_c5_short_timing.log
FLASH_ATTN_B200_SPLIT=0 FLASH_ATTN_B200_LIB=$PWD/build/lib_timing.so timeout 120 python tests/harness/timing.py 1024 0 2>&1 | tee -a gpurun_out/r02_c5_short_timing.log
FLASH_ATTN_B200_SPLIT=0 FLASH_ATTN_B200_LIB=$PWD/build/lib_timing.so timeout 120 python tests/harness/timing.py 8192 1 2>&1 | tee -a gpurun_out/r02_c5_short_timing.log
timeout 60 ./build/softmax_half_bench 935 2>&1 | tee gpurun_out/r02_c5_softmax_half_bench.log
timeout 60 ./build/softmax_bench_base 935 "== base (full rows, 2 warps per sub-partition)" 2>&1 | tee -a gpurun_out/r02_c5_softmax_half_bench.log
timeout 300 python tests/harness/sdpa_compare.py 2>&1 | tee gpurun_out/r02_c5_sdpa_compare.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4

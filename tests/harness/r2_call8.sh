# Round 2, eighth GPU call: in-kernel probes of the cooperative softmax
set -x
mkdir -p gpurun_out
for c in 0 1; do for causal in 0 1; do
  FLASH_ATTN_B200_COOP=$c FLASH_ATTN_B200_LIB=$PWD/build/lib_timing.so timeout 120 python tests/harness/timing.py 8192 $causal
done; done 2>&1 | tee gpurun_out/r02_c8_coop_timing.log

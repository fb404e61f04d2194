# Round 2, call 23: in-kernel timing probes (-DFA_TIMING) of the final layout: D=128 N=8192 full / causal, D=64 cfg 4, short causal
set -x
mkdir -p gpurun_out
export FLASH_ATTN_B200_LIB=$PWD/build/lib_timing.so
( timeout 120 python tests/harness/timing.py 8192 0
  timeout 120 python tests/harness/timing.py 8192 1
  timeout 120 python tests/harness/timing.py 2048 0 32 16 64
  timeout 120 python tests/harness/timing.py 2048 1
  timeout 120 python tests/harness/timing.py 1024 0 ) 2>&1 | tee gpurun_out/r02_c23_timing.log

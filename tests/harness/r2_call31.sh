# Round 2, call 31: timing probes (every tile) of the short causal shapes on the final kernel, pair items and split mode
set -x
mkdir -p gpurun_out
export FLASH_ATTN_B200_LIB=$PWD/build/lib_timing_all.so
( for n in 512 1024; do for sp in 1 0; do echo "== N=$n split=$sp"; FLASH_ATTN_B200_SPLIT=$sp timeout 120 python tests/harness/timing.py $n 1; done; done
  echo "== N=1024 full (pair)"; timeout 120 python tests/harness/timing.py 1024 0 ) 2>&1 | tee gpurun_out/r02_c31_timing_short.log

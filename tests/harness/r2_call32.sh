# Round 2, call 32: watchdog flag word at the end of the dynamic window (flag), split-mode hand-over in two halves (merge2)
set -x
mkdir -p gpurun_out
FLASH_ATTN_B200_LIB=$PWD/build/lib_merge2.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c32_pytest_merge2.log 2>&1; echo pytest rc=$?
tail -n 3 gpurun_out/r02_c32_pytest_merge2.log
timeout 600 python tests/harness/ab_shapes.py build/lib_before_merge2.so build/lib_flag.so build/lib_merge2.so -- 1,32,512,128,1 1,32,768,128,1 1,32,1024,128,1 1,32,512,128,0 4,16,512,64,1 1,32,8192,128,1 2>&1 | tee gpurun_out/r02_c32_ab_merge2.log

# Round 2, call 29: chunk-wise masking (warp-uniform classification of the four 32-column chunks, exponentials of entirely
# masked chunks skipped) vs the element-wise mask
set -x
mkdir -p gpurun_out
FLASH_ATTN_B200_LIB=$PWD/build/lib_cmask.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c29_pytest_cmask.log 2>&1; echo pytest rc=$?
tail -n 3 gpurun_out/r02_c29_pytest_cmask.log
timeout 600 python tests/harness/ab_shapes.py build/lib_tailfinal.so build/lib_cmask.so -- 1,32,512,128,1 1,32,1024,128,1 1,32,2048,128,1 1,32,4096,128,1 1,32,8192,128,1 1,32,1000,128,0 4,16,1024,64,1 2>&1 | tee gpurun_out/r02_c29_ab_cmask.log

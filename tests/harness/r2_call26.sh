# Round 2, call 26: CTA exit without waiting for the bulk stores' writes (tail2) on top of call 25's build (tail)
set -x
mkdir -p gpurun_out
FLASH_ATTN_B200_LIB=$PWD/build/lib_tail2.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c26_pytest_tail2.log 2>&1; echo pytest rc=$?
tail -n 3 gpurun_out/r02_c26_pytest_tail2.log
timeout 600 python tests/harness/ab_shapes.py build/lib_pre_tail.so build/lib_tail.so build/lib_tail2.so -- 1,32,512,128,1 1,32,1024,128,1 1,32,512,128,0 1,32,1024,128,0 1,32,2048,128,1 1,32,8192,128,1 2>&1 | tee gpurun_out/r02_c26_ab_tail2.log

# Round 2, call 20: split-mode merge straight from tensor memory (no shared-memory round trip of slot 1's O) vs the shipped build
set -x
mkdir -p gpurun_out
FLASH_ATTN_B200_LIB=$PWD/build/lib_tmerge.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_boundary_gpu.py -m gpu -x -q > gpurun_out/r02_c20_pytest_tmerge.log 2>&1; echo pytest rc=$?
tail -n 4 gpurun_out/r02_c20_pytest_tmerge.log
timeout 600 python tests/harness/ab_shapes.py build/lib_new.so build/lib_tmerge.so build/lib_new.so@SPLIT=1 build/lib_tmerge.so@SPLIT=1 -- 1,32,512,128,1 1,32,768,128,1 1,32,1024,128,1 1,32,512,128,0 1,32,1024,128,0 1,32,2048,128,1 4,16,1024,64,1 1,32,8192,128,1 2>&1 | tee gpurun_out/r02_c20_ab_tmerge.log

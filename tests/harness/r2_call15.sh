# Round 2, call 15: code-layout variants (wait inside the variant branch, one-CAS watchdog slow path, rescale block rolled) vs base
set -x
mkdir -p gpurun_out
FLASH_ATTN_B200_LIB=$PWD/build/lib_slim.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_boundary_gpu.py -m gpu -x -q > gpurun_out/r02_c15_pytest_slim.log 2>&1; echo pytest slim rc=$?
tail -n 4 gpurun_out/r02_c15_pytest_slim.log
timeout 600 python tests/harness/burst_ab.py build/lib_base.so build/lib_base_wiv.so build/lib_slim.so 2>&1 | tee gpurun_out/r02_c15_burst_ab.log
timeout 600 python tests/harness/ab_shapes.py build/lib_base.so build/lib_base_wiv.so build/lib_slim.so -- 1,32,1024,128,0 1,32,2048,128,0 1,32,2048,128,1 1,32,4096,128,1 32,16,2048,64,0 1,32,1024,128,1 1,32,512,128,1 2>&1 | tee gpurun_out/r02_c15_ab_shapes.log

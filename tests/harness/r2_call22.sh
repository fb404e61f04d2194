# Round 2, call 22: S address and p_full barrier address pinned in registers (FA_PIN_ADDR) vs the build with the pinned scale
set -x
mkdir -p gpurun_out
FLASH_ATTN_B200_LIB=$PWD/build/lib_sreg2.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r02_c22_pytest_sreg2.log 2>&1; echo pytest rc=$?
tail -n 3 gpurun_out/r02_c22_pytest_sreg2.log
timeout 600 python tests/harness/burst_ab.py build/lib_cur.so build/lib_sreg2.so build/lib_pin.so 2>&1 | tee gpurun_out/r02_c22_burst_ab.log
timeout 600 python tests/harness/ab_shapes.py build/lib_cur.so build/lib_sreg2.so build/lib_pin.so -- 1,32,1024,128,1 1,32,2048,128,1 1,32,4096,128,1 1,32,2048,128,0 32,16,2048,64,0 2>&1 | tee gpurun_out/r02_c22_ab_shapes.log

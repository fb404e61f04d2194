# Round 2, call 25: scheduler re-arm moved from the CTA exit path to the producer's last claim, watchdog publish gated on a
# block-local flag (no global round trips on the tail of a CTA) vs the build before
set -x
mkdir -p gpurun_out
FLASH_ATTN_B200_LIB=$PWD/build/lib_tail.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c25_pytest_tail.log 2>&1; echo pytest rc=$?
tail -n 3 gpurun_out/r02_c25_pytest_tail.log
timeout 600 python tests/harness/ab_shapes.py build/lib_pre_tail.so build/lib_tail.so -- 1,32,512,128,1 1,32,1024,128,1 1,32,512,128,0 1,32,1024,128,0 1,32,2048,128,1 1,32,2048,128,0 32,16,2048,64,0 1,32,8192,128,1 2>&1 | tee gpurun_out/r02_c25_ab_tail.log

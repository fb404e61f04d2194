# Round 2, call 14: epilogue warpgroup (FA_EPI_WG) and wait-in-variant (FA_WAIT_IN_VARIANT) against the shipped kernel.
set -x
mkdir -p gpurun_out
for v in epi_a epi_b; do
  FLASH_ATTN_B200_LIB=$PWD/build/lib_$v.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r02_c14_pytest_$v.log 2>&1; echo pytest $v rc=$?
  tail -n 4 gpurun_out/r02_c14_pytest_$v.log
done
timeout 600 python tests/harness/burst_ab.py build/lib_base.so build/lib_base_wiv.so build/lib_epi_a.so build/lib_epi_b.so build/lib_epi_a_wiv.so 2>&1 | tee gpurun_out/r02_c14_burst_ab.log
timeout 600 python tests/harness/ab_shapes.py build/lib_base.so build/lib_base_wiv.so build/lib_epi_a.so build/lib_epi_b.so build/lib_epi_a_wiv.so -- 1,32,1024,128,0 1,32,2048,128,0 1,32,2048,128,1 1,32,4096,128,1 1,32,4096,128,0 32,16,2048,64,0 1,32,1024,128,1 2>&1 | tee gpurun_out/r02_c14_ab_shapes.log

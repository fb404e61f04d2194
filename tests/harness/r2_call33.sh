# Round 2, call 33: epilogue warpgroup on the final kernel, its warps polling epi_full vs sleeping 500 ns between polls
set -x
mkdir -p gpurun_out
timeout 600 python tests/harness/burst_ab.py build/lib_head.so build/lib_epi2.so build/lib_epi2_sleep.so 2>&1 | tee gpurun_out/r02_c33_burst_ab_epi.log
timeout 600 python tests/harness/ab_shapes.py build/lib_head.so build/lib_epi2.so build/lib_epi2_sleep.so -- 1,32,2048,128,1 1,32,2048,128,0 32,16,2048,64,0 2>&1 | tee -a gpurun_out/r02_c33_burst_ab_epi.log

# Round 2, call 16: FMA-pipe exp2 share and threshold re-measured on the new code layout; streamed softmax once more
set -x
mkdir -p gpurun_out
timeout 600 python tests/harness/burst_ab.py build/lib_new.so build/lib_long3.so build/lib_stream.so 2>&1 | tee gpurun_out/r02_c16_burst_ab.log
timeout 600 python tests/harness/ab_shapes.py build/lib_new.so build/lib_polymin1k.so build/lib_stream.so -- 1,32,1024,128,1 1,32,2048,128,1 1,32,3072,128,1 1,32,4096,128,1 2>&1 | tee gpurun_out/r02_c16_ab_polymin.log
timeout 600 python tests/harness/ab_shapes.py build/lib_new.so build/lib_d64p0.so build/lib_d64p2.so build/lib_d64p3.so build/lib_stream.so -- 32,16,2048,64,0 4,32,4096,64,1 2>&1 | tee gpurun_out/r02_c16_ab_d64.log

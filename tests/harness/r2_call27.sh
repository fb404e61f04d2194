# Round 2, call 27: first work index claimed dynamically (before griddepcontrol.wait) vs static blockIdx
set -x
mkdir -p gpurun_out
timeout 600 python tests/harness/ab_shapes.py build/lib_tailfinal.so build/lib_dyn.so -- 1,32,512,128,1 1,32,1024,128,1 1,32,1024,128,0 1,32,2048,128,1 1,32,2048,128,0 1,32,8192,128,1 2>&1 | tee gpurun_out/r02_c27_ab_dynfirst.log

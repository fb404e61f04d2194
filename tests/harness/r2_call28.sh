# Round 2, call 28: pair items vs split mode re-measured on the final kernel (three copies of the library: the override is per handle)
set -x
mkdir -p gpurun_out
timeout 600 python tests/harness/ab_shapes.py build/lib_auto.so build/lib_pairmode.so@SPLIT=0 build/lib_splitmode.so@SPLIT=1 -- 1,32,512,128,1 1,32,768,128,1 1,32,1024,128,1 1,32,1536,128,1 1,32,2048,128,1 1,32,512,128,0 1,32,768,128,0 1,32,1024,128,0 2,32,1024,128,1 4,16,1024,64,1 2>&1 | tee gpurun_out/r02_c28_ab_split.log

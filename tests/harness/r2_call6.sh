# Round 2, sixth GPU call: FMA-pipe exp2 at every length; graph-timed comparison with torch SDPA; cold/hot sweep
set -x
mkdir -p gpurun_out
timeout 900 python tests/harness/ab_shapes.py build/lib_default.so build/lib_polyall.so -- \
   1,32,512,128,1 1,32,1024,128,1 1,32,2048,128,1 1,32,512,128,0 1,32,1024,128,0 1,32,2048,128,0 16,32,1024,128,1 2>&1 | tee gpurun_out/r02_c6_polyall_ab.log
timeout 300 python tests/harness/sdpa_compare.py 2>&1 | tee gpurun_out/r02_c6_sdpa_compare_default.log
FLASH_ATTN_B200_LIB=$PWD/build/lib_polyall.so timeout 300 python tests/harness/sdpa_compare.py 2>&1 | tee gpurun_out/r02_c6_sdpa_compare_polyall.log

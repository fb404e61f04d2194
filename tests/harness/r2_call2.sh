# Round 2, second GPU call: new boundary tests, the new bench record, single-tile items, FMA-pipe exp2 shares
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c2_pytest.log 2>&1; echo pytest rc=$?
tail -n 15 gpurun_out/r02_c2_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c2_bench_default.json 2> gpurun_out/r02_c2_bench_default.err; echo bench rc=$?
cut -c1-1500 gpurun_out/r02_c2_bench_default.json; tail -n 5 gpurun_out/r02_c2_bench_default.err
timeout 300 python bench.py --steps 50 --warmup 10 --workload cfg1_n1024_causal --no-cpu-baseline > gpurun_out/r02_c2_bench_cfg1.json 2>&1; echo bench rc=$?
cut -c1-900 gpurun_out/r02_c2_bench_cfg1.json
# single-tile work items vs pairs, short shapes (hot and cold L2)
timeout 600 python tests/harness/ab_shapes.py build/lib_single.so@ITEM_TILES=2 build/lib_single.so@ITEM_TILES=1 build/lib_default.so -- \
   1,32,512,128,1 1,32,1024,128,1 1,32,2048,128,1 1,32,512,128,0 1,32,1024,128,0 1,32,2048,128,0 1,32,4096,128,1 32,16,2048,64,0 2>&1 | tee gpurun_out/r02_c2_single_tile_ab.log
# FMA-pipe exp2 share: D=128 long (1 of 4 vs 3 of 8), D=64 (0 vs 1 of 4 vs 3 of 8)
timeout 600 python tests/harness/ab_shapes.py build/lib_default.so build/lib_poly3.so build/lib_d64poly1.so build/lib_d64poly3.so -- \
   1,32,8192,128,1 1,32,8192,128,0 32,16,2048,64,0 8,16,8192,64,1 2>&1 | tee gpurun_out/r02_c2_poly_ab.log

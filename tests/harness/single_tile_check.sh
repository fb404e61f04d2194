# Round-2 starting point (not yet run on a GPU): parity and timing of single-tile work items for grids that cannot fill the machine
set -x
mkdir -p gpurun_out
# build/lib_single.so: nvcc ... -DFA_SINGLE_TILE_MODE -shared flash_attention_cuda_b200/csrc/fa_api.cu (built before the gpurun call)
export FLASH_ATTN_B200_LIB=$PWD/build/lib_single.so
FLASH_ATTN_B200_ITEM_TILES=1 timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_parity_gpu.py::test_experimental_pair_kernel_passes_the_same_parity_tests > gpurun_out/single_pytest.log 2>&1; echo pytest rc=$?
tail -n 3 gpurun_out/single_pytest.log
for m in 2 1; do
  for wl in cfg1_n1024_causal; do
    FLASH_ATTN_B200_ITEM_TILES=$m timeout 100 python bench.py --steps 300 --warmup 30 --no-cpu-baseline --e2e-steps 2 --workload $wl | cut -c1-120
  done
done

# Round 2, seventh GPU call: cooperative softmax -- parity, then A/B against one row per thread
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "softmax_forms or cooperative" > gpurun_out/r02_c7_pytest_coop.log 2>&1; echo pytest coop rc=$?
tail -n 12 gpurun_out/r02_c7_pytest_coop.log
FLASH_ATTN_B200_COOP=1 FLASH_ATTN_B200_SPLIT=0 timeout 900 python -m pytest tests -m gpu -x -q -k "not cli and not watchdog" > gpurun_out/r02_c7_pytest_all_coop.log 2>&1; echo pytest all coop=1 rc=$?
tail -n 12 gpurun_out/r02_c7_pytest_all_coop.log
L=flash_attention_cuda_b200/libflashattn_b200.so
timeout 900 python tests/harness/ab_shapes.py $L@COOP=0 $L@COOP=1 -- \
   1,32,8192,128,1 1,32,8192,128,0 1,32,16384,128,1 1,32,4096,128,1 1,32,2048,128,1 1,32,2048,128,0 32,16,2048,64,0 8,16,8192,64,1 2>&1 | tee gpurun_out/r02_c7_coop_ab.log

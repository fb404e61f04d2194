# Round 2, fourth GPU call: split mode (in-CTA split-KV for short sequences)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c4_pytest.log 2>&1; echo pytest rc=$?
tail -n 15 gpurun_out/r02_c4_pytest.log
FLASH_ATTN_B200_SPLIT=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "not full_size and not config3 and not config5 and not causal_long" > gpurun_out/r02_c4_pytest_split1.log 2>&1; echo pytest split=1 rc=$?
tail -n 8 gpurun_out/r02_c4_pytest_split1.log
FLASH_ATTN_B200_SPLIT=0 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r02_c4_pytest_split0.log 2>&1; echo pytest split=0 rc=$?
tail -n 4 gpurun_out/r02_c4_pytest_split0.log
timeout 900 python tests/harness/ab_shapes.py flash_attention_cuda_b200/libflashattn_b200.so@SPLIT=0 flash_attention_cuda_b200/libflashattn_b200.so@SPLIT=1 -- \
   1,32,512,128,1 1,32,768,128,1 1,32,1024,128,1 1,32,2048,128,1 1,32,4096,128,1 1,32,512,128,0 1,32,768,128,0 1,32,1024,128,0 1,32,2048,128,0 1,32,4096,128,0 \
   4,16,1024,64,0 32,16,2048,64,0 1,8,8192,128,1 1,16,4096,128,1 2>&1 | tee gpurun_out/r02_c4_split_ab.log
timeout 300 python tests/harness/sdpa_compare.py 2>&1 | tee gpurun_out/r02_c4_sdpa_compare.log

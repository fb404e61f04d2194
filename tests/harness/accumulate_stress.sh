# In-place accumulate sequence, repeated: the build before the q_full wait for key-less tiles (build/lib_v4c.so) and the current one
set -x
mkdir -p gpurun_out
for lib in build/lib_v4c.so flash_attention_cuda_b200/libflashattn_b200.so; do
  FLASH_ATTN_B200_LIB=$PWD/$lib timeout 150 python tests/harness/accumulate_stress.py 60 2>&1 | tail -n 4
done | tee gpurun_out/accumulate_stress2.log
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/fix_pytest.log 2>&1; echo pytest rc=$?
tail -n 3 gpurun_out/fix_pytest.log

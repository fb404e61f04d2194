# A/B in one box: (1) register budget / deferred-arrive builds, (2) L2 head-group size: time + DRAM bytes
set -x
mkdir -p gpurun_out
python tests/harness/ab_quick.py flash_attention_cuda_b200/libflashattn_b200.so build/lib_216_DFA_NONE.so build/lib_208_DFA_DEFER_ARRIVE.so build/lib_216_DFA_DEFER_ARRIVE.so > gpurun_out/ab_regs_defer.log 2>&1
cat gpurun_out/ab_regs_defer.log
for mb in 64 32 16; do
  echo "== group $mb MB" >> gpurun_out/ab_group.log
  FLASH_ATTN_B200_L2_GROUP_MB=$mb python tests/harness/ab_quick.py flash_attention_cuda_b200/libflashattn_b200.so 2>&1 | head -1 >> gpurun_out/ab_group.log
  FLASH_ATTN_B200_L2_GROUP_MB=$mb timeout 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fa_fwd_kernel -s 3 -c 1 python tests/harness/profile_one.py 1 32 8192 128 1 5 2>&1 | grep -E "dram__|TFLOPS" >> gpurun_out/ab_group.log
done
cat gpurun_out/ab_group.log

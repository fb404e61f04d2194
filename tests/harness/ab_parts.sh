# A/B: P delivered in 2 vs 3 pieces, setmaxnreg 216/72 (variant libs built into build/ with -DFA_P_PARTS / -DFA_REGS_*), then the GPU parity suite on both
set -x
mkdir -p gpurun_out
python tests/harness/ab_quick.py flash_attention_cuda_b200/libflashattn_b200.so build/lib_parts3.so build/lib_r216_72_parts2.so build/lib_r216_72_parts3.so > gpurun_out/ab_parts.log 2>&1
grep ^round gpurun_out/ab_parts.log | cut -c1-150
timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_parity_gpu.py::test_experimental_pair_kernel_passes_the_same_parity_tests > gpurun_out/ab_parts_pytest_default.log 2>&1; echo pytest default rc=$?
tail -n 3 gpurun_out/ab_parts_pytest_default.log
FLASH_ATTN_B200_LIB=$PWD/build/lib_parts3.so timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_parity_gpu.py::test_experimental_pair_kernel_passes_the_same_parity_tests > gpurun_out/ab_parts_pytest_parts3.log 2>&1; echo pytest parts3 rc=$?
tail -n 3 gpurun_out/ab_parts_pytest_parts3.log

# Parity suite + same-box A/B of the max-free softmax (-DFA_SUM_GUARD build in build/lib_sumguard.so) against the default build
set -x
mkdir -p gpurun_out
FLASH_ATTN_B200_LIB=$PWD/build/lib_sumguard.so timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_parity_gpu.py::test_experimental_pair_kernel_passes_the_same_parity_tests > gpurun_out/sumguard_pytest.log 2>&1; echo pytest sumguard rc=$?
tail -n 25 gpurun_out/sumguard_pytest.log | cut -c1-300
python tests/harness/ab_quick.py build/lib_v4c.so build/lib_sumguard.so > gpurun_out/ab_sumguard.log 2>&1
grep ^round gpurun_out/ab_sumguard.log | cut -c1-150

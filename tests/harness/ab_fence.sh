# A/B: scheduler fence between the two pieces of P (default build) against the previous default (build/lib_v4c.so), then parity
set -x
mkdir -p gpurun_out
python tests/harness/ab_quick.py build/lib_v4c.so flash_attention_cuda_b200/libflashattn_b200.so > gpurun_out/ab_fence.log 2>&1
grep ^round gpurun_out/ab_fence.log | cut -c1-150
timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_parity_gpu.py::test_experimental_pair_kernel_passes_the_same_parity_tests > gpurun_out/fence_pytest.log 2>&1; echo pytest rc=$?
tail -n 3 gpurun_out/fence_pytest.log

# Round 2, call 21: speculative reference (FA_SPEC: piece 0's exponentials start against the row's current reference, the vote
# follows) and the scale pinned in a register (FA_SCALE_REG) vs the shipped build
set -x
mkdir -p gpurun_out
for v in spec spec_sreg; do
FLASH_ATTN_B200_LIB=$PWD/build/lib_$v.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r02_c21_pytest_$v.log 2>&1; echo pytest $v rc=$?
tail -n 3 gpurun_out/r02_c21_pytest_$v.log
done
timeout 600 python tests/harness/burst_ab.py build/lib_cur.so build/lib_sreg.so build/lib_spec.so build/lib_spec_sreg.so 2>&1 | tee gpurun_out/r02_c21_burst_ab.log
timeout 600 python tests/harness/ab_shapes.py build/lib_cur.so build/lib_sreg.so build/lib_spec.so build/lib_spec_sreg.so -- 1,32,1024,128,1 1,32,2048,128,1 1,32,4096,128,1 1,32,2048,128,0 32,16,2048,64,0 2>&1 | tee gpurun_out/r02_c21_ab_shapes.log

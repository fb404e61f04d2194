# Round 2, first GPU call: streamed softmax (new default) against the classic form (-DFA_NO_STREAM)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
for n in base stream; do timeout 60 ./build/softmax_bench_$n 935 "== $n"; done > gpurun_out/r02_softmax_bench.log 2>&1
cat gpurun_out/r02_softmax_bench.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_stream_pytest.log 2>&1; echo pytest rc=$?
tail -n 5 gpurun_out/r02_stream_pytest.log
timeout 300 python tests/harness/burst_ab.py build/lib_classic.so build/lib_default.so 2>&1 | tee gpurun_out/r02_stream_burst_ab.log
for l in timing_classic timing; do
  for c in 0 1; do FLASH_ATTN_B200_LIB=$PWD/build/lib_$l.so timeout 120 python tests/harness/timing.py 8192 $c; done
done 2>&1 | tee gpurun_out/r02_stream_timing.log
timeout 600 python tests/harness/ab_quick.py build/lib_classic.so build/lib_default.so 2>&1 | tee gpurun_out/r02_stream_ab_quick.log
timeout 300 python tests/harness/sdpa_compare.py 2>&1 | tee gpurun_out/r02_stream_sdpa_compare.log

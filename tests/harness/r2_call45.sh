# Round 2, call 45: default bench line with the e2e leg ahead of the sustained leg (two processes)
set -x
mkdir -p gpurun_out
for rep in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c45_bench_ours_$rep.json 2> gpurun_out/r02_c45_bench_ours_$rep.err; echo bench rc=$?
  python tests/harness/print_value.py ours$rep gpurun_out/r02_c45_bench_ours_$rep.json e2e sustained | cut -c1-330
done

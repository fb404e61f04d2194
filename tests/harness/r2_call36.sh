# Round 2, call 36: final tree once more (host pipeline changed): GPU suite, smoke, the default bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo pytest rc=$?
tail -n 3 gpurun_out/r02_final_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2 | tee gpurun_out/r02_final_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_ours.json 2> gpurun_out/r02_final_bench_ours.err; echo bench rc=$?
cut -c1-700 gpurun_out/r02_final_bench_ours.json

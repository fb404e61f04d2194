# Round 2, call 38: the tree with the ours-vs-V9 GPU test: whole GPU suite, the V9 comparison's printed record, smoke, both bench arms
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c38_pytest_gpu.log 2>&1; echo pytest rc=$?
tail -n 4 gpurun_out/r02_c38_pytest_gpu.log
timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -k v9 -s -q > gpurun_out/r02_c38_ours_vs_v9.log 2>&1; echo v9 rc=$?
grep "ours-oracle" gpurun_out/r02_c38_ours_vs_v9.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -n 1 | tee gpurun_out/r02_c38_smoke.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_c38_bench_reference_v9.json 2>/dev/null; echo ref rc=$?
cut -c1-200 gpurun_out/r02_c38_bench_reference_v9.json
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c38_bench_ours.json 2> gpurun_out/r02_c38_bench_ours.err; echo bench rc=$?
cut -c1-300 gpurun_out/r02_c38_bench_ours.json

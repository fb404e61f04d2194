# Last check of the tree as committed: GPU suite, smoke, one bench line (compute-sanitizer is closed on this pool)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/sanity_pytest_gpu.log 2>&1; echo pytest rc=$?
tail -n 3 gpurun_out/sanity_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 200 python bench.py --steps 50 --warmup 5 > gpurun_out/sanity_bench.json 2> gpurun_out/sanity_bench.err; echo bench rc=$?
tail -n 1 gpurun_out/sanity_bench.json | cut -c1-160

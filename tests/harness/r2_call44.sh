# Round 2, call 44: final tree with the zero-copy O store as the default of flash_attn_fwd_host: GPU suite, smoke, both bench arms
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c44_pytest_gpu.log 2>&1; echo pytest rc=$?
tail -n 3 gpurun_out/r02_c44_pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -n 1 | tee gpurun_out/r02_c44_smoke.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_c44_bench_reference_v9.json 2>/dev/null; echo ref rc=$?
python tests/harness/print_value.py reference gpurun_out/r02_c44_bench_reference_v9.json e2e | cut -c1-260
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_c44_bench_ours.json 2> gpurun_out/r02_c44_bench_ours.err; echo bench rc=$?
python tests/harness/print_value.py ours gpurun_out/r02_c44_bench_ours.json e2e | cut -c1-260

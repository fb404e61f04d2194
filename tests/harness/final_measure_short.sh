# Round-end measurement pass on one B200, trimmed to ~5 GPU-minutes (everything lands in gpurun_out/, tag = $1)
T=${1:-r01}
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo pytest rc=$?
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo smoke rc=$?
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference_v9.json 2> gpurun_out/${T}_bench_reference_v9.err; echo ref rc=$?
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/${T}_bench_ours.json 2> gpurun_out/${T}_bench_ours.err; echo ours rc=$?
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 2 --workload cfg2_n8192_full > gpurun_out/${T}_bench_ours_full.json 2>/dev/null
timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 2 --workload cfg4_d64_n2048_full > gpurun_out/${T}_bench_ours_cfg4.json 2>/dev/null
timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --e2e-steps 2 --workload cfg1_n1024_causal > gpurun_out/${T}_bench_ours_cfg1.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/${T}_ncu_launches.log 2>&1; echo ncu-list rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_kernel -s 3 -c 1 -o gpurun_out/${T}_causal_n8192 -f python tests/harness/profile_one.py 1 32 8192 128 1 5 > gpurun_out/${T}_ncu_full.log 2>&1; echo ncu-full rc=$?
timeout 200 ./flash_attention 1024 1 > gpurun_out/${T}_cli_1024_causal.log 2>&1; echo cli rc=$?
tail -n 3 gpurun_out/${T}_pytest_gpu.log; tail -n 2 gpurun_out/${T}_smoke.log; tail -n 12 gpurun_out/${T}_cli_1024_causal.log
for f in ours ours_full ours_cfg4 ours_cfg1 reference_v9; do tail -n 1 gpurun_out/${T}_bench_$f.json | cut -c1-120; done

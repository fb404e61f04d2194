# Round 2, final 1-GPU pass: the GPU suite, smoke, the bench records, the reference-style harness, ncu launch list + full capture
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo pytest rc=$?
tail -n 6 gpurun_out/r02_final_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2 | tee gpurun_out/r02_final_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_ours.json 2> gpurun_out/r02_final_bench_ours.err; echo bench rc=$?
cut -c1-400 gpurun_out/r02_final_bench_ours.json
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_bench_reference_v9.json 2>/dev/null; echo ref rc=$?
cut -c1-300 gpurun_out/r02_final_bench_reference_v9.json
for wl in cfg2_n8192_full cfg1_n1024_causal cfg4_d64_n2048_full cfg2_n16384_causal cfg2_n2048_causal; do
  timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --workload $wl > gpurun_out/r02_final_bench_ours_$wl.json 2>/dev/null; echo $wl rc=$?
  cut -c1-200 gpurun_out/r02_final_bench_ours_$wl.json
done
timeout 300 python tests/harness/sdpa_compare.py 2>&1 | tee gpurun_out/r02_final_sdpa_compare.log
timeout 900 ./flash_attention > gpurun_out/r02_final_cli_full_harness.log 2>&1; echo cli rc=$?
tail -n 40 gpurun_out/r02_final_cli_full_harness.log
# ncu: launch list of a short bench run, then one full capture of the dominant kernel (plain runs first, same command lines)
python bench.py --steps 3 --warmup 3 --sustain-s 0 --no-cpu-baseline --e2e-steps 2 > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --sustain-s 0 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_bench.log 2>&1; echo ncu-list rc=$?
python tests/harness/profile_one.py 1 32 8192 128 1 5 > gpurun_out/plain_profile_one.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fa_fwd_kernel -s 3 -c 1 -o gpurun_out/r02_final_causal_n8192 \
    python tests/harness/profile_one.py 1 32 8192 128 1 5 > gpurun_out/ncu_full.log 2>&1; echo ncu-full rc=$?
tail -n 3 gpurun_out/ncu_full.log

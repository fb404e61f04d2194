# Round 2, call 19 (8 GPUs): final-tree bench record with every multi-GPU leg, per-rank breakdown of a gathered step, and the
# gathered form with its two kernels on two streams (FLASH_ATTN_GATHER_OVERLAP=1)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_final_8gpu_bench.json 2> gpurun_out/r02_final_8gpu_bench.err; echo bench rc=$?
timeout 200 $TR --master-port 29702 tests/harness/gather_breakdown.py 10 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$" | tee gpurun_out/r02_final_8gpu_gather_breakdown.log
for rep in 1 2; do for ov in 0 1; do
  FLASH_ATTN_GATHER_OVERLAP=$ov timeout 200 $TR --master-port 2971$ov bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --workload cfg5_ring_n131072_causal --ring-exchange gather 2>/dev/null > gpurun_out/tmp_ov.json
  python tests/harness/print_value.py "overlap=$ov rep=$rep" gpurun_out/tmp_ov.json | tee -a gpurun_out/r02_final_8gpu_gather_overlap.log
done; done
python tests/harness/print_value.py "bench" gpurun_out/r02_final_8gpu_bench.json cp_cfg5 cp_parity strong_cfg3 e2e
tail -n 3 gpurun_out/r02_final_8gpu_bench.err

# Round 2, call 24 (4 GPUs): final-tree bench record with every multi-GPU leg
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_final_4gpu_bench.json 2> gpurun_out/r02_final_4gpu_bench.err; echo bench rc=$?
python tests/harness/print_value.py "bench" gpurun_out/r02_final_4gpu_bench.json cp_cfg5 cp_parity strong_cfg3
tail -n 3 gpurun_out/r02_final_4gpu_bench.err

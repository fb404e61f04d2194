set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 8 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/mg8_default.json 2> gpurun_out/mg8_default.err; echo default rc=$?
timeout 300 $TR bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline --workload cfg3_b16_n8192_causal > gpurun_out/mg8_cfg3.json 2> gpurun_out/mg8_cfg3.err; echo cfg3 rc=$?
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --workload cfg5_ring_n131072_causal > gpurun_out/mg8_cfg5.json 2> gpurun_out/mg8_cfg5.err; echo cfg5 rc=$?
timeout 300 $TR tests/harness/ring_check.py 8192 > gpurun_out/mg8_ring_check.log 2>&1; echo ringcheck rc=$?
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --workload cfg3_b16_n8192_causal > gpurun_out/mg1_cfg3.json 2> gpurun_out/mg1_cfg3.err; echo cfg3x1 rc=$?
tail -n 1 gpurun_out/mg8_default.json gpurun_out/mg8_cfg3.json gpurun_out/mg8_cfg5.json gpurun_out/mg1_cfg3.json | cut -c1-400
tail -n 4 gpurun_out/mg8_ring_check.log

# 4-GPU scaling points: cfg5 with peer pulls, default workload (weak)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517"
timeout 200 $TR bench.py --gpus 4 --steps 5 --warmup 2 --no-cpu-baseline --workload cfg5_ring_n131072_causal --ring-exchange pull > gpurun_out/pull4_cfg5_pull.json 2> gpurun_out/pull4_cfg5_pull.err; echo cfg5 pull rc=$?
tail -n 1 gpurun_out/pull4_cfg5_pull.json | cut -c1-200
timeout 200 $TR bench.py --gpus 4 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/mg4_default.json 2> gpurun_out/mg4_default.err; echo default rc=$?
tail -n 1 gpurun_out/mg4_default.json | cut -c1-200

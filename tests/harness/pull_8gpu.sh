# 8-GPU run of the peer-pull context-parallel path: numerics at world 8, then cfg5 (N=131072 causal)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 200 $TR tests/harness/ring_check.py 8192 pull > gpurun_out/pull8_ring_check.log 2>&1; echo ringcheck rc=$?
grep -c PASS gpurun_out/pull8_ring_check.log; grep -h "FAIL\|Error" gpurun_out/pull8_ring_check.log | head -5
timeout 200 $TR bench.py --gpus 8 --steps 5 --warmup 2 --no-cpu-baseline --workload cfg5_ring_n131072_causal --ring-exchange pull > gpurun_out/pull8_cfg5_pull.json 2> gpurun_out/pull8_cfg5_pull.err; echo cfg5 pull rc=$?
tail -n 1 gpurun_out/pull8_cfg5_pull.json | cut -c1-260
tail -n 3 gpurun_out/pull8_cfg5_pull.err

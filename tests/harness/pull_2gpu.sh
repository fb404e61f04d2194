# 2-GPU validation of the peer-pull context-parallel path: parity (pytest + ring_check), then cfg5 with both exchanges
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "peer or pull" > gpurun_out/pull2_pytest.log 2>&1; echo pytest rc=$?
tail -n 5 gpurun_out/pull2_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
for ex in pull sendrecv; do
  timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 2 --no-cpu-baseline --workload cfg5_ring_n131072_causal --ring-exchange $ex > gpurun_out/pull2_cfg5_$ex.json 2> gpurun_out/pull2_cfg5_$ex.err; echo cfg5 $ex rc=$?
  tail -n 1 gpurun_out/pull2_cfg5_$ex.json | cut -c1-300
  tail -n 3 gpurun_out/pull2_cfg5_$ex.err
done

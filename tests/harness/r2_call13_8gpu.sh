# Round 2, thirteenth GPU call (8 GPUs): gathered context parallelism, one copy stream vs odd/even hops on two
set -x
mkdir -p gpurun_out
for rep in 1 2; do for st in 1 2; do
  echo "streams=$st rep=$rep"
  FLASH_ATTN_GATHER_STREAMS=$st timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2957$st bench.py --gpus 8 --workload cfg5_ring_n131072_causal --ring-exchange gather --steps 5 2>/dev/null | cut -c1-160
done; done 2>&1 | tee gpurun_out/r02_c13_gather_streams.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29579 bench.py --gpus 8 --workload cfg5_heads_n131072_causal --steps 5 --no-cpu-baseline --sustain-s 0 --e2e-steps 2 2>/dev/null | cut -c1-200 | tee -a gpurun_out/r02_c13_gather_streams.log

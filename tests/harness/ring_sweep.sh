# ring context parallelism (cfg 5) with different amounts of room left for NCCL.
#   bash tests/harness/ring_sweep.sh NGPUS "margin[:nccl_max_ctas] ..."     results -> gpurun_out/ring${N}_m*.json
N=${1:-4}
CASES=${2:-"0 8:8 16"}
mkdir -p gpurun_out
for c in $CASES; do
  m=${c%%:*}; x=""; [ "$c" != "$m" ] && x=${c##*:}
  out=gpurun_out/ring${N}_m${m}${x:+_ctas$x}
  env FLASH_ATTN_RING_SM_MARGIN=$m ${x:+NCCL_MAX_CTAS=$x} timeout 240 python -m torch.distributed.run --nnodes=1 \
      --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 2 \
      --no-cpu-baseline --workload cfg5_ring_n131072_causal > $out.json 2> $out.err
  echo "case $c rc=$? $(tail -n 1 $out.json | cut -c1-130)"
done

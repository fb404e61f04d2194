# Round 2, ninth GPU call (8 GPUs): the driver's scaling run as it will be launched, plus the in-kernel peer-read probe
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -n 8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_c9_bench_8gpu.json 2> gpurun_out/r02_c9_bench_8gpu.err; echo bench rc=$?
tail -c 5000 gpurun_out/r02_c9_bench_8gpu.json; tail -n 15 gpurun_out/r02_c9_bench_8gpu.err
CUDA_VISIBLE_DEVICES=0,1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tests/harness/peer_tma_probe.py 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" | tee gpurun_out/r02_c9_peer_tma_probe.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c9_bench_4gpu.json 2> gpurun_out/r02_c9_bench_4gpu.err; echo bench4 rc=$?
tail -c 2500 gpurun_out/r02_c9_bench_4gpu.json

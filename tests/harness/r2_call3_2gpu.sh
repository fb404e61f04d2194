# Round 2, third GPU call (2 GPUs): the multi-GPU legs of bench.py and the 2-GPU parity test
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -x -q -k "context_parallel or peer_block" > gpurun_out/r02_c3_pytest_2gpu.log 2>&1; echo pytest rc=$?
tail -n 5 gpurun_out/r02_c3_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_c3_bench_2gpu.json 2> gpurun_out/r02_c3_bench_2gpu.err; echo bench rc=$?
tail -c 4000 gpurun_out/r02_c3_bench_2gpu.json; tail -n 20 gpurun_out/r02_c3_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_c3_bench_2gpu_ref.json 2> gpurun_out/r02_c3_bench_2gpu_ref.err; echo ref rc=$?
tail -c 1500 gpurun_out/r02_c3_bench_2gpu_ref.json; tail -n 5 gpurun_out/r02_c3_bench_2gpu_ref.err

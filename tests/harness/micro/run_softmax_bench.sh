# tests/harness/micro/softmax_bench.cu, three builds of the kernel's softmax_tile; binaries are built into build/ beforehand
mkdir -p gpurun_out
for n in base sumguard fence; do ./build/softmax_bench_$n 935 "== $n"; done 2>&1 | tee gpurun_out/softmax_bench.log

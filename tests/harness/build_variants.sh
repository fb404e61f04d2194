# Builds the experimental variants of the library and the softmax microbenchmark into build/ (git-ignored, travels with gpurun).
#   bash tests/harness/build_variants.sh
set -e
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p build
for v in "default:" "stream:-DFA_STREAM" "timing:-DFA_TIMING" $EXTRA_VARIANTS; do
  n=${v%%:*}; f=${v#*:}
  $NV $f -shared flash_attention_cuda_b200/csrc/fa_api.cu -o build/lib_$n.so &
done
for v in "base:" "stream:-DBENCH_STREAM"; do
  n=${v%%:*}; f=${v#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -maxrregcount=208 $f -o build/softmax_bench_$n tests/harness/micro/softmax_bench.cu &
done
wait
ls -la build/lib_*.so build/softmax_bench_*
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/softmax_half_bench tests/harness/micro/softmax_half_bench.cu

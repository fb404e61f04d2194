# Builds the experimental variants of the library and the softmax microbenchmark into build/ (git-ignored, travels with gpurun).
#   bash tests/harness/build_variants.sh
set -e
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p build
cp flash_attention_cuda_b200/libflashattn_b200.so build/lib_default.so
for v in "stream:-DFA_STREAM_S" "single:-DFA_SINGLE_TILE_MODE" "sumguard:-DFA_SUM_GUARD" "fence:-DFA_SCHED_FENCE"; do
  n=${v%%:*}; f=${v#*:}
  $NV $f -shared flash_attention_cuda_b200/csrc/fa_api.cu -o build/lib_$n.so &
done
for v in "base:" "stream:-DFA_STREAM_S" "sumguard:-DFA_SUM_GUARD" "fence:-DFA_SCHED_FENCE"; do
  n=${v%%:*}; f=${v#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -maxrregcount=208 $f -o build/softmax_bench_$n tests/harness/micro/softmax_bench.cu &
done
wait
ls -la build/lib_*.so build/softmax_bench_*

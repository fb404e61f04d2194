# Round 2, call 37: ncu --set full captures of the D=64 kernel (cfg 4), the non-causal N=8192 launch and the split-mode cfg 1 launch
set -x
mkdir -p gpurun_out
for spec in "cfg4_d64:32 16 2048 64 0" "full_n8192:1 32 8192 128 0" "cfg1_n1024:1 32 1024 128 1"; do
  name=${spec%%:*}; args=${spec#*:}
  python tests/harness/profile_one.py $args 5 > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:fa_fwd_kernel -s 3 -c 1 -o gpurun_out/r02_final_$name \
      python tests/harness/profile_one.py $args 5 > gpurun_out/ncu_$name.log 2>&1; echo $name rc=$?
  tail -n 1 gpurun_out/plain_$name.log
done

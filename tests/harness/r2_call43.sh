# Round 2, call 43: in-process A/B of the host pipeline variants, twice (two processes)
set -x
mkdir -p gpurun_out
for rep in 1 2; do timeout 120 python tests/harness/host_ab.py 8 2>&1 | tee -a gpurun_out/r02_c43_host_ab.log; done

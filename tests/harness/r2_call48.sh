# Round 2, call 48: chunk kernels on a narrow grid under the direct O store (in-process A/B, output checked against the first variant)
mkdir -p gpurun_out
timeout 60 python tests/harness/host_ab.py 6 ctas 2>&1 | tee gpurun_out/r02_c48_host_ab_ctas.log

# Round 2, call 47: the library as finally built (chunk boundaries refactored into host_chunk_bounds): host-entry GPU tests + smoke
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -x -q -k "host or reference_harness or cli_single" > gpurun_out/r02_c47_pytest_host.log 2>&1; echo pytest rc=$?
tail -n 2 gpurun_out/r02_c47_pytest_host.log
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -n 1

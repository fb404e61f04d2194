"""Sustained A/B of library builds in ONE process sequence (same box, power-capped regime):
   python tests/harness/ab.py lib1.so lib2.so ...   -> TFLOPS for causal/full N=8192 (300 steps each), 2 rounds"""
import json
import os
import subprocess
import sys

libs = sys.argv[1:]
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for rnd in range(2):
    for lib in libs:
        row = []
        for wl in ("cfg2_n8192_causal", "cfg2_n8192_full"):
            env = dict(os.environ, FLASH_ATTN_B200_LIB=lib)
            out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--no-cpu-baseline", "--steps", "300",
                                  "--e2e-steps", "2", "--workload", wl], env=env, capture_output=True, text=True).stdout
            try:
                d = json.loads(out.strip().splitlines()[-1])
                row.append(f"{wl.split('_')[-1]} {d['value']:7.1f} ({d['clocks']['sm_mhz']} MHz {','.join(d['clocks']['reasons'])})")
            except Exception as e:
                row.append(f"{wl}: ERROR {e} {out[-200:]}")
        print(f"round {rnd} {os.path.basename(lib):28s} " + " | ".join(row), flush=True)

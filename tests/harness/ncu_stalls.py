"""Summarises an .ncu-rep (run here, no GPU needed): key raw metrics + stall samples per code region.
   python tests/harness/ncu_stalls.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum"]
for h, u, v in zip(hdr, units, vals):
    if h in keys:
        print(f"{h} [{u}] = {v}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
per = []
for i, r in enumerate(data):
    n = int(r[idx["# Samples"]] or 0)
    st = {s: int(r[idx[s]] or 0) for s in stalls}
    for s in stalls:
        tot[s] += st[s]
    per.append((n, i, r[idx["Source"]], st, int(r[idx["Instructions Executed"]] or 0)))
T = sum(p[0] for p in per)
print("total samples", T, "instructions", len(data))
for s, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
    print(f"  {s:28s} {v:8d} {100 * v / T:5.1f}%")
print("top instructions:")
for n, i, s, st, ie in sorted(per, key=lambda x: -x[0])[:topn]:
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"{n:6d} {100 * n / T:4.1f}% line={i:5d} ex={ie:8d} {s[:64]:64s} {top}")
# opcode histogram weighted by executed count
ops = {}
for n, i, s, st, ie in per:
    op = s.replace("@P0", "").replace("@!P0", "").split()
    op = [x for x in op if not x.startswith("@")]
    if op:
        ops[op[0].split(".")[0]] = ops.get(op[0].split(".")[0], 0) + ie
print("executed warp-instructions by opcode (top 25):")
for k, v in sorted(ops.items(), key=lambda x: -x[1])[:25]:
    print(f"  {k:12s} {v}")

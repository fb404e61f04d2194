"""Per-source-line stall samples of an .ncu-rep captured with -lineinfo / --import-source on.
   python tests/harness/ncu_lines.py rep.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
cur_file = ""
agg = []
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) >= len(hdr) - 1 and r[0].isdigit():
        i_s = hdr.index("# Samples")
        i_e = hdr.index("Instructions Executed")
        st = {h: num(r[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and i < len(r)}
        agg.append((num(r[i_s]), cur_file, int(r[0]), r[1].strip()[:90], num(r[i_e]), st))
T = sum(a[0] for a in agg)
print("total samples", T)
for n, f, ln, src, ie, st in sorted(agg, key=lambda x: -x[0])[:topn]:
    top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:2] if v)
    print(f"{n:6d} {100.0 * n / T:5.1f}%  {f}:{ln:<4d} ex={ie:<9d} {src}   [{top}]")

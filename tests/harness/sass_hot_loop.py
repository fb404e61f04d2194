"""SASS listing of the softmax hot loop (unmasked 128x128 tile of fa_fwd_kernel<128, 1, false>) with the scheduling fields
ptxas encoded in each instruction's control word (Volta+ 128-bit encoding: bits 105-108 stall cycles, 109 yield,
110-112 / 113-115 write / read scoreboard, 116-121 wait mask).  Runs here, no GPU needed.
   python tests/harness/sass_hot_loop.py [lib.so] > profiles/rNN_softmax_hot_loop.sass.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "flash_attention_cuda_b200", "libflashattn_b200.so")
KERNEL = "fa_fwd_kernelILi128ELi1ELb0"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if "Function :" in l and KERNEL in l)
end = next((i for i in range(start + 1, len(sass)) if "Function :" in sass[i]), len(sass))
ins = []      # (address, text, high word)
for i in range(start, end):
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?)\s*;\s*/\* 0x([0-9a-f]+) \*/", sass[i])
    if m and i + 1 < end:
        h = re.match(r"\s*/\* 0x([0-9a-f]+) \*/", sass[i + 1])
        ins.append((m.group(1), m.group(2), int(h.group(1), 16) if h else 0))
# candidate bodies: from a group of tcgen05.ld of S (LDTM.x32 x3 back to back) to the second SYNCS.ARRIVE behind it
bodies = []
is_ld = [t.startswith("LDTM.x32") for _, t, _ in ins]
for i in range(len(ins) - 64):
    # the four loads of S: an LDTM.x32 with none in the 12 instructions before it and three more within the next 64
    if is_ld[i] and not any(is_ld[max(0, i - 12):i]) and sum(is_ld[i + 1:i + 64]) >= 3:
        arrives, j = 0, i
        while j < len(ins) and arrives < 2:
            arrives += "SYNCS.ARRIVE" in ins[j][1]
            j += 1
        bodies.append((i, j))
# the unmasked variant is the one without the column-limit selects
i0, i1 = min(bodies, key=lambda b: sum(1 for x in ins[b[0]:b[1]] if x[1].startswith(("FSEL", "ISETP", "SEL"))))
body = ins[i0:i1]
ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t, _ in body)
stall = lambda h: (h >> 41) & 0xF
ver = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().split("\n")[-2]
print(f"SASS of the softmax hot loop (unmasked 128x128 tile, fa::softmax_tile inlined into fa_fwd_kernel<128,1,false>), {os.path.basename(lib)}.")
print("Extracted by tests/harness/sass_hot_loop.py (cuobjdump -sass); columns: address, instruction, then the scheduling fields of the")
print("control word: stall cycles, yield, write / read scoreboard, wait mask.  From the tcgen05.ld of S to the arrival on p_full[piece 1];")
print("the rarely taken rescale block (rolled loop: LDTM / FMUL2 / STTM of O) and the cold watchdog path of its wait sit inside the range.")
print(f"nvcc: {ver}")
print(f"{len(body)} instructions, sum of encoded stall cycles {sum(stall(h) for _, _, h in body)} (floor for one warp alone, cold blocks included)")
print("opcode histogram: " + ", ".join(f"{k} {v}" for k, v in ops.most_common()))
print()
for a, t, h in body:
    print(f"/*{a}*/ {t:88s} stall={stall(h):2d} y={(h >> 45) & 1} wb={(h >> 46) & 7} rb={(h >> 49) & 7} wait={(h >> 52) & 0x3F:06b}")

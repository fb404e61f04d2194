"""Where a step of gathered context parallelism (ring.gather_attention_forward, BASELINE config 5) spends its time, per rank.
Run under torchrun on 2 / 4 / 8 GPUs.  Per rank and step, CUDA events on the compute stream bracket
   begin   own chunks in place, flags reset, "every block is in place" all-reduce, pulls queued
   low     kernel of the low Q chunk   (first r + 1 slots)
   high    kernel of the high Q chunk  (all 2P - r slots; its first wave of work items is paced by the arrival of the slots)
   end     join with the copy stream + "everyone is done reading" all-reduce
and the pulls are timed on their own stream (first copy queued -> last flag raised).  Printed: one line per rank (ms, mean over
the timed steps) and the step time as bench.py measures it.
   python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tests/harness/gather_breakdown.py [steps]"""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import flash_attention_cuda_b200 as fa  # noqa: E402
from flash_attention_cuda_b200 import ring  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
B, H, N, D = 1, 32, 131072, 128
C = N // (2 * world)
dev = torch.device("cuda", lr)
gk = ring.gathered_kv(B, H, C, D, dev)
g = torch.Generator(device="cuda").manual_seed(7 + rank)
q = [(torch.rand((B, H, C, D), device="cuda", generator=g) - 0.5).half() for _ in range(2)]
for t in gk.k + gk.v:
    t.copy_((torch.rand((B, H, C, D), device="cuda", generator=g) - 0.5).half())
out = [torch.empty_like(q[0]), torch.empty_like(q[1])]
n = gk.nslots
names = ["begin", "low", "high", "end"]
acc = {k: 0.0 for k in names + ["pulls", "step"]}


def ev():
    return torch.cuda.Event(enable_timing=True)


def one_step(timed):
    e = [ev() for _ in range(5)]
    p0, p1 = ev(), ev()
    cur = torch.cuda.current_stream(dev)
    e[0].record()
    # GatheredKV.begin, with the copy stream bracketed
    gk.begin(gk.k, gk.v)
    p1.record(gk.comm)                 # behind the last flag write of the pulls queued by begin()
    e[1].record()
    rr = C // gk.parts
    fa.flash_attn_fwd_gathered(q[0], gk.k_ptr, gk.v_ptr, out[0], (gk.rank + 1) * C, n * C, True, gk.rank * C, gk.flags, rr)
    e[2].record()
    fa.flash_attn_fwd_gathered(q[1], gk.k_ptr, gk.v_ptr, out[1], n * C, n * C, True, (n - 1) * C, gk.flags, rr)
    e[3].record()
    gk.end()
    e[4].record()
    if timed:
        torch.cuda.synchronize()
        for i, k in enumerate(names):
            acc[k] += e[i].elapsed_time(e[i + 1])
        acc["pulls"] += e[1].elapsed_time(p1)      # from the end of begin() on the compute stream to the last pull landed
        acc["step"] += e[0].elapsed_time(e[4])


for _ in range(3):
    one_step(False)
torch.cuda.synchronize()
dist.barrier()
for _ in range(steps):
    one_step(True)
line = (f"rank {rank}: " + "  ".join(f"{k} {acc[k] / steps:7.3f}" for k in names) +
        f"  | last pull lands {acc['pulls'] / steps:6.3f} ms after begin  | step {acc['step'] / steps:7.3f} ms"
        f"  ({gk.nslots - 2} slots pulled)")
lines = [None] * world
dist.all_gather_object(lines, line)
if rank == 0:
    print(f"gathered context parallelism, B{B} H{H} N{N} D{D} causal, {world} GPUs, {steps} steps (each step synchronised for the events)")
    for ln in lines:
        print(ln)
ring.release_peer_kv()
dist.destroy_process_group()

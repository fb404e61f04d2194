"""How long does one ring hop's K/V exchange take alone, and next to a running attention kernel?
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/harness/ring_comm_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import flash_attention_cuda_b200 as fa  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
B, H, C, D = 1, 32, 8192, 128          # one chunk of the P=8 ring: 64 MiB per tensor, 256 MiB per hop
g = torch.Generator(device="cuda").manual_seed(rank)
q, k0, k1, v0, v1 = ((torch.rand((B, H, C, D), device="cuda", generator=g) - 0.5).half() for _ in range(5))
cur = [k0, k1, v0, v1]
nxt = [torch.empty_like(t) for t in cur]
o = torch.empty_like(q)
comm = torch.cuda.Stream()
to, frm = (rank + 1) % world, (rank - 1) % world


def exchange():
    ops = []
    for a, b in zip(cur, nxt):
        ops += [dist.P2POp(dist.isend, a, to), dist.P2POp(dist.irecv, b, frm)]
    comm.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(comm):
        reqs = dist.batch_isend_irecv(ops)
    return reqs


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def comm_only():
    for r in exchange():
        r.wait()
    torch.cuda.current_stream().wait_stream(comm)


def compute_only():
    for kk, vv in ((k0, v0), (k1, v1)):
        fa.flash_attn_fwd(q, kk, vv, causal=False, out=o)


def both():
    reqs = exchange()
    compute_only()
    for r in reqs:
        r.wait()
    torch.cuda.current_stream().wait_stream(comm)


t_comm = timed(comm_only)
res = [f"comm alone {t_comm:.3f} ms ({4 * k0.numel() * 2 / t_comm / 1e6:.0f} GB/s per direction)"]
for margin in (0, 8, 16, 32):
    fa.set_sm_margin(margin)
    res.append(f"margin {margin}: compute {timed(compute_only):.3f} ms, compute+comm {timed(both):.3f} ms")
if rank == 0:
    print("\n".join(res), flush=True)
dist.destroy_process_group()

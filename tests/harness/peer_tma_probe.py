"""SURVEY 8 f2 as written: the kernel's producer warp TMA-loads K/V straight from the neighbour's HBM over NVLink
(the tensor maps are simply encoded on the mapped peer address), against the shipped form (copy engine pulls the
chunk into local HBM, the kernel reads it locally).  One hop of the P=8 context-parallel schedule on 2 GPUs:
Q chunk [1, 32, 8192, 128] against one K/V chunk of the same size (64 MiB + 64 MiB).
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/harness/peer_tma_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import flash_attention_cuda_b200 as fa  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B, H, C, D = 1, 32, 8192, 128
nbytes = B * H * C * D * 2
blk, handle, ptr = fa.peer_alloc(2 * nbytes, dev)
mine = blk.view(torch.float16).view(2, B, H, C, D)
g = torch.Generator(device="cuda").manual_seed(7 + rank)
mine.copy_(torch.rand(mine.shape, device="cuda", generator=g) - 0.5)
q = (torch.rand((B, H, C, D), device="cuda", generator=g) - 0.5).half()
handles = [None] * world
dist.all_gather_object(handles, handle)
other = (rank + 1) % world
pptr = fa.peer_open(handles[other])
peer = fa._as_tensor(pptr, 2 * nbytes, dev).view(torch.float16).view(2, B, H, C, D)   # the neighbour's block, mapped
land = torch.empty_like(mine)
o_part = torch.empty(B * H * C, D, dtype=torch.float32, device="cuda")
ml = torch.empty(B * H * C, 2, dtype=torch.float32, device="cuda")
o = torch.empty_like(q)
side = torch.cuda.Stream()
torch.cuda.synchronize(); dist.barrier()


def timed(fn, n=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


res = {}
for causal in (False, True):
    tag = "diagonal (causal)" if causal else "unmasked"
    res[f"{tag}: K/V in local HBM"] = timed(lambda: fa.flash_attn_fwd_partial(q, mine[0], mine[1], o_part, ml, causal, 0, 0, False))
    res[f"{tag}: K/V read by TMA from the peer's HBM"] = timed(lambda: fa.flash_attn_fwd_partial(q, peer[0], peer[1], o_part, ml, causal, 0, 0, False))
res["copy engine pull of the chunk (128 MiB) alone"] = timed(lambda: fa.peer_copy(land.data_ptr(), pptr, 2 * nbytes))


def pull_then_local():       # the shipped form, no overlap: pull, then the kernel on the landing buffer
    fa.peer_copy(land.data_ptr(), pptr, 2 * nbytes)
    fa.flash_attn_fwd_partial(q, land[0], land[1], o_part, ml, False, 0, 0, False)


def pull_overlapped():       # the shipped form as the driver runs it: the pull of the next hop under this hop's kernel
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fa.peer_copy(land.data_ptr(), pptr, 2 * nbytes, side)
    fa.flash_attn_fwd_partial(q, mine[0], mine[1], o_part, ml, False, 0, 0, False)
    torch.cuda.current_stream().wait_stream(side)


res["unmasked: pull, then kernel on the landing buffer (serial)"] = timed(pull_then_local)
res["unmasked: kernel on local K/V with the next pull in flight"] = timed(pull_overlapped)
# correctness of the peer-read form: same partial state as from a local copy of the same bytes
fa.peer_copy(land.data_ptr(), pptr, 2 * nbytes)
fa.flash_attn_fwd_partial(q, land[0], land[1], o_part, ml, True, 0, 0, False)
ref = o_part.clone()
fa.flash_attn_fwd_partial(q, peer[0], peer[1], o_part, ml, True, 0, 0, False)
torch.cuda.synchronize()
same = bool(torch.equal(ref, o_part))
if rank == 0:
    fl = 4.0 * B * H * C * C * D
    print(f"one context-parallel hop at P=8: Q chunk x K/V chunk of [1, 32, 8192, 128], {world} GPUs, max over ranks")
    for k, v in res.items():
        extra = ""
        if "K/V" in k or "kernel" in k:
            extra = f"   {fl / (2 if 'causal' in k else 1) / v / 1e9:7.1f} TFLOPS"
        if "copy engine" in k:
            extra = f"   {2 * nbytes / v / 1e6:7.1f} GB/s"
        print(f"  {k:62s} {v:8.3f} ms{extra}")
    print(f"  peer-read partial state bit-identical to the local-copy one: {same}")
torch.cuda.synchronize(); dist.barrier()
fa.peer_close(pptr)
dist.barrier()
del mine, peer, blk
fa.peer_free(ptr)
dist.destroy_process_group()

#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: FlashAttention forward, FP16, on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-cpu]
                    [--workload NAME] [--sweep] [--sustain-s S] [--flush-l2] [--no-multi]

One "step" = one forward pass over one synthetic batch of the named workload (default: the shape
BASELINE.json's metric is quoted on -- B=1 H=32 D=128 seq=8192 causal, the README table's shape at
the north-star target length).  Prints ONE JSON line (rank 0):

  value     whole-job forward TFLOPS (FLOPs = 4*B*H*N^2*D, /2 causal -- the reference's convention,
            flash_attention.cu:938-939), inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       the same metric through the C-ABI call with HOST buffers (flash_attn_fwd_host: H2D of
            Q,K,V from pinned memory + kernel + D2H of O inside the timed region); `copy_only_ms` is the
            same byte traffic with no kernel in between, all ranks at once: the host-side floor of the call
  roofline  tensor-bound: achieved TFLOPS of the kernel vs the measured cuBLAS bf16 peak (burst figure:
            the timed region is milliseconds long)
  sustained a separate leg of >= --sustain-s seconds of back-to-back launches with its own clocks record,
            against the measured SUSTAINED cuBLAS figure
  cpu_baseline  the CPU oracle (port of the reference's cpu_attention) on a bounded row sample

Under torchrun (N > 1) the line additionally carries, measured in the same process group:
  strong_cfg3   BASELINE config 3 (B16 H32 N8192 causal) with its 512 heads split over the ranks, no collective
                (reference flash_attention.cu:120-122: heads are independent)
  cp_cfg5       BASELINE config 5 (B1 H32 N131072 causal) context-parallel over the ranks: `gather` (copy-engine pulls
                into one gathered K/V buffer, two kernels per rank), `pull` (the same pulls, partial states + merge)
                and `sendrecv` (NCCL ring), next to the no-communication bound (the same shape with its heads split)
  cp_parity     context-parallel output rows of every rank against the CPU oracle (gate 2e-3 / 2e-4;
                the process exits non-zero when it fails)

`--impl reference` runs the UNMODIFIED reference kernel (V9, flash_attention_v9_dispatch) rebuilt for
sm_100a from /root/reference into oracle/_ref/libref_v9.so, on the same workload with the same timing
code.  The reference's implementation of this path is a GPU kernel, so that is what the reference arm
times; its CPU check function (cpu_attention) is reported beside it as `cpu_baseline`
(`--impl reference-cpu` times only that).
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

WORKLOADS = {
    # name: (B, H, N, D, causal)
    "cfg1_n1024_causal": (1, 32, 1024, 128, 1),          # BASELINE.json configs[0]
    "cfg2_n512_causal": (1, 32, 512, 128, 1),
    "cfg2_n2048_causal": (1, 32, 2048, 128, 1),
    "cfg2_n4096_causal": (1, 32, 4096, 128, 1),
    "cfg2_n8192_causal": (1, 32, 8192, 128, 1),          # configs[1], headline
    "cfg2_n8192_full": (1, 32, 8192, 128, 0),
    "cfg2_n16384_causal": (1, 32, 16384, 128, 1),
    "cfg2_n16384_full": (1, 32, 16384, 128, 0),
    "cfg3_b16_n8192_causal": (16, 32, 8192, 128, 1),     # configs[2] (strong scaling: B*H sharded)
    "cfg4_d64_n2048_full": (32, 16, 2048, 64, 0),        # configs[3]
    "cfg5_ring_n131072_causal": (1, 32, 131072, 128, 1),  # configs[4]: context parallel over the ranks
    # the same shape with its 32 heads split across the ranks instead of its sequence: the no-communication upper bound
    # SURVEY 8(e) asks to be reported next to the context-parallel number
    "cfg5_heads_n131072_causal": (1, 32, 131072, 128, 1),
}
STRONG_BH_SHARD = ("cfg3_b16_n8192_causal", "cfg5_heads_n131072_causal")
DEFAULT_WORKLOAD = "cfg2_n8192_causal"
NOMINAL_FP16_TFLOPS = 2250.0
L2_BYTES = 126e6
# dram__bytes_read.sum + dram__bytes_write.sum per launch of OUR kernel, from the committed `ncu --set full` capture
NCU_CAPTURES = {"cfg2_n8192_causal": "r02_final_causal_n8192_summary.txt"}


def flops(B, H, N, D, causal):
    f = 4.0 * B * H * N * N * D
    return f / 2 if causal else f


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            m = json.load(f)
        return {"tflops": float(m["bf16_tflops"]), "tflops_sustained": float(m.get("bf16_tflops_sustained", 0)),
                "hbm_gbs": float(m["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while a timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for n in dir(nv):
            if n.startswith("nvmlClocksEventReason") or n.startswith("nvmlClocksThrottleReason"):
                v = getattr(nv, n)
                if isinstance(v, int) and v not in (0,):
                    names.setdefault(v, n.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if bit and (bit & (bit - 1)) == 0 and mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        s = sorted(self.samples)
        pretty = {"GpuIdle": "gpu_idle", "ApplicationsClocksSetting": "applications_clocks_setting",
                  "SwPowerCap": "sw_power_cap", "HwSlowdown": "hw_slowdown", "SyncBoost": "sync_boost",
                  "SwThermalSlowdown": "sw_thermal_slowdown", "HwThermalSlowdown": "hw_thermal_slowdown",
                  "HwPowerBrakeSlowdown": "hw_power_brake_slowdown", "DisplayClockSetting": "display_clock_setting"}
        reasons = sorted({pretty.get(r, r) for r in self.reasons} - {"gpu_idle", "None", "All"})
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(s)}


def ncu_traffic(workload_name):
    """(bytes, source file) of the committed `ncu --set full` capture of this workload's kernel, or (None, None)."""
    f = NCU_CAPTURES.get(workload_name)
    if not f:
        return None, None
    try:
        tot = 0.0
        for line in open(os.path.join(REPO, "profiles", f)):
            if line.startswith("dram__bytes_read.sum [Mbyte]") or line.startswith("dram__bytes_write.sum [Mbyte]"):
                tot += float(line.split("=")[1]) * 1e6
        return (tot or None), "profiles/" + f
    except OSError:
        return None, None


def cpu_baseline(workload, budget_rows=None):
    """The CPU oracle (port of cpu_attention, FA.cu:668-697) on a bounded, evenly spread row sample of
    the same workload, all host threads.  TFLOPS-equivalent = sampled rows' FLOPs / wall time."""
    import numpy as np
    import _oracle
    B, H, N, D, causal = workload
    threads = _oracle.max_threads()
    # ~10-20 s of CPU work: per-thread rate is ~0.5 GFLOP/s
    target_flop = 0.5e9 * threads * 12.0
    per_row = flops(1, 1, N, D, causal) / N
    nrows = int(max(threads * 4, min(B * H * N, target_flop / per_row)))
    if budget_rows:
        nrows = budget_rows
    rng = np.random.default_rng(0)
    heads = min(B * H, 4)
    shape = (1, heads, N, D)
    q, k, v = ((rng.random(shape, dtype=np.float32) - 0.5).astype(np.float16) for _ in range(3))
    stride = max(1, (heads * N) // nrows)
    idx = np.arange(0, heads * N, stride)[:nrows]
    bhs, rows = (idx // N).astype(np.int32), (idx % N).astype(np.int32)
    if causal:
        row_flops = float((4.0 * D * (rows.astype(np.float64) + 1)).sum())   # exact work of the sampled rows
    else:
        row_flops = 4.0 * D * N * len(rows)
    t0 = time.perf_counter()
    _oracle.attention_rows(q, k, v, causal, bhs, rows)
    dt = time.perf_counter() - t0
    return {"value": row_flops / dt / 1e12, "unit": "TFLOPS", "cores": threads, "kind": "port",
            "sample": f"{len(rows)} evenly spaced rows of {heads} heads of the workload (N={N}, D={D}, "
                      f"causal={causal}), {row_flops / 1e9:.1f} GFLOP in {dt:.1f} s; oracle/attn_oracle.c, pthreads"}


# ------------------------------------------------------------------------------------------------
# context parallelism (config 5) -- data, timing and the in-run parity check
# ------------------------------------------------------------------------------------------------
CP_SEED = 4242


def cp_chunk(kind, chunk, B, H, C, D, heads=None):
    """Chunk `chunk` (of 2P) of Q (kind 0), K (1) or V (2) of the global sequence: U(-0.5, 0.5) from a generator
    keyed by (kind, chunk) alone, so every rank -- and the parity check on rank 0 -- can regenerate any chunk."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(CP_SEED + 3 * chunk + kind)
    t = (torch.rand((B, H, C, D), device="cuda", generator=g) - 0.5).half()
    return t if heads is None else t[:, heads].contiguous()


def cp_setup(workload, rank, world, local_rank, exchange):
    """This rank's zig-zag share of the global Q, K, V.  pull: K/V are generated straight into the peer-readable block."""
    import torch
    from flash_attention_cuda_b200 import ring
    B, H, N, D, causal = workload
    C = N // (2 * world)
    mine = ring.zigzag_chunks(rank, world)
    q = [cp_chunk(0, c, B, H, C, D) for c in mine]
    if exchange in ("pull", "gather"):
        # K/V are produced straight into the peer-readable block (pull) / the rank's own slots of its gathered buffer
        px = (ring.peer_kv if exchange == "pull" else ring.gathered_kv)(B, H, C, D, torch.device("cuda", local_rank))
        for slot, c in enumerate(mine):
            px.k[slot].copy_(cp_chunk(1, c, B, H, C, D))
            px.v[slot].copy_(cp_chunk(2, c, B, H, C, D))
        k, v = px.k, px.v
    else:
        k = [cp_chunk(1, c, B, H, C, D) for c in mine]
        v = [cp_chunk(2, c, B, H, C, D) for c in mine]
    return q, k, v


def cp_time(workload, rank, world, local_rank, exchange, steps, warm, barrier, max_over_ranks):
    """ms per context-parallel forward (max over ranks) and the last step's output chunks."""
    import torch
    from flash_attention_cuda_b200 import ring
    causal = workload[4]
    q, k, v = cp_setup(workload, rank, world, local_rank, exchange)
    out = None
    for _ in range(warm):
        out = ring.ring_attention_forward(q, k, v, bool(causal), exchange=exchange)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = ring.ring_attention_forward(q, k, v, bool(causal), exchange=exchange)
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1) / steps), out


def cp_parity(workload, rank, world, outs):
    """Sampled output rows of every rank's two chunks against the CPU oracle on the regenerated global sequence.
    `outs`: {exchange: [o_lo, o_hi]} of this rank.  Returns (on rank 0) {exchange: {max_abs, mean_abs}, ...}."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from flash_attention_cuda_b200 import ring
    B, H, N, D, causal = workload
    C = N // (2 * world)
    heads = [0, H - 1] if H > 1 else [0]
    offs = sorted({0, 1, C // 2 - 1, C // 2, C - 2, C - 1})
    mine = ring.zigzag_chunks(rank, world)
    names = sorted(outs)
    # [exchange][slot][head][off] rows of this rank -> one small tensor, gathered on every rank
    sample = torch.stack([torch.stack([outs[n][slot][0, heads][:, offs] for slot in range(2)]) for n in names])
    gathered = [torch.empty_like(sample) for _ in range(world)]
    dist.all_gather(gathered, sample.contiguous())
    if rank != 0:
        return None
    import _oracle
    # the global K, V (sampled heads only) and the sampled Q rows, regenerated chunk by chunk
    hk = np.empty((1, len(heads), N, D), np.float16)
    hv = np.empty((1, len(heads), N, D), np.float16)
    hq = np.zeros((1, len(heads), N, D), np.float16)
    for c in range(2 * world):
        hk[0, :, c * C:(c + 1) * C] = cp_chunk(1, c, B, H, C, D, heads)[0].cpu().numpy()
        hv[0, :, c * C:(c + 1) * C] = cp_chunk(2, c, B, H, C, D, heads)[0].cpu().numpy()
        hq[0, :, c * C:(c + 1) * C] = cp_chunk(0, c, B, H, C, D, heads)[0].cpu().numpy()
    bhs, rows = [], []
    for r in range(world):
        for c in ring.zigzag_chunks(r, world):
            for hi in range(len(heads)):
                for o in offs:
                    bhs.append(hi)
                    rows.append(c * C + o)
    t0 = time.perf_counter()
    ref = _oracle.attention_rows(hq, hk, hv, causal, np.array(bhs, np.int32), np.array(rows, np.int32))
    dt = time.perf_counter() - t0
    res = {"rows_checked": len(rows), "heads": heads, "oracle_s": round(dt, 2),
           "gate": {"max_abs": _oracle.MAX_ABS_TOL, "mean_abs": _oracle.MEAN_ABS_TOL}, "pass": True}
    for ni, n in enumerate(names):
        got = torch.stack([g[ni] for g in gathered]).reshape(-1, D).cpu().numpy()   # rank, slot, head, off: the order above
        mx, mean = _oracle.diff(got, ref)
        ok = bool(mx <= _oracle.MAX_ABS_TOL and mean <= _oracle.MEAN_ABS_TOL)
        res[n] = {"max_abs": mx, "mean_abs": mean, "pass": ok}
        res["pass"] = res["pass"] and ok
    return res


def bench_ring(args, workload, rank, world, local_rank, barrier, max_over_ranks):
    """`--workload cfg5_ring_n131072_causal` on its own: context parallelism with --ring-exchange.
    world == 1 runs the monolithic kernel as the baseline."""
    import torch
    import torch.distributed as dist
    import flash_attention_cuda_b200 as fa
    from flash_attention_cuda_b200 import ring
    B, H, N, D, causal = workload
    steps, warm = max(2, min(args.steps, 5)), 3
    C = N // (2 * world)
    if world == 1:
        g = torch.Generator(device="cuda").manual_seed(1234)
        q, k, v = ((torch.rand((B, H, N, D), device="cuda", generator=g) - 0.5).half() for _ in range(3))
        o = torch.empty_like(q)
        for _ in range(warm):
            fa.flash_attn_fwd(q, k, v, causal=bool(causal), out=o)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fa.flash_attn_fwd(q, k, v, causal=bool(causal), out=o)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
    else:
        ms, _ = cp_time(workload, rank, world, local_rank, args.ring_exchange, steps, warm, barrier, max_over_ranks)
    if rank == 0:
        fl = flops(B, H, N, D, causal)
        pk = peaks()
        print(json.dumps({
            "metric": "fwd_tflops", "value": round(fl / (ms * 1e-3) / 1e12, 2), "unit": "TFLOPS", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": round(ms, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": args.workload, "B": B, "H": H, "N": N, "D": D, "causal": causal,
                       "parallelism": (f"cp{world} zig-zag, "
                                       + ("copy-engine pulls of the unmasked K/V chunks from the owner's HBM "
                                          "(flash_attn_peer_copy), no communication kernel"
                                          if args.ring_exchange == "pull" else
                                          "copy-engine pulls into one gathered K/V buffer, two kernels per rank, no partial states"
                                          if args.ring_exchange == "gather" else "NCCL send/recv of K/V chunk pairs")
                                       + f" (<= {4 * B * H * C * D * 2 / 2**20:.0f} MiB per hop per rank)") if world > 1
                       else "single GPU, monolithic kernel"},
            "roofline": {"bound": "tensor", "achieved": round(fl / (ms * 1e-3) / 1e12 / world, 2),
                         "peak": pk["tflops_sustained"] or pk["tflops"], "unit": "TFLOP/s",
                         "frac": round(fl / (ms * 1e-3) / 1e12 / world / (pk["tflops_sustained"] or pk["tflops"]), 4),
                         "traffic": None, "peak_source": pk["source"] + ", cuBLAS bf16 sustained (long step)"},
            "gpu_launches": None}))
    if world > 1:
        ring.release_peer_kv()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cpu"])
    ap.add_argument("--ring-exchange", default="gather", choices=["gather", "pull", "sendrecv"],
                    help="cfg5 only: how the ranks get at each other's K/V (flash_attention_cuda_b200/ring.py)")
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--sweep", action="store_true", help="also print the README-style TFLOPS table to stderr")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    ap.add_argument("--sustain-s", type=float, default=2.0,
                    help="seconds of back-to-back launches for the `sustained` sub-record (0 = skip)")
    ap.add_argument("--flush-l2", action="store_true",
                    help="write a 256 MiB buffer between timed launches (cold L2); automatic when the tensors fit L2")
    ap.add_argument("--hot-l2", action="store_true", help="never flush: the reference's method (FA.cu:942-960)")
    ap.add_argument("--no-multi", action="store_true", help="N > 1: skip the strong_cfg3 / cp_cfg5 / cp_parity legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = world if world > 1 else args.gpus
    workload = WORKLOADS[args.workload]
    B, H, N, D, causal = workload

    if args.impl == "reference-cpu":
        if rank == 0:
            cb = cpu_baseline(workload)
            print(json.dumps({
                "impl": "reference", "metric": "fwd_tflops", "value": cb["value"], "unit": "TFLOPS",
                "n_gpus": 0, "steps": 1, "warmup": 0, "ms_per_step": None, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "B": B, "H": H, "N": N, "D": D, "causal": causal},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "TFLOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return float(x)

    # ---- implementations: a launcher taking device pointers + a stream, like FA.cu:606-611 ----
    if args.impl == "ours":
        import flash_attention_cuda_b200 as fa
        L = fa.lib()

        def launch(q, k, v, o, b, h, n, d, c, stream_ptr):
            rc = L.flash_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), b, h, n, d, c, stream_ptr)
            if rc != 0:
                raise RuntimeError(L.flash_attn_error_string(rc).decode())
        launches_of = fa.launch_count
    else:
        ref_so = os.path.join(REPO, "oracle", "_ref", "libref_v9.so")
        if not os.path.exists(ref_so):
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable":
                                  "oracle/_ref/libref_v9.so missing (built from /root/reference by `make -C oracle`)"}))
            return
        R = ctypes.CDLL(ref_so)
        vp = ctypes.c_void_p
        R.ref_v9_dispatch.argtypes = [vp, vp, vp, vp] + [ctypes.c_int] * 5 + [vp]
        counter = [0]

        def launch(q, k, v, o, b, h, n, d, c, stream_ptr):
            if d != 128:
                raise RuntimeError("the reference dispatcher hard-codes head_dim 128 (FA.cu:613)")
            R.ref_v9_dispatch(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), b, h, n, d, c, stream_ptr)
            counter[0] += 1
        launches_of = lambda: counter[0]

    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)

    if args.workload.startswith("cfg5_ring"):
        return bench_ring(args, workload, rank, world, local_rank, barrier, max_over_ranks)
    if args.workload in STRONG_BH_SHARD and world > 1:
        # strong scaling: the B*H heads are split across ranks (no collective), total work fixed
        from flash_attention_cuda_b200.ring import bh_shard
        _, cnt = bh_shard(B * H, rank, world)
        B, H = 1, cnt

    def make_inputs(b, h, n, d, seed):
        g = torch.Generator(device="cuda").manual_seed(seed + 1000 * rank)
        # the reference's distribution: U(-0.5, 0.5) (FA.cu:766-768), generated on device for the big shapes
        q, k, v = ((torch.rand((b, h, n, d), device="cuda", generator=g) - 0.5).half() for _ in range(3))
        return q, k, v, torch.empty_like(q)

    flush_buf = [None]

    def time_kernel(b, h, n, d, c, steps, warmup, flush=False, bufs=None, active=True):
        """ms per launch (max over ranks).  flush: a 256 MiB write between launches evicts the tensors from L2; every
        launch is then timed by its own event pair so that the flush is outside the timed intervals."""
        if not active:           # this rank sits the leg out (single-GPU baselines inside a multi-GPU run)
            barrier(); barrier()
            return max_over_ranks(0.0), None
        q, k, v, o = bufs or make_inputs(b, h, n, d, n)
        for _ in range(warmup):
            launch(q, k, v, o, b, h, n, d, c, sp)
        barrier()
        if not flush:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                launch(q, k, v, o, b, h, n, d, c, sp)
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1) / steps
        else:
            if flush_buf[0] is None:
                flush_buf[0] = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a, z in ev:
                flush_buf[0].fill_(1)
                a.record(stream)
                launch(q, k, v, o, b, h, n, d, c, sp)
                z.record(stream)
            barrier()
            ms = sum(a.elapsed_time(z) for a, z in ev) / steps
        return max_over_ranks(ms), (q, k, v, o)

    # ---- main timed region, with clocks sampled while it runs ----
    nbytes = B * H * N * D * 2
    fits_l2 = 4 * nbytes <= L2_BYTES
    flush = (args.flush_l2 or fits_l2) and not args.hot_l2
    sampler = ClockSampler(local_rank).start()
    l0 = launches_of()
    ms, bufs = time_kernel(B, H, N, D, causal, args.steps, args.warmup, flush=flush)
    clocks = sampler.stop()
    launches = launches_of() - l0 - args.warmup
    fl = flops(B, H, N, D, causal)
    tflops_rank = fl / (ms * 1e-3) / 1e12
    value = tflops_rank * n_gpus if world > 1 else tflops_rank
    hot_ms = None
    if flush:      # the reference's method (back-to-back launches, tensors L2-resident) next to the cold-L2 number
        hot_ms, _ = time_kernel(B, H, N, D, causal, args.steps, 3, flush=False, bufs=bufs)

    # ---- e2e: host buffers through the public call, H2D + kernel + D2H timed every step ----
    e2e_steps = args.e2e_steps or min(args.steps, 20)
    q, k, v, o = bufs
    hq, hk, hv = (torch.empty(q.shape, dtype=torch.float16).pin_memory() for _ in range(3))
    ho = torch.empty(q.shape, dtype=torch.float16).pin_memory()
    hq.copy_(q.cpu()); hk.copy_(k.cpu()); hv.copy_(v.cpu())
    if args.impl == "ours":
        def e2e_step():
            rc = L.flash_attn_fwd_host(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), ho.data_ptr(), B, H, N, D, causal)
            if rc != 0:
                raise RuntimeError(L.flash_attn_error_string(rc).decode())
    else:
        def e2e_step():   # the reference harness's own sequence, FA.cu:774-780
            q.copy_(hq, non_blocking=True); k.copy_(hk, non_blocking=True); v.copy_(hv, non_blocking=True)
            launch(q, k, v, o, B, H, N, D, causal, sp)
            ho.copy_(o, non_blocking=True)
            stream.synchronize()

    def wall_ms(step, n):
        for _ in range(2):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            step()
        barrier()
        return max_over_ranks((time.perf_counter() - t0) * 1e3 / n)

    e2e_ms = wall_ms(e2e_step, e2e_steps)
    e2e_val = fl / (e2e_ms * 1e-3) / 1e12 * (n_gpus if world > 1 else 1)
    # the floor of that call: the same bytes over PCIe (H2D of Q, K, V on one stream, D2H of O on another -- full duplex),
    # no kernel, every rank at once.  With several ranks the host's memory system is shared: this is what e2e can reach.
    side = torch.cuda.Stream()

    def copy_only():
        q.copy_(hq, non_blocking=True); k.copy_(hk, non_blocking=True); v.copy_(hv, non_blocking=True)
        with torch.cuda.stream(side):
            ho.copy_(o, non_blocking=True)
        stream.synchronize()
        side.synchronize()
    copy_ms = wall_ms(copy_only, max(3, e2e_steps // 2))

    # ---- sustained leg: seconds of back-to-back launches, its own clocks record.  Runs AFTER the e2e leg: it leaves the
    # part power-capped and hot for a while, and the legs are separate measurements (both arms run the same order) ----
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(args.steps, int(args.sustain_s * 1e3 / ms) + 1)
        s2 = ClockSampler(local_rank).start()
        sus_ms, _ = time_kernel(B, H, N, D, causal, n_sus, 3, flush=False, bufs=bufs)
        c2 = s2.stop()
        pk0 = peaks()
        sus_tf = fl / (sus_ms * 1e-3) / 1e12
        sustained = {"value": round(sus_tf * (n_gpus if world > 1 else 1), 2), "unit": "TFLOPS", "steps": n_sus,
                     "ms_per_step": round(sus_ms, 5), "seconds": round(n_sus * sus_ms * 1e-3, 2), "clocks": c2,
                     "roofline": {"bound": "tensor", "achieved": round(sus_tf, 2),
                                  "peak": pk0["tflops_sustained"] or None, "unit": "TFLOP/s",
                                  "frac": round(sus_tf / pk0["tflops_sustained"], 4) if pk0["tflops_sustained"] else None,
                                  "peak_source": pk0["source"] + ", cuBLAS bf16 sustained (4 s back to back)"}}

    # ---- optional README-style sweep (stderr, rank 0) ----
    sweep = None
    if args.sweep and world == 1:
        sweep = {}
        for c in (0, 1):
            for n in (512, 768, 1024, 2048, 4096, 8192, 16384):
                cold = 4 * 32 * n * 128 * 2 <= L2_BYTES
                m, _ = time_kernel(1, 32, n, 128, c, 50, 5, flush=cold and not args.hot_l2)
                sweep[f"{'causal' if c else 'full'}_{n}"] = round(flops(1, 32, n, 128, c) / (m * 1e-3) / 1e12, 1)
        print("sweep TFLOPS (B1 H32 D128; cold L2 where the tensors fit L2):", json.dumps(sweep), file=sys.stderr)

    # ---- N > 1: the two multi-GPU configurations BASELINE.json names, in the same process group ----
    multi = {}
    parity_ok = True
    if world > 1 and args.impl == "ours" and not args.no_multi and args.workload == DEFAULT_WORKLOAD:
        from flash_attention_cuda_b200 import ring
        del q, k, v, o, hq, hk, hv, ho, bufs
        torch.cuda.empty_cache()
        pk0 = peaks()
        # -- config 3: B16 H32 N8192 causal, heads split over the ranks, no collective
        B3, H3, N3, D3, c3 = WORKLOADS["cfg3_b16_n8192_causal"]
        _, cnt = ring.bh_shard(B3 * H3, rank, world)
        fl3 = flops(B3, H3, N3, D3, c3)
        st3 = max(5, min(args.steps, 30))
        ms3, b3 = time_kernel(1, cnt, N3, D3, c3, st3, 3)
        # the same slice on ONE GPU with the others idle: same launch length, same power regime
        ms3_slice, _ = time_kernel(1, cnt, N3, D3, c3, st3, 3, bufs=b3, active=(rank == 0))
        del b3
        torch.cuda.empty_cache()
        # and the whole batch on one GPU (a launch `world` times longer: power-capped clocks)
        ms3_full, b3f = time_kernel(1, B3 * H3, N3, D3, c3, max(3, st3 // 4), 2, active=(rank == 0))
        del b3f
        torch.cuda.empty_cache()
        multi["strong_cfg3"] = {
            "workload": "cfg3_b16_n8192_causal", "heads_per_rank": cnt, "steps": st3, "ms": round(ms3, 4),
            "tflops": round(fl3 / (ms3 * 1e-3) / 1e12, 1),
            "one_gpu_same_slice_ms": round(ms3_slice, 4), "efficiency": round(ms3_slice / ms3, 4),
            "one_gpu_whole_batch_ms": round(ms3_full, 4),
            "efficiency_vs_whole_batch_on_one_gpu": round(ms3_full / (world * ms3), 4),
            "note": "efficiency = time of one rank's slice alone on one GPU / time with all ranks running theirs "
                    "(same launch length on both sides); the whole-batch figure is a launch N times longer and runs "
                    "at power-capped clocks"}
        # -- config 5: B1 H32 N131072 causal: heads split (no communication) vs context parallel
        w5 = WORKLOADS["cfg5_ring_n131072_causal"]
        B5, H5, N5, D5, c5 = w5
        fl5 = flops(*w5)
        _, cnt5 = ring.bh_shard(B5 * H5, rank, world)
        ms5_heads, b5 = time_kernel(1, cnt5, N5, D5, c5, 3, 2)
        del b5
        torch.cuda.empty_cache()
        cp = {"workload": "cfg5_ring_n131072_causal", "chunk_rows": N5 // (2 * world),
              "heads_split_bound": {"ms": round(ms5_heads, 3), "tflops": round(fl5 / (ms5_heads * 1e-3) / 1e12, 1),
                                    "heads_per_rank": cnt5}}
        outs = {}
        for ex in ("gather", "pull", "sendrecv"):
            ms5, out = cp_time(w5, rank, world, local_rank, ex, 4, 3, barrier, max_over_ranks)
            outs[ex] = [t.clone() for t in out]
            del out
            if ex != "sendrecv":
                ring.release_peer_kv()         # one peer-readable K/V set at a time
                torch.cuda.empty_cache()
            cp[ex] = {"ms": round(ms5, 3), "tflops": round(fl5 / (ms5 * 1e-3) / 1e12, 1),
                      "efficiency": round(ms5_heads / ms5, 4),
                      "frac_of_sustained_peak_per_gpu": round(fl5 / (ms5 * 1e-3) / 1e12 / world / pk0["tflops_sustained"], 4)
                      if pk0["tflops_sustained"] else None}
        cp["note"] = ("efficiency = heads-split time (same shape, no communication, measured in this run) / context-parallel "
                      "time; gather = copy-engine pulls into one gathered K/V buffer per head, two kernels per rank, no partial "
                      "states; pull = the same pulls, one kernel and one partial state per chunk pair, merged at the end; "
                      "sendrecv = NCCL ring with those kernels")
        multi["cp_cfg5"] = cp
        par = cp_parity(w5, rank, world, outs)
        if rank == 0:
            multi["cp_parity"] = par
            parity_ok = bool(par["pass"])
        del outs
        ring.release_peer_kv()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    traffic, traffic_src = ncu_traffic(args.workload) if args.impl == "ours" else (None, None)
    out = {
        "metric": "fwd_tflops", "value": round(value, 2), "unit": "TFLOPS", "n_gpus": n_gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 5), "higher_is_better": True,
        "scaling": "strong" if (args.workload in STRONG_BH_SHARD and world > 1) else "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": args.workload, "B": B, "H": H, "N": N, "D": D, "causal": causal,
                   "per_gpu": "every rank runs the full workload on its own GPU (batch x heads shard, no collective)"
                              if not (args.workload in STRONG_BH_SHARD and world > 1) else
                              "the workload's B*H heads are split across the ranks (no collective); B,H here are rank 0's share",
                   "l2": (f"inputs {4 * nbytes / 2**20:.0f} MiB per step > 126 MB L2 (no flush needed)" if not fits_l2 else
                          "inputs fit L2: a 256 MiB write between launches evicts them (cold L2), every launch timed by its "
                          "own event pair; `hot_l2_ms_per_step` is the reference's method (FA.cu:942-960)" if flush else
                          "inputs fit L2: hot-L2 timing, reference method (FA.cu:942-960)"),
                   "flops_convention": "4*B*H*N^2*D, /2 causal (FA.cu:938-939)"},
        "clocks": clocks,
        "e2e": {"value": round(e2e_val, 3), "unit": "TFLOPS", "h2d_bytes_per_step": 3 * nbytes,
                "d2h_bytes_per_step": nbytes, "ms_per_step": round(e2e_ms, 4), "steps": e2e_steps,
                "copy_only_ms": round(copy_ms, 4),
                "copy_only_gbs_per_rank": round(4 * nbytes / (copy_ms * 1e-3) / 1e9, 1),
                "frac_of_copy_floor": round(copy_ms / e2e_ms, 4),
                "call": "flash_attn_fwd_host (pinned host Q,K,V -> device by copy engine, kernel, O tiles TMA-stored by the "
                        "kernel's epilogue straight into the pinned host buffer; 8 head chunks pipelined over two streams, "
                        "every byte still crosses PCIe inside the timed region)"
                        if args.impl == "ours" else "H2D x3 + flash_attention_v9_dispatch + D2H (FA.cu:774-780)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": round(tflops_rank, 2), "peak": pk["tflops"], "unit": "TFLOP/s",
                     "frac": round(tflops_rank / pk["tflops"], 4), "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": pk["source"] + ", cuBLAS bf16 burst (timed region of milliseconds)",
                     "frac_of_nominal_2250": round(tflops_rank / NOMINAL_FP16_TFLOPS, 4),
                     "algorithmic_flops_per_launch": fl, "algorithmic_bytes_per_launch": 4 * nbytes,
                     "hbm_gbs_algorithmic": round(4 * nbytes / (ms * 1e-3) / 1e9, 1)},
    }
    if hot_ms is not None:
        out["hot_l2_ms_per_step"] = round(hot_ms, 5)
        out["hot_l2_tflops"] = round(fl / (hot_ms * 1e-3) / 1e12, 2)
    if sustained:
        out["sustained"] = sustained
    if sweep:
        out["sweep_tflops"] = sweep
    out.update(multi)
    if args.impl != "ours":
        out["impl"] = "reference"
        out["reference_kind"] = "V9 kernel (flash_attention_v9_dispatch) rebuilt for sm_100a, unmodified source"
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(workload)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    if not parity_ok:
        raise SystemExit("cp_parity FAILED: context-parallel output differs from the CPU oracle beyond 2e-3 / 2e-4")


if __name__ == "__main__":
    main()

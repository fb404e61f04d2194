"""flash_attention_cuda_b200 -- host-side binding of libflashattn_b200.so (B200 / sm_100a).

The product is the C-ABI shared library (include/flash_attn.h); this module is the thin
ctypes layer the tests and bench.py use to call it with torch tensors.  It mirrors the
reference's launcher surface, ``flash_attention_v9_dispatch`` (reference
flash_attention.cu:606-663): FP16 ``[B, H, N, D]`` tensors, scale ``1/sqrt(D)``, optional
causal mask, launch on the current stream.

There is no fallback: if the CUDA library is missing or fails to load, importing the symbols
raises.  Nothing in here touches ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_PKG_DIR)
# FLASH_ATTN_B200_LIB selects another build of the same library (A/B runs of kernel variants)
LIB_PATH = os.environ.get("FLASH_ATTN_B200_LIB") or os.path.join(_PKG_DIR, "libflashattn_b200.so")

FA_OK = 0
ERRORS = {
    -1: "FA_ERR_BAD_HEAD_DIM",
    -2: "FA_ERR_NULL_PTR",
    -3: "FA_ERR_MISALIGNED",
    -4: "FA_ERR_BAD_SHAPE",
    -5: "FA_ERR_UNSUPPORTED_ARCH",
    -6: "FA_ERR_TENSORMAP",
    -7: "FA_ERR_WORKSPACE",
    -8: "FA_ERR_WATCHDOG",
}

# every symbol include/flash_attn.h declares
EXPORTED_SYMBOLS = (
    "flash_attn_fwd",
    "flash_attn_fwd_bf16",
    "flash_attn_fwd_ex",
    "flash_attn_fwd_gathered",
    "flash_attn_stream_write_flag",
    "flash_attn_finalize",
    "flash_attn_merge",
    "flash_attn_fwd_host",
    "flash_attn_get_kernel_info",
    "flash_attn_set_sm_margin",
    "flash_attn_peer_alloc",
    "flash_attn_peer_open",
    "flash_attn_peer_close",
    "flash_attn_peer_free",
    "flash_attn_peer_copy",
    "flash_attn_peer_copy_2d",
    "flash_attn_status",
    "flash_attn_launch_count",
    "flash_attn_destroy",
    "flash_attn_error_string",
    "flash_attn_version",
)


class FlashAttnError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"flash_attn: {msg} (code {code})")
        self.code = code


class KernelInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in (
        "regs_per_thread", "local_bytes_per_thread", "static_smem_bytes", "dynamic_smem_bytes",
        "threads_per_cta", "ctas", "tmem_columns", "kv_stages", "work_items", "num_sms", "cta_group")]


def build(force: bool = False) -> str:
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(_PKG_DIR, "csrc", f) for f in ("fa_api.cu", "fa_fwd_sm100.cuh", "sm100_ptx.cuh")]
    srcs.append(os.path.join(_REPO, "include", "flash_attn.h"))
    if not force and os.path.exists(LIB_PATH):
        if all(os.path.getmtime(s) <= os.path.getmtime(LIB_PATH) for s in srcs if os.path.exists(s)):
            return LIB_PATH
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", srcs[0], "-o", LIB_PATH]
    subprocess.run(cmd, check=True, cwd=_REPO)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    """Load libflashattn_b200.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `make` or flash_attention_cuda_b200.build(); "
            "there is no CPU or PyTorch fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
    L.flash_attn_fwd.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]
    L.flash_attn_fwd.restype = ci
    if hasattr(L, "flash_attn_fwd_bf16"):   # absent from archived A/B builds of older kernels
        L.flash_attn_fwd_bf16.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]
        L.flash_attn_fwd_bf16.restype = ci
    L.flash_attn_fwd_ex.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ll, ll, ci, vp]
    L.flash_attn_fwd_ex.restype = ci
    if hasattr(L, "flash_attn_fwd_gathered"):   # absent from archived A/B builds of older kernels
        L.flash_attn_fwd_gathered.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ll, ll, vp, ci, vp]
        L.flash_attn_fwd_gathered.restype = ci
        L.flash_attn_stream_write_flag.argtypes = [vp, ci, vp]
        L.flash_attn_stream_write_flag.restype = ci
        L.flash_attn_peer_copy_2d.argtypes = [vp, ctypes.c_size_t, vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, vp]
        L.flash_attn_peer_copy_2d.restype = ci
    L.flash_attn_finalize.argtypes = [vp, vp, vp, ll, ci, vp]
    L.flash_attn_finalize.restype = ci
    if hasattr(L, "flash_attn_merge"):   # absent from archived A/B builds of older kernels
        L.flash_attn_merge.argtypes = [vp, vp, vp, ci, ll, ci, vp]
        L.flash_attn_merge.restype = ci
    L.flash_attn_fwd_host.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci]
    L.flash_attn_fwd_host.restype = ci
    L.flash_attn_get_kernel_info.argtypes = [ci, ci, ci, ci, ci, ctypes.POINTER(KernelInfo)]
    L.flash_attn_get_kernel_info.restype = ci
    if hasattr(L, "flash_attn_set_sm_margin"):   # absent from archived A/B builds of older kernels
        L.flash_attn_set_sm_margin.argtypes = [ci]
        L.flash_attn_set_sm_margin.restype = ci
    if hasattr(L, "flash_attn_peer_alloc"):   # absent from archived A/B builds of older kernels
        L.flash_attn_peer_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp), ctypes.c_char_p]
        L.flash_attn_peer_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
        L.flash_attn_peer_close.argtypes = [vp]
        L.flash_attn_peer_free.argtypes = [vp]
        L.flash_attn_peer_copy.argtypes = [vp, vp, ctypes.c_size_t, vp]
        for f in (L.flash_attn_peer_alloc, L.flash_attn_peer_open, L.flash_attn_peer_close,
                  L.flash_attn_peer_free, L.flash_attn_peer_copy):
            f.restype = ci
    L.flash_attn_launch_count.argtypes = []
    L.flash_attn_launch_count.restype = ctypes.c_ulonglong
    L.flash_attn_destroy.argtypes = []
    L.flash_attn_destroy.restype = None
    L.flash_attn_error_string.argtypes = [ci]
    L.flash_attn_error_string.restype = ctypes.c_char_p
    L.flash_attn_version.argtypes = []
    L.flash_attn_version.restype = ctypes.c_char_p
    L.flash_attn_debug_work_item.argtypes = [ci, ci, ci, ci, ci, ci, ci, ll] + [ctypes.POINTER(ci)] * 5
    L.flash_attn_debug_work_item.restype = ci
    if hasattr(L, "flash_attn_debug_work_item_split"):   # absent from archived A/B builds of older kernels
        L.flash_attn_debug_work_item_split.argtypes = [ci, ci, ci, ci, ci, ci, ci, ll] + [ctypes.POINTER(ci)] * 5
        L.flash_attn_debug_work_item_split.restype = ci
        L.flash_attn_debug_set_split.argtypes = [ci]
        L.flash_attn_debug_set_split.restype = None
        L.flash_attn_debug_uses_split.argtypes = [ci, ci, ci, ci]
        L.flash_attn_debug_uses_split.restype = ci
    if hasattr(L, "flash_attn_debug_tiles_per_item"):   # absent from archived A/B builds of older kernels
        L.flash_attn_debug_tiles_per_item.argtypes = [ci]
        L.flash_attn_debug_tiles_per_item.restype = ci
    L.flash_attn_debug_status.argtypes = [ctypes.POINTER(ctypes.c_uint)]
    L.flash_attn_debug_status.restype = ci
    if hasattr(L, "flash_attn_status"):   # absent from archived A/B builds of older kernels
        L.flash_attn_status.argtypes = [ctypes.POINTER(ctypes.c_uint)]
        L.flash_attn_status.restype = ci
        L.flash_attn_debug_trip_watchdog.argtypes = [ctypes.c_uint, vp]
        L.flash_attn_debug_trip_watchdog.restype = ci
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != FA_OK:
        raise FlashAttnError(rc, lib().flash_attn_error_string(rc).decode())


def _stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


def _check_qkv(q, k, v, dtypes, what):
    """The data contract of the C ABI (include/flash_attn.h): CUDA, one 16-bit type, contiguous [B, H, N, D], one device."""
    if not (q.is_cuda and k.is_cuda and v.is_cuda):
        raise ValueError(f"{what} needs CUDA tensors (there is no CPU path)")
    if q.dtype not in dtypes or k.dtype != q.dtype or v.dtype != q.dtype:
        raise TypeError(f"{what} takes {' or '.join(str(d).replace('torch.', '') for d in dtypes)} tensors, all of one type")
    if q.dim() != 4 or k.dim() != 4 or v.dim() != 4 or k.shape != v.shape:
        raise ValueError("q must be [B, H, Nq, D] and k, v [B, H, Nkv, D]")
    if q.shape[:2] != k.shape[:2] or q.shape[3] != k.shape[3]:
        raise ValueError("q, k, v must agree in B, H and D")
    if not (q.is_contiguous() and k.is_contiguous() and v.is_contiguous()):
        raise ValueError("q, k, v must be contiguous [B, H, N, D] (a slice along N is not)")
    if k.device != q.device or v.device != q.device:
        raise ValueError("q, k, v must live on one device")


def _check_f32(t, shape, name, device):
    import torch
    if not t.is_cuda or t.device != device:
        raise ValueError(f"{name} must be a CUDA tensor on {device}")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError(f"{name} must be contiguous float32")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, not {tuple(t.shape)}")


def flash_attn_fwd(q, k, v, causal: bool = True, out=None, stream=None):
    """O = softmax(Q K^T / sqrt(D) [+ causal mask]) V for FP16 (or BF16) CUDA tensors [B, H, N, D].

    Same contract as the reference dispatcher (flash_attention.cu:606-663); enqueues on the
    current (or given) torch stream and returns the output tensor without synchronising."""
    import torch
    _check_qkv(q, k, v, (torch.float16, torch.bfloat16), "flash_attn_fwd")
    if q.shape != k.shape:
        raise ValueError("q, k, v must all be [B, H, N, D]")
    B, H, N, D = q.shape
    if out is None:
        out = torch.empty_like(q)
    if out.dtype != q.dtype:
        raise TypeError("out must have the dtype of q, k, v")
    if out.shape != q.shape or out.device != q.device or not out.is_contiguous():
        raise ValueError("out must be a contiguous [B, H, N, D] tensor on the device of q")
    entry = lib().flash_attn_fwd if q.dtype == torch.float16 else lib().flash_attn_fwd_bf16
    with torch.cuda.device(q.device):
        rc = entry(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(),
                   B, H, N, D, 1 if causal else 0, _stream_ptr(stream))
    check(rc)
    return out


def flash_attn_fwd_partial(q, k, v, o_partial, ml, causal: bool, q_offset: int, kv_offset: int,
                           accumulate: bool, stream=None):
    """One K/V block of a longer sequence -> (o_partial fp32, ml) partial state (include/flash_attn.h).
    FP16 only: the partial / merge path has no BF16 instantiation."""
    import torch
    _check_qkv(q, k, v, (torch.float16,), "flash_attn_fwd_partial")
    B, H, Nq, D = q.shape
    Nkv = k.shape[2]
    if o_partial.numel() != B * H * Nq * D or ml.numel() != B * H * Nq * 2:
        raise ValueError("o_partial must hold B*H*Nq*D and ml B*H*Nq*2 float32 values")
    _check_f32(o_partial, o_partial.shape, "o_partial", q.device)
    _check_f32(ml, ml.shape, "ml", q.device)
    with torch.cuda.device(q.device):
        rc = lib().flash_attn_fwd_ex(q.data_ptr(), k.data_ptr(), v.data_ptr(), o_partial.data_ptr(),
                                     ml.data_ptr(), B, H, Nq, Nkv, D, 1 if causal else 0,
                                     q_offset, kv_offset, 1 if accumulate else 0, _stream_ptr(stream))
    check(rc)


def flash_attn_fwd_gathered(q, k_ptr: int, v_ptr: int, out, Nkv: int, kv_head_rows: int, causal: bool, q_offset: int,
                            ready=None, ready_rows: int = 0, stream=None):
    """Attention of q [B, H, Nq, D] (FP16, at global positions q_offset..) against a GATHERED K/V buffer: raw device pointers
    to B*H heads of `kv_head_rows` rows each, the first Nkv in use (include/flash_attn.h).  `ready`: optional int32 CUDA
    tensor of per-chunk flags (one per `ready_rows` rows) the kernel waits on while copies are still landing."""
    import torch
    if not q.is_cuda or q.dtype != torch.float16 or q.dim() != 4 or not q.is_contiguous():
        raise TypeError("q must be a contiguous float16 CUDA tensor [B, H, Nq, D]")
    if out.shape != q.shape or out.dtype != q.dtype or out.device != q.device or not out.is_contiguous():
        raise ValueError("out must match q")
    if ready is not None and (ready.dtype != torch.int32 or ready.device != q.device or not ready.is_contiguous()):
        raise TypeError("ready must be a contiguous int32 tensor on the device of q")
    B, H, Nq, D = q.shape
    with torch.cuda.device(q.device):
        rc = lib().flash_attn_fwd_gathered(q.data_ptr(), k_ptr, v_ptr, out.data_ptr(), B, H, Nq, Nkv, D, 1 if causal else 0,
                                           q_offset, kv_head_rows, ready.data_ptr() if ready is not None else None,
                                           ready_rows, _stream_ptr(stream))
    check(rc)
    return out


def _check_f16_out(out, rows, D, name="out"):
    import torch
    if not out.is_cuda or out.dtype != torch.float16 or not out.is_contiguous():
        raise TypeError(f"{name} must be a contiguous float16 CUDA tensor (the merge writes FP16)")
    if out.numel() != rows * D:
        raise ValueError(f"{name} must hold rows * D = {rows * D} elements, not {out.numel()}")


def flash_attn_finalize(o_partial, ml, out, stream=None):
    import torch
    D = o_partial.shape[-1]
    rows = o_partial.numel() // D
    _check_f32(o_partial, o_partial.shape, "o_partial", out.device)
    _check_f32(ml, ml.shape, "ml", out.device)
    if ml.numel() != rows * 2:
        raise ValueError("ml must hold rows * 2 float32 values")
    _check_f16_out(out, rows, D)
    with torch.cuda.device(out.device):
        rc = lib().flash_attn_finalize(o_partial.data_ptr(), ml.data_ptr(), out.data_ptr(), rows, D, _stream_ptr(stream))
    check(rc)
    return out


def flash_attn_merge(o_partials, mls, out, stream=None):
    """Merge `splits` partial states (o_partials [S, rows, D] fp32, mls [S, rows, 2]) into the FP16 tensor `out`."""
    import torch
    if o_partials.dim() != 3:
        raise ValueError("o_partials must be [splits, rows, D]")
    S, rows, D = o_partials.shape
    _check_f32(o_partials, (S, rows, D), "o_partials", out.device)
    _check_f32(mls, (S, rows, 2), "mls", out.device)
    _check_f16_out(out, rows, D)
    with torch.cuda.device(out.device):
        rc = lib().flash_attn_merge(o_partials.data_ptr(), mls.data_ptr(), out.data_ptr(), S, rows, D,
                                    _stream_ptr(stream))
    check(rc)
    return out


def kernel_info(B: int, H: int, N: int, D: int, causal: bool) -> dict:
    info = KernelInfo()
    check(lib().flash_attn_get_kernel_info(B, H, N, D, 1 if causal else 0, ctypes.byref(info)))
    return {n: getattr(info, n) for n, _ in KernelInfo._fields_}


def watchdog_status(sync: bool = True) -> dict:
    """The kernel watchdog record of the current device (include/flash_attn.h: flash_attn_status), by default after
    a device synchronise.  aborted == 1: a kernel gave up on a barrier and the next call will raise FA_ERR_WATCHDOG."""
    buf = (ctypes.c_uint * 4)()
    L = lib()
    check((L.flash_attn_debug_status if sync or not hasattr(L, "flash_attn_status") else L.flash_attn_status)(buf))
    return {"aborted": int(buf[0]), "tag": int(buf[1]), "block": int(buf[2]), "thread": int(buf[3])}


def set_sm_margin(sms: int) -> int:
    """Leave `sms` SMs free in every later launch (room for NCCL kernels); returns the previous value."""
    return int(lib().flash_attn_set_sm_margin(int(sms)))


PEER_HANDLE_BYTES = 64


class _DevicePtr:
    """Raw device memory seen through __cuda_array_interface__ (lets torch alias it without owning it)."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = ptr, nbytes
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _as_tensor(ptr: int, nbytes: int, device):
    import torch
    return torch.as_tensor(_DevicePtr(ptr, nbytes), device=device)


def peer_alloc(nbytes: int, device=None):
    """(uint8 CUDA tensor over a peer-readable block on the current device, 64-byte handle, raw pointer)."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ptr = ctypes.c_void_p()
    handle = ctypes.create_string_buffer(PEER_HANDLE_BYTES)
    with torch.cuda.device(dev):
        check(lib().flash_attn_peer_alloc(nbytes, ctypes.byref(ptr), handle))
    return _as_tensor(ptr.value, nbytes, dev), handle.raw, ptr.value


def peer_open(handle: bytes) -> int:
    """Map another process's block (same node) into this process; returns the device pointer."""
    ptr = ctypes.c_void_p()
    check(lib().flash_attn_peer_open(handle, ctypes.byref(ptr)))
    return ptr.value


def peer_close(ptr: int) -> None:
    check(lib().flash_attn_peer_close(ptr))


def peer_free(ptr: int) -> None:
    check(lib().flash_attn_peer_free(ptr))


def peer_copy(dst_ptr: int, src_ptr: int, nbytes: int, stream=None) -> None:
    """Enqueue a copy-engine transfer (any mix of local and mapped peer pointers) on a torch stream."""
    check(lib().flash_attn_peer_copy(dst_ptr, src_ptr, nbytes, _stream_ptr(stream)))


def peer_copy_2d(dst_ptr: int, dpitch: int, src_ptr: int, spitch: int, width: int, height: int, stream=None) -> None:
    """`height` rows of `width` bytes, pitches in bytes (copy engine; local and mapped peer pointers alike)."""
    check(lib().flash_attn_peer_copy_2d(dst_ptr, dpitch, src_ptr, spitch, width, height, _stream_ptr(stream)))


def stream_write_flag(flag_ptr: int, value: int, stream=None) -> None:
    """*flag = value, ordered behind the work already queued on the stream; a stream memory operation, no kernel."""
    check(lib().flash_attn_stream_write_flag(flag_ptr, value, _stream_ptr(stream)))


def launch_count() -> int:
    return int(lib().flash_attn_launch_count())


def work_item(w: int, B: int, H: int, Nq: int, Nkv: int, D: int, causal: bool, shift: int = 0, split: bool = False):
    """Host mirror of the device work decomposition (scheduler tests).  split: the short-sequence mode in which an item
    is one Q tile and n0 / n1 are its even / odd KV tiles (fa_fwd_sm100.cuh)."""
    vals = [ctypes.c_int() for _ in range(5)]
    fn = lib().flash_attn_debug_work_item_split if split else lib().flash_attn_debug_work_item
    rc = fn(w, B, H, Nq, Nkv, D, 1 if causal else 0, shift, *[ctypes.byref(x) for x in vals])
    check(rc)
    total, bh, q0, n0, n1 = (x.value for x in vals)
    return {"total": total, "bh": bh, "q0": q0, "n0": n0, "n1": n1, "n": n0 + n1 if split else max(n0, n1)}


def set_split(mode) -> None:
    """Work decomposition of flash_attn_fwd from now on: None = automatic, False = pair items, True = split mode."""
    lib().flash_attn_debug_set_split(-1 if mode is None else (1 if mode else 0))


def uses_split(B: int, H: int, N: int, causal: bool) -> bool:
    return bool(lib().flash_attn_debug_uses_split(B, H, N, 1 if causal else 0))


def tiles_per_item(D: int) -> int:
    """128-row Q tiles one pair-mode work item covers."""
    return int(lib().flash_attn_debug_tiles_per_item(D))

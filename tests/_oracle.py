"""ctypes binding of the checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

liboracle.so     -- C restatement of the reference's cpu_attention (oracle/attn_oracle.c)
_ref/libref_v9.so -- the reference's own translation unit compiled from /root/reference
                     (cpu_attention + flash_attention_v9_dispatch), when it has been built.
"""
import ctypes
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref_v9.so")

_vp = ctypes.c_void_p


def build():
    src = os.path.join(ORACLE_DIR, "attn_oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(src) > os.path.getmtime(ORACLE_SO):
        subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, capture_output=True)
    if os.path.exists("/root/reference/flash_attention.cu") and not os.path.exists(REF_SO):
        subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=True, capture_output=True)


_o = None
_r = None


def oracle():
    global _o
    if _o is None:
        build()
        _o = ctypes.CDLL(ORACLE_SO)
        _o.fa_oracle_diff.restype = ctypes.c_double
        _o.fa_oracle_max_threads.restype = ctypes.c_int
    return _o


def ref():
    """The reference's own compiled TU, or None when oracle/_ref was not built."""
    global _r
    if _r is None and os.path.exists(REF_SO):
        _r = ctypes.CDLL(REF_SO)
    return _r


def _p(a):
    return a.ctypes.data_as(_vp)


def _u16(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float16:
        a = a.view(np.uint16)
    assert a.dtype == np.uint16
    return a


def attention(q, k, v, causal, threads=0):
    """q,k,v: [B,H,N,D] float16 (or uint16 bit patterns). Returns float16 [B,H,N,D]."""
    q, k, v = _u16(q), _u16(k), _u16(v)
    B, H, N, D = q.shape
    out = np.empty_like(q)
    oracle().fa_oracle_attention(_p(q), _p(k), _p(v), _p(out), B, H, N, D, int(bool(causal)), threads)
    return out.view(np.float16)


def attention_rows(q, k, v, causal, bhs, rows, threads=0):
    """Row-sampled oracle: returns float16 [len(rows), D] for (bhs[i], rows[i])."""
    q, k, v = _u16(q), _u16(k), _u16(v)
    B, H, N, D = q.shape
    bhs = np.ascontiguousarray(bhs, dtype=np.int32)
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    out = np.empty((len(rows), D), np.uint16)
    oracle().fa_oracle_rows(_p(q), _p(k), _p(v), _p(out), N, D, int(bool(causal)), _p(bhs), _p(rows),
                            len(rows), threads)
    return out.view(np.float16)


def ref_attention(q, k, v, causal):
    """The reference's own cpu_attention (single thread), via oracle/_ref."""
    r = ref()
    assert r is not None, "oracle/_ref/libref_v9.so not built"
    q, k, v = _u16(q), _u16(k), _u16(v)
    B, H, N, D = q.shape
    out = np.empty_like(q)
    r.ref_cpu_attention(_p(q), _p(k), _p(v), _p(out), B, H, N, D, int(bool(causal)))
    return out.view(np.float16)


def fill_ref_rand(shape, seed=42):
    """The reference harness's input generator (interleaved glibc rand(), FA.cu:764-769)."""
    n = int(np.prod(shape))
    q = np.empty(n, np.uint16)
    k = np.empty(n, np.uint16)
    v = np.empty(n, np.uint16)
    oracle().fa_oracle_fill_ref_rand(_p(q), _p(k), _p(v), ctypes.c_size_t(n), ctypes.c_uint(seed))
    return (q.view(np.float16).reshape(shape), k.view(np.float16).reshape(shape),
            v.view(np.float16).reshape(shape))


def diff(a, b):
    """(max_abs, mean_abs) of two float16 arrays, computed like the reference check (FA.cu:781-783)."""
    a, b = _u16(a).ravel(), _u16(b).ravel()
    assert a.size == b.size
    mean = ctypes.c_double()
    mx = oracle().fa_oracle_diff(_p(a), _p(b), ctypes.c_size_t(a.size), ctypes.byref(mean))
    return float(mx), float(mean.value)


def checksum(a):
    a = _u16(a).ravel()
    s, sa = ctypes.c_double(), ctypes.c_double()
    oracle().fa_oracle_checksum(_p(a), ctypes.c_size_t(a.size), ctypes.byref(s), ctypes.byref(sa))
    return float(s.value), float(sa.value)


def merge_partials(o_part, ml_part):
    """o_part [S, rows, D] fp32, ml_part [S, rows, 2] fp32 -> float16 [rows, D] (FA.cu:575-597)."""
    o_part = np.ascontiguousarray(o_part, dtype=np.float32)
    ml_part = np.ascontiguousarray(ml_part, dtype=np.float32)
    S, rows, D = o_part.shape
    out = np.empty((rows, D), np.uint16)
    oracle().fa_oracle_merge_partials(_p(o_part), _p(ml_part), S, ctypes.c_long(rows), D, _p(out))
    return out.view(np.float16)


def max_threads():
    return int(oracle().fa_oracle_max_threads())


# tolerance of the north-star gate
MAX_ABS_TOL = 2e-3
MEAN_ABS_TOL = 2e-4

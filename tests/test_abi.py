"""CPU-only: the C-ABI library loads, exports every symbol include/flash_attn.h declares, and
validates its arguments before touching the device (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

import flash_attention_cuda_b200 as fa

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    fa.build()
    return fa.lib()


def declared_symbols():
    hdr = open(os.path.join(REPO, "include", "flash_attn.h")).read()
    hdr = hdr.split("#ifdef __cplusplus\n} /* extern")[0]          # C part only
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(flash_attn_\w+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(fa.EXPORTED_SYMBOLS)


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.flash_attn_version()
    for code in fa.ERRORS:
        assert len(lib.flash_attn_error_string(code)) > 0
    assert lib.flash_attn_error_string(0) == b"success"


def test_argument_validation_happens_before_any_device_work(lib):
    buf = (ctypes.c_uint16 * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    # head_dim other than 64/128 (the reference silently mis-indexes, FA.cu:613)
    assert lib.flash_attn_fwd(p, p, p, p, 1, 1, 16, 96, 1, None) == -1
    assert lib.flash_attn_fwd(None, p, p, p, 1, 1, 16, 128, 1, None) == -2
    assert lib.flash_attn_fwd(p, p, p, p, 0, 1, 16, 128, 1, None) == -4
    assert lib.flash_attn_fwd(p, p, p, p, 1, 1, 0, 128, 1, None) == -4
    mis = ctypes.c_void_p(p.value + 2)
    assert lib.flash_attn_fwd(mis, p, p, p, 1, 1, 16, 128, 1, None) == -3
    assert lib.flash_attn_fwd_ex(p, p, p, p, None, 1, 1, 16, 16, 128, 1, 0, 0, 0, None) == -2
    assert lib.flash_attn_finalize(p, p, p, 4, 100, None) == -1
    assert lib.flash_attn_fwd_host(p, p, p, None, 1, 1, 16, 128, 1) == -2


def test_python_wrapper_refuses_cpu_tensors():
    import torch
    x = torch.zeros(1, 1, 16, 128, dtype=torch.float16)
    with pytest.raises(ValueError):
        fa.flash_attn_fwd(x, x, x)


def test_no_oracle_on_the_product_path():
    # the product sources must not reference the checkers
    for root in (os.path.join(REPO, "flash_attention_cuda_b200"), os.path.join(REPO, "include")):
        for dirpath, _, files in os.walk(root):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "liboracle" not in text and "attn_oracle" not in text, f
                    assert "import _oracle" not in text, f

"""Generates tests/golden/*.npz by running the REFERENCE's own cpu_attention
(/root/reference/flash_attention.cu:668-697, compiled unmodified into oracle/_ref/libref_v9.so by
oracle/Makefile).  Run in the build container (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

Each fixture stores the seeded inputs' recipe and the reference output bits, so the restatement in
oracle/attn_oracle.c is pinned against the real reference without needing it at test time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle  # noqa: E402

CASES = [
    # name, (B,H,N,D), causal, generator
    ("refrand_h2_n256_d128_causal", (1, 2, 256, 128), 1, "refrand"),
    ("refrand_h2_n256_d128_full", (1, 2, 256, 128), 0, "refrand"),
    ("refrand_h1_n100_d64_causal", (1, 1, 100, 64), 1, "refrand"),
    ("normal_b2_h2_n129_d128_causal", (2, 2, 129, 128), 1, "normal"),
    ("normal_h3_n65_d64_full", (1, 3, 65, 64), 0, "normal"),
    ("normal_h1_n1_d128_causal", (1, 1, 1, 128), 1, "normal"),
]


def inputs(shape, gen, seed=1234):
    if gen == "refrand":
        return _oracle.fill_ref_rand(shape, 42)
    rng = np.random.default_rng(seed)
    return tuple(rng.standard_normal(shape, dtype=np.float32).astype(np.float16) for _ in range(3))


if __name__ == "__main__":
    assert _oracle.ref() is not None, "build oracle/_ref first (make -C oracle)"
    for name, shape, causal, gen in CASES:
        q, k, v = inputs(shape, gen)
        o = _oracle.ref_attention(q, k, v, causal)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), q=q.view(np.uint16), k=k.view(np.uint16),
                            v=v.view(np.uint16), o=o.view(np.uint16), causal=np.int32(causal))
        print(name, o.shape, _oracle.checksum(o))

"""Generate tests/golden/ref_vectors.npz from the REFERENCE ITSELF.

Run in the build container (needs oracle/_ref/libvrdx_ref.so, i.e. /root/reference compiled
by oracle/Makefile).  Inputs come from the reference's DataGenerator(seed).Generate(n, bits)
(bench/data_generator.cc:12-27); outputs from CpuBenchmark::Sort / SortKeyValue
(bench/cpu_benchmark.cc:19-53) — the check the reference applies to its own GPU path
(bench/bench.cc:41-64).  The fixture travels to the GPU box, /root/reference does not.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu_oracle as o  # noqa: E402

# (seed, n, bits): sizes straddle the reference's 4096-key partition and our tile sizes
CASES = [
    (42, 1, 32), (42, 2, 32), (7, 33, 32), (7, 257, 32), (1, 4095, 32), (1, 4096, 32), (1, 4097, 32),
    (2, 8191, 32), (2, 8193, 32), (3, 10000, 32), (4, 9001, 8), (5, 9001, 4), (6, 5000, 0),
    (8, 12289, 1), (9, 20011, 16),
]


def main():
    assert o.have_ref(), "build oracle/_ref first: make -C oracle"
    out = {}
    for seed, n, bits in CASES:
        k, v = o.ref_generate(seed, n, bits)
        sk, _ = o.ref_sort_keys(k)
        kk, kv, _ = o.ref_sort_key_value(k, v)
        tag = f"s{seed}_n{n}_b{bits}"
        out[tag + "_keys"] = k
        out[tag + "_values"] = v
        out[tag + "_sorted"] = sk
        out[tag + "_kv_keys"] = kk
        out[tag + "_kv_values"] = kv
    # keys equal to the reference's 0xFFFFFFFF padding sentinel, with distinguishable values
    k = np.array([0xFFFFFFFF, 5, 0xFFFFFFFF, 0, 5, 0xFFFFFFFF, 0xFFFFFFFE] * 700, dtype=np.uint32)
    v = np.arange(k.size, dtype=np.uint32)
    kk, kv, _ = o.ref_sort_key_value(k, v)
    out["sentinel_keys"], out["sentinel_values"] = k, v
    out["sentinel_kv_keys"], out["sentinel_kv_values"] = kk, kv
    out["cases"] = np.array(CASES, dtype=np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

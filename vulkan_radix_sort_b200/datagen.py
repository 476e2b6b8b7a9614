"""Synthetic sort inputs — host-side mirror of the reference's bench/data_generator.cc.

``DataGenerator(seed).generate(size, bits)`` reproduces, bit for bit, what the reference's
``DataGenerator(int seed).Generate(size, bits)`` returns when built with libstdc++
(bench/data_generator.cc:8-27): ``std::mt19937(seed)``; keys = the first ``size`` draws of
``uniform_int_distribution<uint32_t>(0, 2^bits-1)``, values = the next ``size`` full-range
draws.  With libstdc++ a full-range draw is the raw MT19937 output and a power-of-two
range draw is the top ``bits`` bits of the raw output (Lemire multiply-shift with a zero
rejection threshold), one engine call per draw either way.  tests/test_oracle.py pins this
against the compiled reference generator.

The extra distributions are the adversarial inputs named in BASELINE.md §3 / SURVEY.md §8(d).
"""
from __future__ import annotations

import numpy as np


class DataGenerator:
    """Mirror of the reference's DataGenerator (bench/data_generator.h:13-24)."""

    def __init__(self, seed: int = 1):
        self._bitgen = np.random.MT19937()
        # std::mt19937(seed) seeding == Knuth init_genrand(seed) == numpy "legacy" seeding
        self._bitgen._legacy_seeding(int(seed) & 0xFFFFFFFF)

    def raw(self, n: int) -> np.ndarray:
        n = int(n)
        out = np.empty(n, dtype=np.uint32)
        chunk = 1 << 24  # random_raw yields uint64: bound the transient footprint
        for a in range(0, n, chunk):
            b = min(n, a + chunk)
            out[a:b] = self._bitgen.random_raw(b - a)
        return out

    def generate(self, size: int, bits: int = 32):
        """-> (keys, values), like SortData Generate(size, bits) (bench/data_generator.cc:12-27)."""
        raw_keys = self.raw(size)
        if bits >= 32:
            keys = raw_keys
        elif bits <= 0:
            keys = np.zeros(size, dtype=np.uint32)
        else:
            keys = (raw_keys >> np.uint32(32 - bits)).astype(np.uint32)
        values = self.raw(size)
        return keys, values


# Distributions for the adversarial configs (BASELINE.json configs[3]).
DISTRIBUTIONS = (
    "uniform", "skewed", "bits8", "bits4", "all_zero", "all_ones", "sorted", "reverse",
    "sentinel_mix",
)


def make_keys(dist: str, n: int, seed: int = 1) -> np.ndarray:
    """Keys of distribution ``dist`` (uint32, length n)."""
    gen = DataGenerator(seed)
    if dist == "uniform":
        return gen.generate(n, 32)[0]
    if dist == "skewed":  # AND of three consecutive draws: each bit set w.p. 1/8
        r = gen.raw(3 * n).reshape(n, 3)
        return (r[:, 0] & r[:, 1] & r[:, 2]).astype(np.uint32)
    if dist == "bits8":
        return gen.generate(n, 8)[0]
    if dist == "bits4":
        return gen.generate(n, 4)[0]
    if dist == "all_zero":
        return gen.generate(n, 0)[0]
    if dist == "all_ones":  # collides with the reference's 0xFFFFFFFF padding sentinel
        return np.full(n, 0xFFFFFFFF, dtype=np.uint32)
    if dist == "sorted":
        return np.sort(gen.generate(n, 32)[0])
    if dist == "reverse":
        return np.sort(gen.generate(n, 32)[0])[::-1].copy()
    if dist == "sentinel_mix":  # half the keys are exactly the padding sentinel
        k = gen.generate(n, 32)[0]
        m = gen.raw(n) & np.uint32(1)
        return np.where(m == 1, np.uint32(0xFFFFFFFF), k).astype(np.uint32)
    raise ValueError(f"unknown distribution {dist!r}")

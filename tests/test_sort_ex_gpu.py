"""GPU parity of vrdxCudaCmdSortEx (key types, descending order, bit sub-ranges).

The extension has no counterpart in the reference; its definition is oracle.sort_ex (a stable sort by a bit field of
the order-preserving unsigned image of the key), which tests/test_oracle.py pins to NumPy's typed sorts.  Bar: keys
and values bit-exact.  Also here: the regression test for pads of a tail tile in the order-free first pass.
"""
import numpy as np
import pytest
import torch

from vulkan_radix_sort_b200 import api
from vulkan_radix_sort_b200.datagen import DataGenerator, make_keys

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

KEY_TYPES = {"uint32": 0, "int32": 1, "float32": 2}


def run_ex(sorter, bits, values=None, count=None, **kw):
    dk = torch.from_numpy(bits.view(np.int32).copy()).to(DEV)
    dv = torch.from_numpy(values.view(np.int32).copy()).to(DEV) if values is not None else None
    cnt = torch.tensor([count], dtype=torch.int32, device=DEV) if count is not None else None
    sorter.sort_ex(dk, dv, count_buffer=cnt, max_count=bits.size, **kw)
    torch.cuda.synchronize()
    return dk.cpu().numpy().view(np.uint32), (dv.cpu().numpy().view(np.uint32) if dv is not None else None)


def float_bits(n, seed):
    rng = np.random.default_rng(seed)
    f = (rng.standard_normal(n) * 1e3).astype(np.float32)
    f[::17] = 0.0
    f[1::19] = -0.0
    f[2::23] = np.inf
    f[3::29] = -np.inf
    f[4::31] = np.float32(1e-42)
    f[5::41] = np.nan
    b = f.view(np.uint32).copy()
    b[6::43] |= np.uint32(0x80000000)  # some negative NaNs / negated values
    return b


@pytest.mark.parametrize("n", [1, 255, 4096, 6144, 6145, 100_003, 1_300_001])
@pytest.mark.parametrize("key_type", list(KEY_TYPES))
@pytest.mark.parametrize("descending", [False, True])
def test_key_types_and_order_keys_and_pairs(sorter, oracle, n, key_type, descending):
    bits = float_bits(n, n) if key_type == "float32" else DataGenerator(n).generate(n)[0]
    vals = np.arange(n, dtype=np.uint32)
    kt = KEY_TYPES[key_type]
    ek, ev = oracle.sort_ex(bits, vals, key_type=kt, descending=descending)
    gk, _ = run_ex(sorter, bits, key_type=kt, descending=descending)
    assert np.array_equal(gk, ek)
    gk, gv = run_ex(sorter, bits, vals, key_type=kt, descending=descending)
    assert np.array_equal(gk, ek) and np.array_equal(gv, ev)


@pytest.mark.parametrize("two_runs", ["-1", "1"], ids=["two_runs_auto", "two_runs_always"])
@pytest.mark.parametrize("key_type,descending", [("int32", False), ("float32", True), ("uint32", True)])
def test_block_free_tiles_with_key_codecs(oracle, key_type, descending, two_runs, monkeypatch):
    """Keys-only reduce-then-scan over all 32 bits of typed / descending keys: the GENERIC instantiations of the
    block-free tile paths (csrc/vrdx_kernels.cuh TileBlockFree; one-run tiles in pass 0, one- and two-run tiles in
    pass 1 at this size, and — forced — in every later pass of the low-entropy input)."""
    from vulkan_radix_sort_b200 import Sorter
    monkeypatch.setenv("VRDX_TWO_RUNS", two_runs)
    s = Sorter(0, algorithm=api.VRDX_CUDA_ALGORITHM_REDUCE_THEN_SCAN)
    kt = KEY_TYPES[key_type]
    n = (1 << 21) + 777
    rng = np.random.default_rng(11)
    inputs = [float_bits(n, 7) if key_type == "float32" else DataGenerator(12).generate(n)[0]]
    low = rng.choice(1 << 16, size=40, replace=False).astype(np.uint32)          # long runs below the third digit
    inputs.append((low[rng.integers(0, 40, n)] | (rng.integers(0, 1 << 16, n, dtype=np.uint32) << np.uint32(16))).astype(np.uint32))
    for bits in inputs:
        ek, _ = oracle.sort_ex(bits, np.arange(n, dtype=np.uint32), key_type=kt, descending=descending)
        gk, _ = run_ex(s, bits, key_type=kt, descending=descending)
        assert np.array_equal(gk, ek)
    s.close()


def test_typed_tensors_sort_like_torch(sorter):
    g = torch.Generator(device="cpu").manual_seed(5)
    f = torch.randn(500_000, generator=g).to(DEV)
    mine = f.clone()
    sorter.sort_ex(mine)                                   # dtype float32 -> FLOAT32 keys
    assert torch.equal(mine, torch.sort(f).values)
    i = torch.randint(-2**31, 2**31 - 1, (500_000,), generator=g, dtype=torch.int64).to(torch.int32).to(DEV)
    mine = i.clone()
    sorter.sort_ex(mine, descending=True)                  # dtype int32 -> INT32 keys
    assert torch.equal(mine, torch.sort(i, descending=True).values)


@pytest.mark.parametrize("bits_range", [(0, 8), (8, 16), (0, 16), (8, 24), (4, 13), (20, 32), (3, 32), (0, 31), (31, 32),
                                        (0, 1), (5, 5)])
@pytest.mark.parametrize("n", [777, 70_001, 2_000_003])
def test_bit_sub_range_is_stable_on_the_other_bits(any_sorter, oracle, bits_range, n):
    b, e = bits_range
    bits = DataGenerator(b * 37 + e).generate(n)[0]
    vals = np.arange(n, dtype=np.uint32)
    ek, ev = oracle.sort_ex(bits, vals, begin_bit=b, end_bit=e)
    gk, _ = run_ex(any_sorter, bits, begin_bit=b, end_bit=e, key_type=0)
    assert np.array_equal(gk, ek)                          # keys-only must be stable too: other bits differ
    gk, gv = run_ex(any_sorter, bits, vals, begin_bit=b, end_bit=e, key_type=0)
    assert np.array_equal(gk, ek) and np.array_equal(gv, ev)


@pytest.mark.parametrize("key_type,descending,bits_range", [("float32", True, (0, 32)), ("int32", False, (16, 32)),
                                                            ("float32", False, (12, 28)), ("uint32", True, (0, 24))])
def test_every_flavour_and_indirect_count(any_sorter, oracle, key_type, descending, bits_range):
    mx, count = 400_007, 250_013
    b, e = bits_range
    bits = float_bits(mx, 3) if key_type == "float32" else DataGenerator(9).generate(mx)[0]
    vals = np.arange(mx, dtype=np.uint32)[::-1].copy()
    kt = KEY_TYPES[key_type]
    ek, ev = oracle.sort_ex(bits[:count], vals[:count], key_type=kt, descending=descending, begin_bit=b, end_bit=e)
    gk, gv = run_ex(any_sorter, bits, vals, count=count, key_type=kt, descending=descending, begin_bit=b, end_bit=e)
    assert np.array_equal(gk[:count], ek) and np.array_equal(gv[:count], ev)
    assert np.array_equal(gk[count:], bits[count:]) and np.array_equal(gv[count:], vals[count:])  # tail untouched
    gk, _ = run_ex(any_sorter, bits, count=count, key_type=kt, descending=descending, begin_bit=b, end_bit=e)
    assert np.array_equal(gk[:count], ek) and np.array_equal(gk[count:], bits[count:])


@pytest.mark.parametrize("dist", ["all_zero", "all_ones", "bits4", "sorted", "sentinel_mix"])
def test_low_entropy_inputs_with_codec(sorter, oracle, dist):
    n = 300_001
    bits = make_keys(dist, n, 2)
    vals = np.arange(n, dtype=np.uint32)
    for kt, desc, (b, e) in ((1, True, (0, 32)), (2, False, (0, 32)), (0, False, (24, 32)), (1, False, (28, 32))):
        ek, ev = oracle.sort_ex(bits, vals, key_type=kt, descending=desc, begin_bit=b, end_bit=e)
        gk, gv = run_ex(sorter, bits, vals, key_type=kt, descending=desc, begin_bit=b, end_bit=e)
        assert np.array_equal(gk, ek) and np.array_equal(gv, ev), (dist, kt, desc, b, e)


def test_default_key_info_equals_vrdx_cmd_sort(sorter, oracle):
    n = 123_457
    k, v = DataGenerator(4).generate(n)
    gk, gv = run_ex(sorter, k, v, key_type=0)
    ek, ev = oracle.sort_key_value(k, v)
    assert np.array_equal(gk, ek) and np.array_equal(gv, ev)


def test_invalid_key_info_is_a_sticky_error(sorter):
    d = torch.zeros(16, dtype=torch.int32, device=DEV)
    with pytest.raises(RuntimeError):
        sorter.sort_ex(d, begin_bit=9, end_bit=8)
    with pytest.raises(RuntimeError):
        sorter.sort_ex(d, end_bit=33)
    with pytest.raises(RuntimeError):
        sorter.sort_ex(d, key_type=7)
    sorter.sort_ex(d)  # the sorter is still usable
    torch.cuda.synchronize()


def test_launch_counts_follow_the_pass_count(sorter):
    d = torch.randint(0, 2**31 - 1, (100_000,), dtype=torch.int32, device=DEV)
    sorter.sort_ex(d.clone(), key_type=0)
    full = sorter.last_launch_count
    sorter.sort_ex(d.clone(), key_type=0, begin_bit=0, end_bit=16)
    assert sorter.last_launch_count == full - 2            # two passes fewer
    sorter.sort_ex(d.clone(), key_type=0, begin_bit=0, end_bit=8)
    assert sorter.last_launch_count == full - 3 + 1        # one pass + the copy back
    sorter.sort_ex(d.clone(), key_type=0, begin_bit=4, end_bit=4)
    assert sorter.last_launch_count == 0
    torch.cuda.synchronize()


# ---- regression: pads of the tail tile in the order-free first pass of a keys-only sort

@pytest.mark.parametrize("algorithm", ["ONESWEEP", "REDUCE_THEN_SCAN"])
def test_tail_tile_real_keys_with_the_pad_digit_are_not_dropped(algorithm, oracle):
    """Every key of the last, partial tile holds digit 0xFF in pass 0 — the digit the 0xFFFFFFFF pads carry.  If pads
    could take slots before real keys (an unordered ranking), real keys would be replaced by 0xFFFFFFFF."""
    from vulkan_radix_sort_b200 import Sorter
    s = Sorter(0, algorithm=getattr(api, "VRDX_CUDA_ALGORITHM_" + algorithm))
    tile = int(s.properties.keysTileSize)
    rng = np.random.default_rng(8)
    try:
        for tail in (1, 5, 31, 33, 100, 513, tile // 2 + 7, tile - 1):
            n = 3 * max(tile, 6144) + tail
            k = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
            k[-tail - 64:] |= np.uint32(0xFF)              # digit 0 == 0xFF, upper bits random
            k[-tail - 64:] &= np.uint32(0x7FFFFFFF)        # ... and none of them equals the pad itself
            d = torch.from_numpy(k.view(np.int32)).to(DEV)
            for _ in range(3):
                w = d.clone()
                s.sort(w)
                torch.cuda.synchronize()
                assert np.array_equal(w.cpu().numpy().view(np.uint32), np.sort(k)), (algorithm, tail)
    finally:
        s.close()


# ---- 64-bit keys (two chained key-value sorts: low word, then high word)

def _bits64(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "float64":
        f = rng.standard_normal(n) * 1e6
        f[::11] = np.inf
        f[1::13] = -np.inf
        f[2::17] = 5e-324
        f[3::19] = 0.0
        f[4::23] = -0.0
        f[5::29] = np.nan
        return f.view(np.uint64).copy()
    u = rng.integers(0, 1 << 64, n, dtype=np.uint64)
    u[::7] &= np.uint64(0xFFFFFFFF)
    u[1::7] &= np.uint64(0xFFFFFFFF00000000)
    u[2::7] = u[0]                                        # exact duplicates
    return u


@pytest.mark.parametrize("n", [1, 4097, 300_007, 5_000_001])
@pytest.mark.parametrize("kind", ["uint64", "int64", "float64"])
@pytest.mark.parametrize("descending", [False, True])
def test_keys64_bit_exact(sorter, oracle, n, kind, descending):
    bits = _bits64(kind, n, n)
    kt = {"uint64": 0, "int64": 1, "float64": 2}[kind]
    want = oracle.sort_keys64(bits, key_type=kt, descending=descending)
    d = torch.from_numpy(bits.view(np.int64).copy()).to(DEV)
    sorter.sort_keys64(d, key_type=kt, descending=descending)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint64), want)


def test_keys64_typed_tensors_and_indirect_count(any_sorter, oracle):
    g = torch.Generator(device="cpu").manual_seed(3)
    f = torch.randn(400_003, generator=g, dtype=torch.float64).to(DEV)
    mine = f.clone()
    any_sorter.sort_keys64(mine)
    assert torch.equal(mine, torch.sort(f).values)
    i = torch.randint(-2**62, 2**62, (400_003,), generator=g, dtype=torch.int64).to(DEV)
    count = 250_001
    cnt = torch.tensor([count], dtype=torch.int32, device=DEV)
    mine = i.clone()
    any_sorter.sort_keys64(mine, descending=True, count_buffer=cnt, max_count=i.numel())
    torch.cuda.synchronize()
    assert torch.equal(mine[:count], torch.sort(i[:count], descending=True).values)
    assert torch.equal(mine[count:], i[count:])            # tail untouched

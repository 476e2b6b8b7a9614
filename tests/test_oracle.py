"""CPU tests: the oracle is pinned to the reference before anything trusts it.

1. against the committed golden vectors (tests/golden/ref_vectors.npz, produced by the
   reference's own CpuBenchmark + DataGenerator, see tests/golden/make_golden.py);
2. where oracle/_ref is present, against the reference library live on fresh inputs;
3. the host-side generator mirror (vulkan_radix_sort_b200.datagen) against the reference's.
"""
import os

import numpy as np
import pytest

from vulkan_radix_sort_b200.datagen import DISTRIBUTIONS, DataGenerator, make_keys

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _cases(golden):
    return [tuple(int(x) for x in row) for row in golden["cases"]]


def test_golden_fixture_present(golden):
    assert len(_cases(golden)) >= 10


def test_oracle_matches_golden_vectors(oracle, golden):
    for seed, n, bits in _cases(golden):
        tag = f"s{seed}_n{n}_b{bits}"
        k, v = golden[tag + "_keys"], golden[tag + "_values"]
        assert np.array_equal(oracle.sort_keys(k), golden[tag + "_sorted"]), tag
        ok, ov = oracle.sort_key_value(k, v)
        assert np.array_equal(ok, golden[tag + "_kv_keys"]), tag
        assert np.array_equal(ov, golden[tag + "_kv_values"]), tag  # pins stability
        pk, pv = oracle.sort_partitioned(k, v, n)
        assert np.array_equal(pk, golden[tag + "_kv_keys"]), tag
        assert np.array_equal(pv, golden[tag + "_kv_values"]), tag


def test_oracle_sentinel_keys_keep_their_values(oracle, golden):
    k, v = golden["sentinel_keys"], golden["sentinel_values"]
    for fn in (oracle.sort_key_value, lambda a, b: oracle.sort_partitioned(a, b, a.size)):
        ok, ov = fn(k, v)
        assert np.array_equal(ok, golden["sentinel_kv_keys"])
        assert np.array_equal(ov, golden["sentinel_kv_values"])


def test_generator_mirror_matches_golden_inputs(golden):
    for seed, n, bits in _cases(golden):
        tag = f"s{seed}_n{n}_b{bits}"
        k, v = DataGenerator(seed).generate(n, bits)
        assert np.array_equal(k, golden[tag + "_keys"]), tag
        assert np.array_equal(v, golden[tag + "_values"]), tag


def test_generator_first_key_seed42():
    # DataGenerator(42).Generate(n) starts with 1608637542 under libstdc++
    assert int(DataGenerator(42).generate(4)[0][0]) == 1608637542


def test_structural_restatement_equals_net_effect(oracle):
    # partition boundaries of the reference (4096) and tail handling, all distributions
    for dist in DISTRIBUTIONS:
        for n in (1, 2, 4095, 4096, 4097, 12288, 12289, 30001):
            k = make_keys(dist, n, seed=11)
            v = np.arange(n, dtype=np.uint32)
            a_k, a_v = oracle.sort_key_value(k, v)
            b_k, b_v = oracle.sort_partitioned(k, v, n)
            assert np.array_equal(a_k, b_k) and np.array_equal(a_v, b_v), (dist, n)
            assert np.array_equal(a_k, np.sort(k, kind="stable"))
            assert np.array_equal(a_v, np.argsort(k, kind="stable").astype(np.uint32))


def test_indirect_contract_tail_untouched(oracle):
    # count < max: only [0,count) is sorted, [count,max) keeps the caller's data
    mx, count = 12288, 4113
    k = make_keys("uniform", mx, seed=3)
    v = make_keys("uniform", mx, seed=4)
    ok, ov = oracle.sort_partitioned(k, v, count)
    ek, ev = oracle.sort_key_value(k[:count], v[:count])
    assert np.array_equal(ok[:count], ek) and np.array_equal(ov[:count], ev)
    assert np.array_equal(ok[count:], k[count:]) and np.array_equal(ov[count:], v[count:])
    ok2, _ = oracle.sort_partitioned(k, None, 0)
    assert np.array_equal(ok2, k)


def test_property_checkers(oracle):
    k = make_keys("bits8", 50000, seed=5)
    v = np.arange(k.size, dtype=np.uint32)
    sk, sv = oracle.sort_key_value(k, v)
    assert oracle.is_sorted(sk) and not oracle.is_sorted(k)
    assert oracle.multiset_fingerprint(k, v) == oracle.multiset_fingerprint(sk, sv)
    assert oracle.check_stable_permutation(k, sk, sv)
    bad = sv.copy()
    i = int(np.flatnonzero(sk[1:] == sk[:-1])[0])
    bad[i], bad[i + 1] = bad[i + 1], bad[i]  # break stability only
    assert not oracle.check_stable_permutation(k, sk, bad)
    assert oracle.multiset_fingerprint(sk, bad) == oracle.multiset_fingerprint(sk, sv)  # same multiset
    lost = sv.copy()
    lost[7] = lost[8]  # a dropped / duplicated element changes the multiset
    assert oracle.multiset_fingerprint(sk, lost) != oracle.multiset_fingerprint(sk, sv)


def test_oracle_against_live_reference(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this machine")
    for seed, n, bits in [(1, 1 << 18, 32), (2, 100003, 32), (3, 65537, 8), (4, 70000, 0)]:
        k, v = oracle.ref_generate(seed, n, bits)
        mk, mv = DataGenerator(seed).generate(n, bits)
        assert np.array_equal(k, mk) and np.array_equal(v, mv)
        rk, _ = oracle.ref_sort_keys(k)
        assert np.array_equal(oracle.sort_keys(k), rk)
        rkk, rkv, _ = oracle.ref_sort_key_value(k, v)
        ok, ov = oracle.sort_key_value(k, v)
        assert np.array_equal(ok, rkk) and np.array_equal(ov, rkv)
    for dist in ("skewed", "all_ones", "sentinel_mix", "sorted", "reverse"):
        k = make_keys(dist, 50001, seed=9)
        v = np.arange(k.size, dtype=np.uint32)
        rkk, rkv, _ = oracle.ref_sort_key_value(k, v)
        ok, ov = oracle.sort_partitioned(k, v, k.size)
        assert np.array_equal(ok, rkk) and np.array_equal(ov, rkv), dist


# ---- key-type / order / bit-range extension: the oracle's definition is pinned to NumPy's typed sorts

def _special_floats(rng, n):
    f = rng.standard_normal(n).astype(np.float32) * np.float32(1e3)
    f[:: 17] = 0.0
    f[1:: 19] = -0.0
    f[2:: 23] = np.inf
    f[3:: 29] = -np.inf
    f[4:: 31] = np.float32(1e-42)      # subnormal
    f[5:: 37] = np.float32(-1e-42)
    return f


def test_sort_ex_matches_numpy_typed_sorts(oracle):
    rng = np.random.default_rng(11)
    u = rng.integers(0, 1 << 32, 50_000, dtype=np.uint64).astype(np.uint32)
    got, _ = oracle.sort_ex(u)
    assert np.array_equal(got, np.sort(u)) and np.array_equal(got, oracle.sort_keys(u))
    i = u.view(np.int32)
    got, _ = oracle.sort_ex(i, key_type=oracle.KEY_INT32)
    assert np.array_equal(got.view(np.int32), np.sort(i))
    got, _ = oracle.sort_ex(i, key_type=oracle.KEY_INT32, descending=True)
    assert np.array_equal(got.view(np.int32), np.sort(i)[::-1])
    f = _special_floats(rng, 50_000)
    got, _ = oracle.sort_ex(f, key_type=oracle.KEY_FLOAT32)
    gf = got.view(np.float32)
    assert np.array_equal(gf, np.sort(f))                     # -0.0 == +0.0 for NumPy's comparison ...
    z = np.flatnonzero(gf == 0.0)
    assert np.all(np.diff(np.signbit(gf[z]).astype(np.int8)) <= 0)  # ... and the total order puts -0 before +0
    got, _ = oracle.sort_ex(f, key_type=oracle.KEY_FLOAT32, descending=True)
    assert np.array_equal(got.view(np.float32), np.sort(f)[::-1])


def test_sort_ex_nan_total_order(oracle):
    f = np.array([1.0, np.nan, -1.0, -np.nan, np.inf, -np.inf, 0.0], dtype=np.float32)
    f[3] = np.float32(np.nan).view(np.uint32).__or__(np.uint32(0x80000000)).view(np.float32)  # negative NaN
    got, _ = oracle.sort_ex(f, key_type=oracle.KEY_FLOAT32)
    g = got.view(np.float32)
    assert np.isnan(g[0]) and np.signbit(g[0]) and np.isnan(g[-1]) and not np.signbit(g[-1])
    assert np.array_equal(g[1:-1], np.array([-np.inf, -1.0, 0.0, 1.0, np.inf], dtype=np.float32))


def test_sort_ex_bit_range_is_a_stable_field_sort(oracle):
    rng = np.random.default_rng(12)
    u = rng.integers(0, 1 << 32, 20_000, dtype=np.uint64).astype(np.uint32)
    v = np.arange(u.size, dtype=np.uint32)
    for b, e in ((0, 8), (8, 24), (4, 13), (20, 32), (0, 32), (7, 7)):
        k, p = oracle.sort_ex(u, v, begin_bit=b, end_bit=e)
        field = (k.astype(np.uint64) >> np.uint64(b)) & np.uint64((1 << (e - b)) - 1)
        assert np.all(np.diff(field.astype(np.int64)) >= 0)
        same = np.diff(field.astype(np.int64)) == 0
        assert np.all(np.diff(p.astype(np.int64))[same] > 0)          # ties keep input order
        assert np.array_equal(u[p], k)
    k, p = oracle.sort_ex(u, v, begin_bit=7, end_bit=7)
    assert np.array_equal(k, u) and np.array_equal(p, v)               # empty range: nothing moves


def test_sort_keys64_matches_numpy_typed_sorts(oracle):
    rng = np.random.default_rng(13)
    u = rng.integers(0, 1 << 64, 30_000, dtype=np.uint64)
    u[::7] &= np.uint64(0xFFFFFFFF)                     # equal high words: the low word decides
    u[1::7] &= np.uint64(0xFFFFFFFF00000000)            # equal low words
    assert np.array_equal(oracle.sort_keys64(u), np.sort(u))
    i = u.view(np.int64)
    assert np.array_equal(oracle.sort_keys64(i, key_type=oracle.KEY_INT32).view(np.int64), np.sort(i))
    assert np.array_equal(oracle.sort_keys64(i, key_type=oracle.KEY_INT32, descending=True).view(np.int64), np.sort(i)[::-1])
    f = rng.standard_normal(30_000) * 1e6
    f[::11] = np.inf
    f[1::13] = -np.inf
    f[2::17] = 5e-324                                    # subnormal
    assert np.array_equal(oracle.sort_keys64(f, key_type=oracle.KEY_FLOAT32).view(np.float64), np.sort(f))
    assert np.array_equal(oracle.sort_keys64(f, key_type=oracle.KEY_FLOAT32, descending=True).view(np.float64), np.sort(f)[::-1])

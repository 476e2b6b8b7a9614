"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle.

Bit-exact bar (integer work): keys AND values must equal the oracle's stable sort, which
tests/test_oracle.py pins to the reference's CpuBenchmark (bench/cpu_benchmark.cc:19-53) — the
very check the reference applies to its GPU path (bench/bench.cc:41-64).  At BASELINE.json's
full sizes the check is by size-independent properties (sortedness, multiset fingerprint,
stability through an identity payload).
"""
import os

import numpy as np
import pytest
import torch

from vulkan_radix_sort_b200 import api
from vulkan_radix_sort_b200.datagen import DISTRIBUTIONS, DataGenerator, make_keys

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz")
DEV = "cuda:0"


def to_dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(a.view(np.int32)).to(DEV)


def to_np(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().view(np.uint32)


def gpu_sort_keys(sorter, keys):
    d = to_dev(keys)
    sorter.sort(d)
    torch.cuda.synchronize()
    return to_np(d)


def gpu_sort_kv(sorter, keys, values):
    dk, dv = to_dev(keys), to_dev(values)
    sorter.sort_key_value(dk, dv)
    torch.cuda.synchronize()
    return to_np(dk), to_np(dv)


def tile_sizes(sorter):
    p = sorter.properties
    return int(p.keysTileSize), int(p.keyValueTileSize)


# ------------------------------------------------------------------ golden vectors

def test_golden_vectors_from_reference(sorter):
    g = np.load(GOLDEN)
    for seed, n, bits in (tuple(int(x) for x in r) for r in g["cases"]):
        tag = f"s{seed}_n{n}_b{bits}"
        k, v = g[tag + "_keys"], g[tag + "_values"]
        assert np.array_equal(gpu_sort_keys(sorter, k), g[tag + "_sorted"]), tag
        ok, ov = gpu_sort_kv(sorter, k, v)
        assert np.array_equal(ok, g[tag + "_kv_keys"]), tag
        assert np.array_equal(ov, g[tag + "_kv_values"]), tag
    ok, ov = gpu_sort_kv(sorter, g["sentinel_keys"], g["sentinel_values"])
    assert np.array_equal(ok, g["sentinel_kv_keys"]) and np.array_equal(ov, g["sentinel_kv_values"])


# ------------------------------------------------------------------ small-N functional sweep

def test_small_n_sweep_keys_and_pairs(any_sorter, oracle):
    sorter = any_sorter
    tk, tkv = tile_sizes(sorter)
    sizes = {1, 2, 3, 31, 32, 33, 255, 256, 257, 511, 512, 513, 4095, 4096, 4097,
             tk - 1, tk, tk + 1, tkv - 1, tkv, tkv + 1, 2 * tk + 1, 3 * tkv - 5, 262143, 1 << 18}
    for n in sorted(sizes):
        k, v = DataGenerator(100 + n % 7).generate(n)
        assert np.array_equal(gpu_sort_keys(sorter, k), oracle.sort_keys(k)), n
        ok, ov = gpu_sort_kv(sorter, k, v)
        ek, ev = oracle.sort_key_value(k, v)
        assert np.array_equal(ok, ek) and np.array_equal(ov, ev), n


def test_zero_elements_is_a_noop(sorter):
    k = to_dev(np.array([3, 1, 2], dtype=np.uint32))
    storage = torch.empty(sorter.storage_requirements(3).size, dtype=torch.uint8, device=DEV)
    sorter.sort(k, count=0, storage=storage)
    torch.cuda.synchronize()
    assert to_np(k).tolist() == [3, 1, 2]
    assert sorter.last_launch_count == 0


@pytest.mark.parametrize("dist", DISTRIBUTIONS)
def test_distributions_bit_exact(any_sorter, oracle, dist):
    sorter = any_sorter
    for n in (100003, (1 << 20) + 17):
        k = make_keys(dist, n, seed=21)
        v = np.arange(n, dtype=np.uint32)  # identity payload makes stability directly visible
        assert np.array_equal(gpu_sort_keys(sorter, k), oracle.sort_keys(k)), (dist, n)
        ok, ov = gpu_sort_kv(sorter, k, v)
        ek, ev = oracle.sort_key_value(k, v)
        assert np.array_equal(ok, ek), (dist, n)
        assert np.array_equal(ov, ev), (dist, n)


def test_matches_structural_restatement(sorter, oracle):
    # the partition-structured restatement of upsweep/spine/downsweep gives the same bytes
    n = 50001
    k, v = DataGenerator(5).generate(n)
    ok, ov = gpu_sort_kv(sorter, k, v)
    pk, pv = oracle.sort_partitioned(k, v, n)
    assert np.array_equal(ok, pk) and np.array_equal(ov, pv)


def test_live_reference_when_available(sorter, oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not shipped")
    n = 1 << 18  # the reference's own verification size (bench/bench.cc:17,164-166)
    k, v = oracle.ref_generate(1, n, 32)
    rk, _ = oracle.ref_sort_keys(k)
    assert np.array_equal(gpu_sort_keys(sorter, k), rk)
    rkk, rkv, _ = oracle.ref_sort_key_value(k, v)
    ok, ov = gpu_sort_kv(sorter, k, v)
    assert np.array_equal(ok, rkk) and np.array_equal(ov, rkv)


# ------------------------------------------------------------------ indirect variants

@pytest.mark.parametrize("count_frac", [1.0, 0.5, 0.0003, 0.0])
def test_indirect_count_below_max_leaves_tail_untouched(any_sorter, oracle, count_frac):
    sorter = any_sorter
    mx = 300007
    count = int(mx * count_frac)
    k, v = DataGenerator(31).generate(mx)
    cnt = torch.tensor([count], dtype=torch.int32, device=DEV)
    # keys-only indirect
    dk = to_dev(k)
    sorter.sort_indirect(dk, cnt, max_count=mx)
    torch.cuda.synchronize()
    out = to_np(dk)
    assert np.array_equal(out[:count], oracle.sort_keys(k[:count]))
    assert np.array_equal(out[count:], k[count:])
    # key-value indirect
    dk, dv = to_dev(k), to_dev(v)
    sorter.sort_key_value_indirect(dk, dv, cnt, max_count=mx)
    torch.cuda.synchronize()
    ok, ov = to_np(dk), to_np(dv)
    ek, ev = oracle.sort_partitioned(k, v, count)  # structural restatement carries the tail contract
    assert np.array_equal(ok, ek) and np.array_equal(ov, ev)


def test_indirect_count_above_max_is_clamped(sorter, oracle):
    mx = 70001
    k, _ = DataGenerator(8).generate(mx)
    cnt = torch.tensor([mx + 12345], dtype=torch.int32, device=DEV)
    dk = to_dev(k)
    sorter.sort_indirect(dk, cnt, max_count=mx)
    torch.cuda.synchronize()
    assert np.array_equal(to_np(dk), oracle.sort_keys(k))


def test_single_allocation_with_offsets_like_vulkan_benchmark(sorter, oracle):
    # bench/vulkan_benchmark.cc:356-358,386-388: keys, values and the count live in ONE buffer at
    # offsets 0, inout_size, 2*inout_size; the same VkBuffer is passed three times.
    n = 123457
    inout = (4 * n + 15) // 16 * 16
    k, v = DataGenerator(77).generate(n)
    buf = torch.zeros(2 * inout + 16, dtype=torch.uint8, device=DEV)
    buf[0:4 * n] = torch.from_numpy(k.view(np.uint8)).to(DEV)
    buf[inout:inout + 4 * n] = torch.from_numpy(v.view(np.uint8)).to(DEV)
    buf[2 * inout:2 * inout + 4] = torch.from_numpy(np.array([n], dtype=np.uint32).view(np.uint8)).to(DEV)
    storage = torch.empty(sorter.storage_requirements(n, True).size + 48, dtype=torch.uint8, device=DEV)
    stream = torch.cuda.current_stream().cuda_stream
    base = buf.data_ptr()
    api.vrdxCmdSortKeyValueIndirect(stream, sorter.handle, n, base, 2 * inout, base, 0, base, inout,
                                    storage.data_ptr(), 48, None, 0)
    sorter.check()
    torch.cuda.synchronize()
    host = buf.cpu().numpy()
    ok = host[0:4 * n].view(np.uint32)
    ov = host[inout:inout + 4 * n].view(np.uint32)
    ek, ev = oracle.sort_key_value(k, v)
    assert np.array_equal(ok, ek) and np.array_equal(ov, ev)


def test_four_byte_aligned_key_pointer(sorter, oracle):
    # offsets that are only 4-byte aligned still sort correctly (scalar head/tail paths)
    n = 99991
    k, v = DataGenerator(13).generate(n + 3)
    for shift in (1, 2, 3):
        dk = to_dev(k)
        storage = torch.empty(sorter.storage_requirements(n).size, dtype=torch.uint8, device=DEV)
        api.vrdxCmdSort(torch.cuda.current_stream().cuda_stream, sorter.handle, n, dk.data_ptr(), 4 * shift,
                        storage.data_ptr(), 0, None, 0)
        sorter.check()
        torch.cuda.synchronize()
        out = to_np(dk)
        assert np.array_equal(out[shift:shift + n], oracle.sort_keys(k[shift:shift + n]))
        assert np.array_equal(out[:shift], k[:shift]) and np.array_equal(out[shift + n:], k[shift + n:])


# ------------------------------------------------------------------ API behaviour

def test_storage_reuse_across_sizes_and_kinds(any_sorter, oracle):
    sorter = any_sorter
    # one storage buffer, sized for the largest sort, reused uninitialised (garbage-filled) by
    # smaller sorts of both kinds — the sort must reset all of its own state in-stream.
    big = 400001
    storage = torch.empty(sorter.storage_requirements(big, True).size, dtype=torch.uint8, device=DEV)
    storage.fill_(0xA5)
    for n in (big, 17, 70001, big, 8193):
        k, v = DataGenerator(n % 11).generate(n)
        dk = to_dev(k)
        sorter.sort(dk, storage=storage)
        dk2, dv2 = to_dev(k), to_dev(v)
        sorter.sort_key_value(dk2, dv2, storage=storage)
        torch.cuda.synchronize()
        assert np.array_equal(to_np(dk), oracle.sort_keys(k)), n
        ek, ev = oracle.sort_key_value(k, v)
        assert np.array_equal(to_np(dk2), ek) and np.array_equal(to_np(dv2), ev), n


def test_idempotent_on_sorted_input(sorter, oracle):
    k, v = DataGenerator(3).generate(200001)
    ok, ov = gpu_sort_kv(sorter, k, v)
    ok2, ov2 = gpu_sort_kv(sorter, ok, ov)
    assert np.array_equal(ok, ok2) and np.array_equal(ov, ov2)


def test_concurrent_streams_with_distinct_storage(sorter, oracle):
    n = 250007
    streams = [torch.cuda.Stream() for _ in range(3)]
    data = [DataGenerator(40 + i).generate(n) for i in range(3)]
    dev = [(to_dev(k), to_dev(v)) for k, v in data]
    stor = [torch.empty(sorter.storage_requirements(n, True).size, dtype=torch.uint8, device=DEV) for _ in range(3)]
    torch.cuda.synchronize()
    for s, (dk, dv), st in zip(streams, dev, stor):
        with torch.cuda.stream(s):
            sorter.sort_key_value(dk, dv, storage=st, stream=s)
    torch.cuda.synchronize()
    for (k, v), (dk, dv) in zip(data, dev):
        ek, ev = oracle.sort_key_value(k, v)
        assert np.array_equal(to_np(dk), ek) and np.array_equal(to_np(dv), ev)


def test_cuda_graph_capture_and_replay(any_sorter, oracle):
    sorter = any_sorter
    # "record once, replay many": the enqueue must be capture-safe (no host sync, no host reads)
    n = 150001
    k1, v1 = DataGenerator(51).generate(n)
    k2, v2 = DataGenerator(52).generate(n)
    dk, dv = to_dev(k1), to_dev(v1)
    cnt = torch.tensor([n], dtype=torch.int32, device=DEV)
    storage = torch.empty(sorter.storage_requirements(n, True).size, dtype=torch.uint8, device=DEV)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            sorter.sort_key_value_indirect(dk, dv, cnt, max_count=n, storage=storage, stream=s)
    torch.cuda.synchronize()
    for (k, v, c) in ((k1, v1, n), (k2, v2, n // 3)):
        dk.copy_(to_dev(k)); dv.copy_(to_dev(v)); cnt.fill_(c)
        g.replay()
        torch.cuda.synchronize()
        ek, ev = oracle.sort_partitioned(k, v, c)
        assert np.array_equal(to_np(dk), ek) and np.array_equal(to_np(dv), ev)


def test_query_pool_timestamps(sorter):
    n = 1 << 20
    k, _ = DataGenerator(1).generate(n)
    dk = to_dev(k)
    res, pool = api.vrdxCudaCreateQueryPool(api.cuda_device(0), api.QUERY_COUNT)
    assert res == api.VK_SUCCESS
    res, _ = api.vrdxCudaGetQueryPoolResults(pool)
    assert res == api.VK_NOT_READY
    sorter.sort(dk, query_pool=pool, query=0)
    torch.cuda.synchronize()
    res, ts = api.vrdxCudaGetQueryPoolResults(pool)
    assert res == api.VK_SUCCESS
    assert ts[0] == 0 and all(b >= a for a, b in zip(ts, ts[1:])) and ts[14] > 0
    # per-pass "downsweep" intervals exist (reference: bench/vulkan_benchmark.cc:330-337)
    assert all(ts[4 + 3 * p] > ts[3 + 3 * p] for p in range(4))
    api.vrdxCudaDestroyQueryPool(pool)


def test_query_pool_inside_cuda_graph(any_sorter, oracle):
    # GPU-written timestamps are ordinary kernel work, so they can be captured and replayed
    sorter = any_sorter
    n = (1 << 20) + 3
    k, _ = DataGenerator(9).generate(n)
    dk = to_dev(k)
    storage = torch.empty(sorter.storage_requirements(n).size, dtype=torch.uint8, device=DEV)
    res, pool = api.vrdxCudaCreateQueryPool(api.cuda_device(0), 2 * api.QUERY_COUNT)
    assert res == api.VK_SUCCESS
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            sorter.sort(dk, storage=storage, stream=s, query_pool=pool, query=api.QUERY_COUNT)  # second half of the pool
    torch.cuda.synchronize()
    for rep in range(2):
        dk.copy_(to_dev(k))
        g.replay()
        torch.cuda.synchronize()
        assert np.array_equal(to_np(dk), oracle.sort_keys(k))
        res, ts = api.vrdxCudaGetQueryPoolResults(pool, api.QUERY_COUNT, api.QUERY_COUNT)
        assert res == api.VK_SUCCESS
        assert ts[0] == 0 and ts[14] > 0 and all(b >= a for a, b in zip(ts, ts[1:]))
    res, _ = api.vrdxCudaGetQueryPoolResults(pool, 0, api.QUERY_COUNT)
    assert res == api.VK_NOT_READY  # the first half of the pool was never written
    api.vrdxCudaDestroyQueryPool(pool)


def test_errors_are_sticky_not_fatal(sorter):
    api.vrdxCmdSort(None, sorter.handle, 16, None, 0, None, 0, None, 0)  # NULL buffers
    assert api.vrdxCudaGetLastError(sorter.handle) != 0
    assert api.vrdxCudaGetLastError(sorter.handle) == 0  # reading clears
    k = to_dev(np.array([2, 1], dtype=np.uint32))
    sorter.sort(k)
    torch.cuda.synchronize()
    assert to_np(k).tolist() == [1, 2]


def test_launch_count_reported(sorter):
    k = to_dev(DataGenerator(1).generate(50000)[0])
    sorter.sort(k)
    assert sorter.last_launch_count >= 5  # reset + histogram + 4 passes
    torch.cuda.synchronize()


# ------------------------------------------------------------------ reduce-then-scan: block-free tiles

def _rts_sorter(reserved=None):
    from vulkan_radix_sort_b200 import Sorter
    return Sorter(0, algorithm=api.VRDX_CUDA_ALGORITHM_REDUCE_THEN_SCAN, reserved=reserved)


@pytest.mark.parametrize("two_runs", ["-1", "0", "1"], ids=["two_runs_auto", "two_runs_never", "two_runs_always"])
@pytest.mark.parametrize("reserved", [None, (2, 2), (4, 4)], ids=["256x20", "256x16", "512x16"])
def test_block_free_tiles_one_run_two_runs_three_runs(oracle, reserved, two_runs, monkeypatch):
    """Keys-only reduce-then-scan tiles whose keys agree below the digit (one run), or are exactly two such
    runs, take TileBlockFree (csrc/vrdx_kernels.cuh); tiles of three and more runs rank stably.  The crafted
    inputs put all three kinds of tile into passes 1, 2 and 3; the result must equal the oracle's bit for bit."""
    # which passes run the two-run kernel flavours is decided from the count (run length on uniform keys); the
    # developer switch forces them on / off so that every pass of the crafted inputs meets both flavours
    monkeypatch.setenv("VRDX_TWO_RUNS", two_runs)
    s = _rts_sorter(reserved)
    tile = int(s.properties.keysTileSize)
    rng = np.random.default_rng(77)
    cases = []
    for n in ((1 << 21) + 777, 3 << 20, 40 * tile, 40 * tile + 1):
        cases.append(("uniform", rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)))
        for m in (1, 2, 3, 40, 300, 5000):  # distinct values of the low 16 bits: runs of ~n/m keys in pass 2
            low = rng.choice(1 << 16, size=m, replace=False).astype(np.uint32)
            k = low[rng.integers(0, m, n)] | (rng.integers(0, 1 << 16, n, dtype=np.uint32) << np.uint32(16))
            cases.append((f"low16x{m}", k.astype(np.uint32)))
        for m in (2, 7, 200):  # ... of the low 24 bits: long runs in pass 3 as well
            low = rng.choice(1 << 24, size=m, replace=False).astype(np.uint32)
            k = low[rng.integers(0, m, n)] | (rng.integers(0, 1 << 8, n, dtype=np.uint32) << np.uint32(24))
            cases.append((f"low24x{m}", k.astype(np.uint32)))
        # runs of exactly tile, tile/2 and tile+1 keys below the second digit: boundaries on and off the tile edges
        for run in (tile, tile // 2, tile + 1, 2 * tile - 1):
            k = ((np.arange(n, dtype=np.uint64) // run) % 256).astype(np.uint32) | \
                (rng.integers(0, 1 << 24, n, dtype=np.uint32) << np.uint32(8))
            cases.append((f"run{run}", rng.permutation(k)))
    for name, k in cases:
        assert np.array_equal(gpu_sort_keys(s, k), oracle.sort_keys(k)), (name, len(k))
    # the same inputs as key-value sorts never take the block-free path; spot-check that the tables still serve them
    name, k = cases[3]
    v = np.arange(len(k), dtype=np.uint32)
    ok, ov = gpu_sort_kv(s, k, v)
    ek, ev = oracle.sort_key_value(k, v)
    assert np.array_equal(ok, ek) and np.array_equal(ov, ev), name
    s.close()


# ------------------------------------------------------------------ BASELINE.json full sizes

def _property_check(oracle, k_in, k_out, v_out=None):
    assert oracle.is_sorted(k_out)
    if v_out is None:
        assert oracle.multiset_fingerprint(k_in) == oracle.multiset_fingerprint(k_out)
    else:
        assert oracle.check_stable_permutation(k_in, k_out, v_out)


def test_all_flavours_agree_at_2_pow_25(any_sorter, oracle):
    n = 1 << 25
    k, v = DataGenerator(2).generate(n)
    ok, ov = gpu_sort_kv(any_sorter, k, v)
    ek, ev = oracle.sort_key_value(k, v)
    assert np.array_equal(ok, ek) and np.array_equal(ov, ev)


def test_count_of_2_pow_30_and_above_uses_32bit_counts(sorter, oracle):
    # above the 30-bit look-back cells the library switches to reduce-then-scan on its own
    # (the reference's size arithmetic wraps here, src/vk_radix_sort.h.in:105-114)
    n = (1 << 30) + 12345
    free, _ = torch.cuda.mem_get_info()
    if free < 12 * (1 << 30):
        pytest.skip("not enough device memory")
    g = torch.Generator(device=DEV)
    g.manual_seed(7)
    dk = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=DEV, generator=g)
    ref_sum = int(dk.to(torch.int64).sum().item())
    sorter.sort(dk)
    torch.cuda.synchronize()
    u = dk.view(torch.uint32)
    assert int(dk.to(torch.int64).sum().item()) == ref_sum
    # sortedness on the device in slices (unsigned compare through int64)
    step = 1 << 27
    for a in range(0, n - 1, step):
        b = min(n, a + step + 1)
        w = u[a:b].to(torch.int64)
        assert bool((w[1:] >= w[:-1]).all())
    del dk, u
    torch.cuda.empty_cache()


@pytest.mark.parametrize("log2n", [25, 28])
def test_full_size_uniform_keys_and_pairs(sorter, oracle, log2n):
    n = 1 << log2n
    k = DataGenerator(1).generate(n)[0]
    dk = to_dev(k)
    sorter.sort(dk)
    torch.cuda.synchronize()
    out = to_np(dk)
    _property_check(oracle, k, out)
    if log2n == 25:
        assert np.array_equal(out, oracle.sort_keys(k))  # full bit-exact comparison still cheap here
    del dk
    dk = to_dev(k)
    dv = torch.arange(n, dtype=torch.int32, device=DEV)
    sorter.sort_key_value(dk, dv)
    torch.cuda.synchronize()
    _property_check(oracle, k, to_np(dk), to_np(dv))


@pytest.mark.parametrize("n", [(1 << 25) - 1, 1 << 25, (1 << 26) + 4099, (1 << 27) - 1, 1 << 27])
def test_counts_on_both_sides_of_the_auto_crossovers(sorter, oracle, n):
    """AUTO switches from onesweep to reduce-then-scan at 2^25 keys / 2^27 pairs; the storage is laid out for the
    composition that runs (csrc/vrdx_api.cu SorterLayout).  Property check on both sides of both thresholds."""
    k = DataGenerator(n % 97).generate(n)[0]
    dk = to_dev(k)
    sorter.sort(dk)
    torch.cuda.synchronize()
    _property_check(oracle, k, to_np(dk))
    del dk
    dk = to_dev(k)
    dv = torch.arange(n, dtype=torch.int32, device=DEV)
    sorter.sort_key_value(dk, dv)
    torch.cuda.synchronize()
    _property_check(oracle, k, to_np(dk), to_np(dv))


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("dist", ["skewed", "bits8", "bits4", "all_zero", "all_ones", "sorted", "reverse", "sentinel_mix"])
def test_indirect_adversarial_non_power_of_two(sorter, oracle, dist, seed):
    # BASELINE.json configs[3] / SURVEY 8(d) config 4: max = 2^27, device count = 2^27 - 4099, seeds {1, 2, 3}
    if seed > 1 and dist in ("all_zero", "all_ones"):
        pytest.skip("constant input: the seed does not change it")
    mx = 1 << 27
    n = mx - 4099
    k = make_keys(dist, mx, seed=seed)
    cnt = torch.tensor([n], dtype=torch.int32, device=DEV)
    dk = to_dev(k)
    dv = torch.arange(mx, dtype=torch.int32, device=DEV)
    sorter.sort_key_value_indirect(dk, dv, cnt, max_count=mx)
    torch.cuda.synchronize()
    ok, ov = to_np(dk), to_np(dv)
    assert oracle.check_stable_permutation(k[:n], ok[:n], ov[:n])
    assert np.array_equal(ok[n:], k[n:])
    assert np.array_equal(ov[n:], np.arange(n, mx, dtype=np.uint32))
    del dv
    dk = to_dev(k)
    sorter.sort_indirect(dk, cnt, max_count=mx)
    torch.cuda.synchronize()
    out = to_np(dk)
    assert np.array_equal(out[:n], ok[:n]) and np.array_equal(out[n:], k[n:])


def test_two_host_threads_share_one_sorter(sorter, oracle):
    """SURVEY 8(b) threading row: the sorter is immutable after creation, so two host threads may record
    into different command buffers (streams) with different storage at the same time."""
    import threading
    n = 1 << 20
    inputs = [DataGenerator(100 + t).generate(n) for t in range(2)]
    results, errors = [None, None], []

    def work(t):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                storage = torch.empty(sorter.storage_requirements(n, True).size, dtype=torch.uint8, device=DEV)
                for rep in range(8):
                    dk, dv = to_dev(inputs[t][0]), to_dev(inputs[t][1])
                    sorter.sort_key_value(dk, dv, storage=storage, stream=stream)
                stream.synchronize()
                results[t] = (to_np(dk), to_np(dv))
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for t in range(2):
        ek, ev = oracle.sort_key_value(*inputs[t])
        assert np.array_equal(results[t][0], ek) and np.array_equal(results[t][1], ev)


def test_one_process_drives_two_devices(oracle):
    """Kernel attributes (> 48 KB dynamic shared memory) are per device: a second sorter on another GPU of the
    same process must prepare that GPU too (round-1 ADVICE, vrdx_api.cu `static prepared`)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    from vulkan_radix_sort_b200 import Sorter
    n = (1 << 22) + 77          # large enough for the 128 KB lane-private histogram kernel
    outs = []
    sorters = [Sorter(d) for d in (0, 1)]
    k, v = DataGenerator(5).generate(n)
    for d, s in enumerate(sorters):
        with torch.cuda.device(d):
            dk = torch.from_numpy(k.view(np.int32)).to(f"cuda:{d}")
            dv = torch.from_numpy(v.view(np.int32)).to(f"cuda:{d}")
            s.sort_key_value(dk, dv)
            torch.cuda.synchronize(d)
            outs.append((dk.cpu().numpy().view(np.uint32), dv.cpu().numpy().view(np.uint32)))
            # the multi-GPU histogram kernel with its largest shared-memory footprint (12 bits x 4 prefixes = 64 KB)
            hist = torch.zeros(4 << 12, dtype=torch.int32, device=f"cuda:{d}")
            pref = torch.arange(4, dtype=torch.int32, device=f"cuda:{d}")
            api.load_library().vrdxDistCmdPrefixHistogram(torch.cuda.current_stream(d).cuda_stream, s.handle, n,
                                                          dk.data_ptr(), 0, 8, 12, 4, pref.data_ptr(), 0, hist.data_ptr(), 0)
            s.check()
            torch.cuda.synchronize(d)
            got = hist.cpu().numpy().astype(np.int64).reshape(4, 1 << 12)
            ks = np.sort(k)
            for p in range(4):
                sel = ks[(ks >> 20) == p]
                assert np.array_equal(got[p], np.bincount((sel >> 8) & 0xFFF, minlength=1 << 12)), (d, p)
    ek, ev = oracle.sort_key_value(k, v)
    for ok_, ov_ in outs:
        assert np.array_equal(ok_, ek) and np.array_equal(ov_, ev)
    for s in sorters:
        s.close()

"""GPU tests of the multi-GPU building blocks (single GPU) and, when at least two GPUs are visible,
of the whole distributed sort over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(a.view(np.int32)).cuda()


@pytest.fixture(scope="module")
def backends():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dist_numpy_backend import NumpyBackend
    from vulkan_radix_sort_b200.dist import CudaBackend
    b = CudaBackend(0)
    yield b, NumpyBackend()
    b.close()


@pytest.mark.parametrize("dist_name", ["uniform", "bits8", "all_ones", "skewed"])
def test_prefix_histogram_matches_numpy(backends, dist_name):
    from vulkan_radix_sort_b200.datagen import make_keys
    cuda_b, np_b = backends
    for n in (1, 4099, 1000003):
        k = make_keys(dist_name, n, seed=3)
        kt = torch.from_numpy(k.view(np.int32).copy())
        for shift, bits, prefixes in ((24, 8, [0]), (16, 8, [int(k[0]) >> 24, 0xFF, 0]), (8, 8, [int(k[n // 2]) >> 16, 1]),
                                      (0, 8, [int(k[-1]) >> 8, int(k[0]) >> 8, 0xFFFFFF]),
                                      (12, 12, [int(k[0]) >> 24, 0xFF, 0, 1, 2, 3, 4]),
                                      (0, 12, [int(k[-1]) >> 12, int(k[0]) >> 12, 0xFFFFF])):
            p = torch.tensor(prefixes, dtype=torch.int64)
            want = np_b.prefix_histogram(kt, n, shift, bits, p)
            got = cuda_b.prefix_histogram(_dev(k), n, shift, bits, p.cuda()).cpu()
            assert torch.equal(got, want), (dist_name, n, shift, bits)


@pytest.mark.parametrize("dist_name", ["uniform", "bits4", "all_ones", "sentinel_mix"])
def test_class_count_matches_numpy(backends, dist_name):
    from vulkan_radix_sort_b200.datagen import make_keys
    cuda_b, np_b = backends
    for n, splitters in ((1, []), (4097, [7]), (100003, [3, 0x40000000, 0xFFFFFFFF]),
                         ((1 << 20) + 5, [0, 1, 2, 0x7FFFFFFF, 0x80000000, 0xC0000000, 0xFFFFFFFE])):
        k = make_keys(dist_name, n, seed=6)
        spl = torch.tensor(splitters, dtype=torch.int64)
        want = np_b.class_count(torch.from_numpy(k.view(np.int32).copy()), n, spl)
        got = cuda_b.class_count(_dev(k), n, spl.cuda()).cpu()
        assert torch.equal(got, want), (dist_name, n)


@pytest.mark.parametrize("dist_name", ["uniform", "bits4", "all_zero", "sentinel_mix"])
def test_partition_groups_by_class(backends, dist_name):
    from vulkan_radix_sort_b200.datagen import make_keys
    cuda_b, _ = backends
    for n, splitters in ((5, []), (4096, [7]), (100003, [3, 0x40000000, 0xFFFFFFFF]),
                         (1 << 20, [0, 1, 2, 0x7FFFFFFF, 0x80000000, 0xC0000000, 0xFFFFFFFE])):
        k = make_keys(dist_name, n, seed=5)
        u = np.array(splitters, dtype=np.uint64)
        k64 = k.astype(np.uint64)
        cls = (2 * (k64[:, None] > u[None, :]).sum(axis=1) + (k64[:, None] == u[None, :]).any(axis=1)) if u.size \
            else np.zeros(n, dtype=np.int64)
        sizes = np.bincount(cls, minlength=2 * len(splitters) + 1)
        starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
        out = torch.zeros(n, dtype=torch.int32, device="cuda")
        cuda_b.partition(_dev(k), n, torch.tensor(splitters, dtype=torch.int64, device="cuda"),
                         torch.tensor(starts, dtype=torch.int64, device="cuda"), out)
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(np.uint32)
        for c, (a, sz) in enumerate(zip(starts, sizes)):           # same multiset per class, any order inside
            assert np.array_equal(np.sort(got[a:a + sz]), np.sort(k[cls == c])), (dist_name, n, c)


@pytest.mark.parametrize("n", [4099, 300_007, (1 << 21) + 3])
def test_partition_scatter_to_eight_destinations_on_one_gpu(backends, n):
    """The fused partition + exchange kernel with the table an 8-GPU run builds, all eight "peers" being regions of
    one local buffer: destination boundaries at class edges, inside tie classes (twice inside the same one) and an
    empty destination.  Any order inside a class is legal, so regions are compared as multisets."""
    cuda_b, _ = backends
    rng = np.random.default_rng(n)
    k = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    k[rng.random(n) < 0.2] = 0x40000000          # a heavy tie on one splitter
    k[rng.random(n) < 0.05] = 0xFFFFFFFF         # ties on the largest key
    k[rng.random(n) < 0.01] = 5
    splitters = [5, 1000, 0x40000000, 0x80000000, 0xC0000000, 0xFFFFFFF0, 0xFFFFFFFF]
    u = np.array(splitters, dtype=np.uint64)
    k64 = k.astype(np.uint64)
    cls = 2 * (k64[:, None] > u[None, :]).sum(axis=1) + (k64[:, None] == u[None, :]).any(axis=1)
    sizes = np.bincount(cls, minlength=2 * len(splitters) + 1)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    tie = 5                                       # class of keys == 0x40000000
    cuts = sorted({0, int(starts[2]), int(starts[tie] + sizes[tie] // 3), int(starts[tie] + 2 * sizes[tie] // 3),
                   int(starts[8]), int(starts[11]), int(starts[13] + sizes[13] // 2)})
    first_pos = cuts[:5] + [cuts[4]] + cuts[5:] + [n]      # destination 4 is empty
    assert len(first_pos) == 9 and first_pos == sorted(first_pos)
    gap = 64                                      # guard words between the regions: nothing may be written there
    region_off = [fp + gap * (j + 1) for j, fp in enumerate(first_pos[:8])]
    out = torch.full((n + gap * 10,), 0x5A5A5A5A, dtype=torch.int32, device="cuda")
    dest_ptrs = [out.data_ptr() + 4 * off for off in region_off]
    cuda_b.partition_scatter(_dev(k), n, torch.tensor(splitters, dtype=torch.int64, device="cuda"),
                             torch.tensor(starts[:-1], dtype=torch.int64, device="cuda"), dest_ptrs, first_pos)
    torch.cuda.synchronize()
    got = out.cpu().numpy().view(np.uint32)
    ordered = np.sort(k)                          # class order == key order, so position p holds ordered[p] up to ties
    written = np.zeros(got.size, dtype=bool)
    for j in range(8):
        a, b = first_pos[j], first_pos[j + 1]
        region = got[region_off[j]: region_off[j] + (b - a)]
        written[region_off[j]: region_off[j] + (b - a)] = True
        assert np.array_equal(np.sort(region), ordered[a:b]), (n, j)
    assert np.all(got[~written] == 0x5A5A5A5A)    # guards and slack untouched


def _nccl_worker(rank, world, port, dist_name, n, q, fused=False, strategy="exact"):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from vulkan_radix_sort_b200.datagen import make_keys
        from vulkan_radix_sort_b200.dist import CudaBackend, SharedReceive, distributed_sort
        k = make_keys(dist_name, n + 13 * rank, seed=1 + rank)
        backend = CudaBackend(rank)
        shared = SharedReceive(backend, int(n * 1.1) + 1024) if fused else None
        for _ in range(2):  # twice: the receive buffers are reused across sorts
            recv, cnt, plan = distributed_sort(backend, torch.from_numpy(k.view(np.int32).copy()).cuda(), k.size,
                                               shared=shared, strategy=strategy)
            torch.cuda.synchronize()
        q.put((rank, k, recv[:cnt].cpu().numpy().view(np.uint32).copy(), plan.targets))
        if shared is not None:
            shared.close()
        backend.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("strategy", ["exact", "sampled"])
@pytest.mark.parametrize("fused", [False, True], ids=["nccl_all_to_all", "fused_peer_stores"])
@pytest.mark.parametrize("dist_name", ["uniform", "all_zero", "skewed"])
def test_distributed_sort_over_nccl(oracle, dist_name, fused, strategy):
    if strategy == "sampled" and not fused:
        pytest.skip("sampled splitters are covered with the fused exchange")
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs at least two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, dist_name, 3_000_017, q, fused, strategy)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=300) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    union = np.concatenate([r[1] for r in results])
    assert np.array_equal(np.concatenate([r[2] for r in results]), oracle.sort_keys(union))
    for rank, _, out, targets in results:
        want = targets[rank + 1] - targets[rank]
        if strategy == "exact":
            assert out.size == want
        else:
            assert abs(out.size - want) <= 0.03 * want + 2

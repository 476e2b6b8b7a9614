"""CPU tests of the multi-GPU host logic: world_size 2 / 3 / 4 over gloo, device kernels replaced
by the NumPy test double.  Checks that the concatenation over ranks equals the oracle's sort of
the union, that the partition is balanced to +-1 key for every distribution (ties split by source
rank), and that the plan is identical on all ranks."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, dist_name, n_per_rank, q, strategy="exact"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dist_numpy_backend import NumpyBackend
        from vulkan_radix_sort_b200.datagen import make_keys
        from vulkan_radix_sort_b200.dist import distributed_sort
        n = n_per_rank + (rank * 37 if dist_name != "empty_rank" else 0)  # ragged local sizes
        if dist_name == "empty_rank" and rank == 0:
            n = 0
        name = "uniform" if dist_name == "empty_rank" else dist_name
        keys_np = make_keys(name, max(n, 1), seed=1 + rank)[:n]
        keys = torch.from_numpy(keys_np.view(np.int32).copy())
        recv, cnt, plan = distributed_sort(NumpyBackend(), keys, n, strategy=strategy)
        out = recv.numpy()[:cnt].view(np.uint32).copy()
        q.put((rank, keys_np, out, plan.total, plan.targets, plan.sizes))
    finally:
        dist.destroy_process_group()


def _run(world, dist_name, n_per_rank, strategy="exact"):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dist_name, n_per_rank, q, strategy)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(results, key=lambda r: r[0])


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("dist_name", ["uniform", "bits4", "all_zero", "all_ones", "skewed", "sorted", "empty_rank"])
def test_distributed_sort_matches_oracle(oracle, world, dist_name):
    if (world == 4 and dist_name != "all_zero") or (world == 3 and dist_name in ("bits4", "all_ones", "sorted")):
        pytest.skip("covered at the other world sizes")
    results = _run(world, dist_name, 20011)
    union = np.concatenate([r[1] for r in results])
    expect = oracle.sort_keys(union)
    got = np.concatenate([r[2] for r in results])
    assert np.array_equal(got, expect)
    total, targets = results[0][3], results[0][4]
    assert total == union.size
    for rank, _, out, t, tg, sizes in results:
        assert (t, tg) == (total, targets) and sizes == results[0][5]      # identical plan everywhere
        assert out.size == targets[rank + 1] - targets[rank]               # balanced to +-1 key, any distribution
        assert abs(out.size - total / world) <= 1


@pytest.mark.parametrize("dist_name", ["uniform", "all_zero", "skewed", "empty_rank"])
def test_sampled_splitters_sort_exactly_and_balance_approximately(oracle, dist_name):
    world = 3
    results = _run(world, dist_name, 20011, strategy="sampled")
    union = np.concatenate([r[1] for r in results])
    got = np.concatenate([r[2] for r in results])
    assert np.array_equal(got, oracle.sort_keys(union))                  # exactly sorted whatever the splitters
    total = results[0][3]
    for rank, _, out, t, tg, sizes in results:
        assert sizes == results[0][5]
        assert abs(out.size - total / world) <= 0.03 * total / world + 2   # ~1 % sampling error; ties are cut exactly

"""Multi-GPU keys-only sort: one process per GPU, torch.distributed for the plumbing.

BASELINE.json configs[4] / SURVEY.md §8(e).  The reference is single-device; this layer is new:

  1. splitters.  "sampled" (what bench.py --gpus N runs): every rank contributes 2^16 keys taken with one
     common stride, the gathered sample is sorted on the device (by vrdxCmdSort itself) and gives the
     splitter values; ONE exact class-count pass (vrdxDistCmdClassCount, 4 B/key) gives every (rank,
     class) size; balance ~1 %.  "exact": a 3-level (12/10/10-bit) distributed histogram search finds the
     key value at every global rank k*N/G (three 4 B/key passes, three all-reduces, three host round
     trips); every rank then receives N/G +- 1 keys for ANY distribution.  In both, ties on a
     splitter value (an all-equal input is one big tie) are cut exactly and handed out by source
     rank, which is legal for a keys-only sort because equal keys are indistinguishable;
  2. local multi-split (vrdxDistCmdPartition) into <= 2G-1 classes (open intervals between
     splitters and one tie class per distinct splitter value), class-ordered, so the keys bound
     for each destination rank are ONE contiguous slice of the output;
  3. exchange.  Default on GPUs: FUSED into the multi-split — vrdxDistCmdPartitionScatter stores
     every key straight into its destination rank's receive buffer (peer memory mapped over CUDA
     IPC), so the data crosses NVLink / NVSwitch while the kernel is still ranking other tiles and
     no separate collective runs; one tiny all-reduce acts as the barrier before the local sort.
     Baseline / CPU path: an all-to-all-v of the contiguous slices (NCCL grouped send/recv, gloo);
  4. local LSD sort of what arrived (vrdxCmdSort through the C-ABI).

Afterwards rank r holds the keys of global ranks [T_r, T_{r+1}) in ascending order, so the
concatenation over ranks is the sorted input — bit-identical to a single-device sort.

The numerical work is behind a small ``Backend`` object: ``CudaBackend`` calls the C-ABI; the
tests supply a NumPy stand-in so that this host logic runs under gloo on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

def search_levels(boundaries: int):
    """(shift, bits) of each level of the splitter search: 12/10/10 bits.  The first level has a
    single (empty) prefix, the later ones one prefix per boundary; 1024 bins x <= 11 prefixes stay
    under 48 KB of shared memory, so the histogram kernel keeps 4 CTAs per SM."""
    return ((20, 12), (10, 10), (0, 10))


def _u32_as_i32(t: torch.Tensor) -> torch.Tensor:
    """int64 tensor of values in [0, 2^32) -> int32 tensor with the same low 32 bits."""
    t = t.to(torch.int64)
    return torch.where(t >= (1 << 31), t - (1 << 32), t).to(torch.int32).contiguous()


# ------------------------------------------------------------------------------ backends

class CudaBackend:
    """The product path: every numerical step is a kernel of libvrdx_b200.so."""

    def __init__(self, device: int):
        from . import api
        from .sorter import Sorter
        self.api = api
        self.sorter = Sorter(device)
        self.device = self.sorter.device

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def prefix_histogram(self, keys: torch.Tensor, count: int, shift: int, bits: int, prefixes: torch.Tensor) -> torch.Tensor:
        """-> int64 [P, 2^bits]: local digit histogram at `shift` of the keys matching each prefix."""
        p = int(prefixes.numel())
        hist = torch.zeros(p << bits, dtype=torch.int32, device=self.device)
        pref32 = _u32_as_i32(prefixes)
        self.api.load_library().vrdxDistCmdPrefixHistogram(
            self._stream(), self.sorter.handle, count, keys.data_ptr(), 0, shift, bits, p, pref32.data_ptr(), 0,
            hist.data_ptr(), 0)
        self.sorter.check()
        return (hist.to(torch.int64) & 0xFFFFFFFF).view(p, 1 << bits)

    def class_count(self, keys: torch.Tensor, count: int, splitters: torch.Tensor) -> torch.Tensor:
        """-> int64 [2m+1]: local number of keys in every class of the given distinct splitters."""
        m = int(splitters.numel())
        spl = _u32_as_i32(splitters) if m else None
        counts = torch.zeros(2 * m + 1, dtype=torch.int32, device=self.device)
        self.api.load_library().vrdxDistCmdClassCount(
            self._stream(), self.sorter.handle, count, keys.data_ptr(), 0, m, spl.data_ptr() if m else None, 0,
            counts.data_ptr(), 0)
        self.sorter.check()
        return counts.to(torch.int64) & 0xFFFFFFFF

    def partition_scatter(self, keys: torch.Tensor, count: int, splitters: torch.Tensor, class_starts: torch.Tensor,
                          dest_ptrs, first_pos) -> None:
        """Fused partition + exchange: keys go straight into the peers' receive buffers."""
        m = int(splitters.numel())
        spl = _u32_as_i32(splitters) if m else None
        cursors = _u32_as_i32(class_starts)
        g = len(dest_ptrs)
        table = torch.empty(8 * g + 4 * (g + 1), dtype=torch.uint8)
        table[:8 * g] = torch.tensor([int(x) for x in dest_ptrs], dtype=torch.int64).view(torch.uint8)
        table[8 * g:] = torch.tensor([int(x) for x in first_pos], dtype=torch.int64).to(torch.int32).view(torch.uint8)
        table_d = table.to(self.device)
        self.api.load_library().vrdxDistCmdPartitionScatter(
            self._stream(), self.sorter.handle, count, keys.data_ptr(), 0, m, spl.data_ptr() if m else None, 0,
            cursors.data_ptr(), 0, g, table_d.data_ptr(), 0)
        self.sorter.check()
        self._keepalive = (spl, cursors, table_d)

    def partition(self, keys: torch.Tensor, count: int, splitters: torch.Tensor, class_starts: torch.Tensor,
                  out: torch.Tensor) -> None:
        m = int(splitters.numel())
        spl = _u32_as_i32(splitters) if m else None
        cursors = _u32_as_i32(class_starts)
        self.api.load_library().vrdxDistCmdPartition(
            self._stream(), self.sorter.handle, count, keys.data_ptr(), 0, m, spl.data_ptr() if m else None, 0,
            cursors.data_ptr(), 0, out.data_ptr(), 0)
        self.sorter.check()

    def local_sort(self, keys: torch.Tensor, count: int, storage: torch.Tensor | None = None) -> None:
        self.sorter.sort(keys, count=count, storage=storage)

    def storage_for(self, max_count: int) -> torch.Tensor:
        return self.sorter.storage_for(max_count, False)

    def close(self):
        self.sorter.close()


class _RawCudaArray:
    """__cuda_array_interface__ view of raw device memory, so torch can wrap it without owning it."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i4", "data": (ptr, False), "version": 3}


class SharedReceive:
    """This rank's receive buffer, allocated so that every other rank on the node can map it
    (cudaMalloc + CUDA IPC), plus the mapped addresses of all peers' buffers."""

    def __init__(self, backend: CudaBackend, capacity: int, group=None):
        import ctypes
        self.api = backend.api
        self.lib = backend.api.load_library()
        self.device = backend.device
        self.capacity = int(capacity)
        self.group = group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = backend.api.cuda_device(self.device.index)
        self._dev = dev
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        res = self.lib.vrdxDistAllocShared(dev, 4 * self.capacity, ctypes.byref(own), handle)
        if res != backend.api.VK_SUCCESS:
            raise RuntimeError(f"vrdxDistAllocShared failed: VkResult {res}")
        self.own_ptr = own.value
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(self.device)
        everyone = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(everyone, mine, group=group)
        self.peer_ptrs, self._opened = [], []
        for r, h in enumerate(everyone):
            if r == rank:
                self.peer_ptrs.append(self.own_ptr)
                continue
            p = ctypes.c_void_p()
            res = self.lib.vrdxDistOpenShared(dev, bytes(h.cpu().numpy().tobytes()), ctypes.byref(p))
            if res != backend.api.VK_SUCCESS:
                raise RuntimeError(f"vrdxDistOpenShared(rank {r}) failed: VkResult {res}")
            self.peer_ptrs.append(p.value)
            self._opened.append(p.value)
        self.tensor = torch.as_tensor(_RawCudaArray(self.own_ptr, self.capacity), device=self.device)
        caps = [torch.empty(1, dtype=torch.int64, device=self.device) for _ in range(world)]
        dist.all_gather(caps, torch.tensor([self.capacity], dtype=torch.int64, device=self.device), group=group)
        self.capacities = [int(c.item()) for c in caps]

    def close(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody may still be storing into a buffer that is about to be freed
        for p in self._opened:
            self.lib.vrdxDistCloseShared(self._dev, p)
        self._opened = []
        if self.own_ptr:
            self.tensor = None
            self.lib.vrdxDistFreeShared(self._dev, self.own_ptr)
            self.own_ptr = None


# ------------------------------------------------------------------------------ host logic

@dataclass
class SplitPlan:
    """Everything every rank needs to know about the exchange (identical on all ranks)."""
    total: int                    # N = sum of local counts
    targets: list                 # T_k, k = 0..G
    splitters: list               # v_k for k = 1..G-1 (may repeat)
    distinct: list                # ascending distinct splitter values u_i
    sizes: list                   # sizes[s][j] = keys rank s sends to rank j
    class_starts: list            # this rank's first output slot per class (2m+1 entries)


def _as_u32_tensor(values, device):
    return torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)


def find_splitters(backend, keys: torch.Tensor, count: int, group=None):
    """The multi-level search (search_levels).  Returns, per boundary k = 1..G-1, Python lists:
    value v_k, global #keys < v_k, global #keys == v_k, local #keys < v_k, local #keys == v_k;
    plus N and the targets T_k.  The histograms are reduced across ranks on the device
    (all-reduce); the few hundred bins of arithmetic per level run on the host."""
    import numpy as np
    world = dist.get_world_size(group)
    device = keys.device
    n_total_t = torch.tensor([count], dtype=torch.int64, device=device)
    dist.all_reduce(n_total_t, group=group)
    total = int(n_total_t.item())
    targets = [k * total // world for k in range(world + 1)]
    nb = world - 1
    if nb == 0 or total == 0:
        return [], [], [], [], [], total, targets
    remaining = np.array(targets[1:world], dtype=np.int64)   # rank of the wanted key inside the current bucket
    prefix = np.zeros(nb, dtype=np.int64)
    less_g = np.zeros(nb, dtype=np.int64)
    less_l = np.zeros(nb, dtype=np.int64)
    eq_g = eq_l = None
    rows = np.arange(nb)
    for level, (shift, bits) in enumerate(search_levels(nb)):
        p = 1 if level == 0 else nb
        h_local_t = backend.prefix_histogram(keys, count, shift, bits, torch.from_numpy(prefix[:p]).to(device))
        both = torch.stack([h_local_t, h_local_t]).contiguous()          # [0] stays local, [1] is reduced
        dist.all_reduce(both[1], group=group)
        both_h = both.cpu().numpy()                                       # one small D2H per level
        h_local = np.broadcast_to(both_h[0], (nb, 1 << bits)) if level == 0 else both_h[0]
        h_global = np.broadcast_to(both_h[1], (nb, 1 << bits)) if level == 0 else both_h[1]
        cum = np.cumsum(h_global, axis=1)                                 # inclusive
        digit = np.minimum((cum <= remaining[:, None]).sum(axis=1), (1 << bits) - 1)
        prev = np.maximum(digit - 1, 0)
        excl_g = np.where(digit > 0, cum[rows, prev], 0)
        excl_l = np.where(digit > 0, np.cumsum(h_local, axis=1)[rows, prev], 0)
        less_g += excl_g
        less_l += excl_l
        remaining = remaining - excl_g
        prefix = prefix * (1 << bits) + digit
        eq_g = h_global[rows, digit]
        eq_l = h_local[rows, digit]
    as_list = lambda a: [int(x) for x in a]
    return as_list(prefix), as_list(less_g), as_list(eq_g), as_list(less_l), as_list(eq_l), total, targets


SAMPLES_PER_RANK = 1 << 16


def _plan_from_class_stats(world, rank, total, targets, values, less_g, eq_g, counts, less_l_all, eq_l_all):
    """Common tail of both strategies.  `values[k]` is the splitter of boundary k+1; less/eq are the
    numbers of keys below / equal to it (globally, and per rank).  Boundary k+1 is placed at global
    rank clamp(T_{k+1}, less_g, less_g + eq_g): inside a tie class the cut is exact, and ties are
    handed out to the left side by source rank order."""
    nb = world - 1
    sizes = []
    for s in range(world):
        a = [0]
        for k in range(nb):
            need_left = min(max(targets[k + 1] - less_g[k], 0), eq_g[k])   # ties that end up left of boundary k+1
            before = sum(eq_l_all[t][k] for t in range(s))                  # ties held by lower ranks go first
            left = min(max(need_left - before, 0), eq_l_all[s][k])
            a.append(less_l_all[s][k] + left)
        a.append(counts[s])
        for k in range(1, len(a)):                                          # repeated splitter values keep positions monotone
            a[k] = max(a[k], a[k - 1])
        sizes.append([a[j + 1] - a[j] for j in range(world)])
    distinct = sorted(set(values))
    starts = [0]
    for u in distinct:                                                      # [interior_0][tie_0][interior_1][tie_1]...[interior_m]
        k = values.index(u)
        starts.append(less_l_all[rank][k])
        starts.append(less_l_all[rank][k] + eq_l_all[rank][k])
    return SplitPlan(total, targets, list(values), distinct, sizes, starts[:2 * len(distinct) + 1])


def make_plan_exact(backend, keys: torch.Tensor, count: int, group=None) -> SplitPlan:
    """Exact quantile splitters (multi-level histogram search): every rank receives N/G +- 1 keys."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    v, less_g, eq_g, less_l, eq_l, total, targets = find_splitters(backend, keys, count, group)
    nb = world - 1
    mine = torch.tensor([count] + less_l + eq_l, dtype=torch.int64, device=keys.device)
    everyone = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(everyone, mine, group=group)
    stats = torch.stack(everyone).cpu().tolist()
    counts = [int(r[0]) for r in stats]
    less_l_all = [[int(x) for x in r[1:1 + nb]] for r in stats]
    eq_l_all = [[int(x) for x in r[1 + nb:1 + 2 * nb]] for r in stats]
    return _plan_from_class_stats(world, rank, total, targets, v, less_g, eq_g, counts, less_l_all, eq_l_all)


def _gather_equal(t: torch.Tensor, world: int, group) -> torch.Tensor:
    """all-gather of equally sized 1-D tensors -> [world, len(t)]."""
    out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
    if t.device.type == "cuda":
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    else:
        dist.all_gather(list(out.view(world, -1).unbind(0)), t.contiguous(), group=group)
    return out.view(world, t.numel())


def make_plan_sampled(backend, keys: torch.Tensor, count: int, group=None) -> SplitPlan:
    """Splitters from a sorted sample, then ONE exact class-count pass over the local keys (the default on
    GPUs: one 4 B/key read and two host round trips instead of three of each).

    Every rank samples its keys with the SAME stride (ceil(largest local count / SAMPLES_PER_RANK), found
    with one all-reduce), so every sample stands for the same number of keys whatever the local counts
    are; the gathered samples are sorted on the device by the library's own sort and the splitter of
    boundary k is the sample at rank k * M / G.  The class counts are exact, so the result is exactly
    sorted and tie classes are still cut exactly; only the balance is approximate (~1 % for uniform keys)."""
    import numpy as np
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    device = keys.device
    nb = world - 1
    s_n = SAMPLES_PER_RANK
    cnt_t = torch.tensor([count], dtype=torch.int64, device=device)
    max_t = cnt_t.clone()
    dist.all_reduce(max_t, op=dist.ReduceOp.MAX, group=group)
    stride_t = torch.clamp((max_t + (s_n - 1)) // s_n, min=1)                  # device-side: no host sync here
    idx = torch.arange(s_n, dtype=torch.int64, device=device) * stride_t
    valid = idx < cnt_t
    if count > 0:
        picked = keys[:count][torch.clamp(idx, max=count - 1)]
        sample = torch.where(valid, picked, torch.full_like(picked, -1))       # pads = 0xFFFFFFFF: they sort last
    else:
        sample = torch.full((s_n,), -1, dtype=keys.dtype, device=device)
    # message: samples | number of valid samples | local count (two 31-bit halves)
    tail = torch.stack([valid.sum(), cnt_t[0] & 0x7FFFFFFF, cnt_t[0] >> 31]).to(keys.dtype)
    pool = _gather_equal(torch.cat([sample, tail]), world, group)                # [world, s_n + 3]
    samples = pool[:, :s_n].contiguous().view(-1)
    backend.local_sort(samples, world * s_n, backend.storage_for(world * s_n))  # ascending as unsigned 32-bit
    m_total = pool[:, s_n].to(torch.int64).sum()
    qidx = torch.clamp((torch.arange(1, world, dtype=torch.int64, device=device) * m_total) // world, max=world * s_n - 1)
    picked_q = samples[qidx] if nb else samples[:0]
    host = torch.cat([pool[:, s_n:].reshape(-1), picked_q]).cpu().tolist()       # host round trip 1
    counts = [int(host[3 * r + 1]) + (int(host[3 * r + 2]) << 31) for r in range(world)]
    total = sum(counts)
    targets = [k * total // world for k in range(world + 1)]
    if nb == 0 or total == 0:
        return SplitPlan(total, targets, [], [], [[counts[s] if j == 0 else 0 for j in range(world)] for s in range(world)], [0])
    values = [int(x) & 0xFFFFFFFF for x in host[3 * world:]]
    distinct = sorted(set(values))
    cls = backend.class_count(keys, count, _as_u32_tensor(distinct, device))
    cls_all = _gather_equal(cls, world, group).cpu().numpy()                      # host round trip 2: [world, 2m+1]
    excl = np.concatenate([np.zeros((world, 1), dtype=np.int64), np.cumsum(cls_all, axis=1)[:, :-1]], axis=1)
    less_l_all, eq_l_all = [], []
    for s in range(world):
        less_l_all.append([int(excl[s, 2 * distinct.index(v) + 1]) for v in values])
        eq_l_all.append([int(cls_all[s, 2 * distinct.index(v) + 1]) for v in values])
    less_g = [sum(less_l_all[s][k] for s in range(world)) for k in range(nb)]
    eq_g = [sum(eq_l_all[s][k] for s in range(world)) for k in range(nb)]
    return _plan_from_class_stats(world, rank, total, targets, values, less_g, eq_g, counts, less_l_all, eq_l_all)


def make_plan(backend, keys: torch.Tensor, count: int, group=None, strategy: str = "exact") -> SplitPlan:
    if strategy == "exact" or not hasattr(backend, "class_count"):
        return make_plan_exact(backend, keys, count, group)
    return make_plan_sampled(backend, keys, count, group)


def distributed_sort(backend, keys: torch.Tensor, count: int | None = None, group=None, recv: torch.Tensor | None = None,
                     part: torch.Tensor | None = None, storage: torch.Tensor | None = None, timers=None,
                     shared: "SharedReceive | None" = None, strategy: str = "exact"):
    """Sort the union of every rank's keys[0:count].  Returns (recv_buffer, recv_count, plan): rank
    r's slice of the globally sorted sequence (global ranks [T_r, T_{r+1})).

    With ``shared`` (a SharedReceive) the exchange is fused into the partition kernel (peer stores
    over NVLink); otherwise the keys are multi-split locally and exchanged with an all-to-all-v."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = int(keys.numel() if count is None else count)
    device = keys.device

    def mark(name):
        if timers is not None:
            timers.mark(name)

    mark("start")
    plan = make_plan(backend, keys, n, group, strategy)
    mark("splitters")
    in_splits = plan.sizes[rank]
    out_splits = [plan.sizes[s][rank] for s in range(world)]
    recv_count = sum(out_splits)
    splitters_t = _as_u32_tensor(plan.distinct, device)
    starts_t = _as_u32_tensor(plan.class_starts, device)
    if shared is not None:
        # the plan is identical on every rank, so every rank checks EVERY destination and they all raise
        # together: nobody stores past the end of a peer's buffer, nobody waits in the barrier for a
        # rank that bailed out (capacities are exchanged once, when the buffers are mapped)
        for j in range(world):
            need = sum(plan.sizes[s_][j] for s_ in range(world))
            if need > shared.capacities[j]:
                raise RuntimeError(f"receive buffer of rank {j} too small: {need} > {shared.capacities[j]}")
        recv = shared.tensor
        # where my keys for destination j start inside j's buffer: after the keys of lower ranks
        first_pos = [0]
        for j in range(world):
            first_pos.append(first_pos[-1] + in_splits[j])
        dest_ptrs = [shared.peer_ptrs[j] + 4 * sum(plan.sizes[s][j] for s in range(rank)) for j in range(world)]
        if n:
            backend.partition_scatter(keys, n, splitters_t, starts_t, dest_ptrs, first_pos)
        mark("partition")
        token = torch.zeros(1, dtype=torch.int32, device=device)
        dist.all_reduce(token, group=group)      # barrier on the stream: every rank's stores have completed
        mark("exchange")
    else:
        if part is None:
            part = torch.empty(max(n, 1), dtype=keys.dtype, device=device)
        if recv is None or recv.numel() < recv_count:
            recv = torch.empty(max(recv_count, 1), dtype=keys.dtype, device=device)
        if n:
            backend.partition(keys, n, splitters_t, starts_t, part)
        mark("partition")
        if world > 1:
            dist.all_to_all_single(recv[:recv_count], part[:n], out_splits, in_splits, group=group)
        else:
            recv[:recv_count].copy_(part[:n])
        mark("exchange")
    if recv_count:
        backend.local_sort(recv, recv_count, storage)
    mark("local_sort")
    return recv, recv_count, plan

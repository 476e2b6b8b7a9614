"""Timing of vrdxCudaCmdSortKeys64 (64-bit keys, two chained key-value sorts) on resident data."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vulkan_radix_sort_b200 import Sorter, api

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
n = 1 << log2n
g = torch.Generator(device="cuda").manual_seed(1)
src = torch.randint(-2**63, 2**63 - 1, (n,), generator=g, dtype=torch.int64, device="cuda")
s = Sorter(0)
size = api.vrdxCudaGetSorterKeys64StorageRequirements(s.handle, n).size
st = torch.empty(int(size), dtype=torch.uint8, device="cuda")
for name, kw in (("uint64 asc", dict(key_type=0)), ("int64 desc", dict(key_type=1, descending=True))):
    ms = []
    for it in range(5):
        k = src.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.sort_keys64(k, storage=st, **kw); e1.record(); torch.cuda.synchronize()
        if it >= 2: ms.append(e0.elapsed_time(e1))
    t = sum(ms) / len(ms)
    ok = bool((k[1:] >= k[:-1]).all()) if "asc" in name and kw["key_type"] == 1 else None
    print(f"N=2^{log2n} keys64 {name:12s} {t:7.3f} ms  {n / t / 1e6:6.2f} GKeys/s  launches={s.last_launch_count}", flush=True)

"""Timing of vrdxCudaCmdSortEx variants (key types, descending, bit sub-ranges) on resident data."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vulkan_radix_sort_b200 import Sorter

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
n = 1 << log2n
g = torch.Generator(device="cuda").manual_seed(1)
bits = torch.randint(-2**31, 2**31 - 1, (n,), generator=g, dtype=torch.int64, device="cuda").to(torch.int32)
flt = torch.randn(n, generator=g, device="cuda")
vals = torch.arange(n, dtype=torch.int32, device="cuda")
s = Sorter(0)
cases = [("uint32 asc [0,32)", bits, dict(key_type=0)), ("int32 asc", bits, dict(key_type=1)),
         ("float32 asc", flt, dict()), ("float32 desc", flt, dict(descending=True)),
         ("uint32 bits [0,16)", bits, dict(key_type=0, begin_bit=0, end_bit=16)),
         ("uint32 bits [8,32)", bits, dict(key_type=0, begin_bit=8, end_bit=32)),
         ("uint32 bits [0,8)", bits, dict(key_type=0, begin_bit=0, end_bit=8))]
for kv in (False, True):
    st = s.storage_for(n, kv)
    for name, src, kw in cases:
        ms = []
        for it in range(6):
            k = src.clone(); v = vals.clone() if kv else None
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.sort_ex(k, v, storage=st, **kw); e1.record(); torch.cuda.synchronize()
            if it >= 2: ms.append(e0.elapsed_time(e1))
        t = sum(ms) / len(ms)
        print(f"N=2^{log2n} {'kv  ' if kv else 'keys'} {name:22s} {t:7.3f} ms  {n / t / 1e6:7.2f} GKeys/s  launches={s.last_launch_count}", flush=True)

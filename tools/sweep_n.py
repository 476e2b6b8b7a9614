"""Developer sweep over N for both algorithms (crossover for AUTO)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vulkan_radix_sort_b200 import Sorter, api
from vulkan_radix_sort_b200.datagen import make_keys
res, pool = api.vrdxCudaCreateQueryPool(api.cuda_device(0), 15)
sorters = {"onesweep": Sorter(0, algorithm=1), "rts": Sorter(0, algorithm=2)}
host = torch.from_numpy(make_keys("uniform", 1 << 28, 1).view(np.int32))
src_all = host.cuda()
vals_all = torch.arange(1 << 28, dtype=torch.int32, device="cuda")
import math
sizes = [1 << 18, 1 << 20, 1 << 21, 3 << 20, 1 << 22, 3 << 21, 1 << 23, 3 << 22, 1 << 24, 3 << 23, 1 << 25, 1 << 26]
if len(sys.argv) > 1:
    sizes = [int(eval(a)) for a in sys.argv[1:]]
for n in sizes:
    line = f"{n:9d} (2^{math.log2(n):.2f})"
    for kv in (False, True):
        for name, s in sorters.items():
            keys = torch.empty(n, dtype=torch.int32, device="cuda"); vals = torch.empty_like(keys)
            st = s.storage_for(n, kv); ts = []
            for r in range(7):
                keys.copy_(src_all[:n]); vals.copy_(vals_all[:n]); torch.cuda.synchronize()
                (s.sort_key_value(keys, vals, storage=st, query_pool=pool) if kv else s.sort(keys, storage=st, query_pool=pool))
                torch.cuda.synchronize()
                rc, t = api.vrdxCudaGetQueryPoolResults(pool)
                if r >= 2: ts.append(t[14] - t[0])
            line += f"  {'kv' if kv else 'k'}:{name}={n/np.median(ts):6.2f}"
    print(line, flush=True)

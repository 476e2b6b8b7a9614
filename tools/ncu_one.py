"""Tiny driver for ncu captures: a few sorts of resident data (keys-only or kv)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vulkan_radix_sort_b200 import Sorter
from vulkan_radix_sort_b200.datagen import make_keys
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
kv = len(sys.argv) > 2 and sys.argv[2] == "kv"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
n = 1 << log2n
src = torch.from_numpy(make_keys("uniform", n, 1).view(np.int32)).cuda()
s = Sorter(0)
for _ in range(reps):
    k = src.clone()
    if kv:
        v = torch.arange(n, dtype=torch.int32, device="cuda")
        s.sort_key_value(k, v)
    else:
        s.sort(k)
    torch.cuda.synchronize()
print("done")

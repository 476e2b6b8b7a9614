"""Developer probe: time the multi-GPU building-block kernels on ONE GPU with the shapes an
8-GPU run produces (7 prefixes / 7 splitters, 2^29 keys)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vulkan_radix_sort_b200.dist import CudaBackend
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 29
n = 1 << log2n
g = torch.Generator(device="cuda"); g.manual_seed(1)
keys = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device="cuda", generator=g)
b = CudaBackend(0)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
for world in (2, 8):
    nb = world - 1
    bounds = [(k + 1) * (1 << 32) // world for k in range(nb)]
    for shift, bits in ((24, 8), (12, 12), (0, 12), (20, 12), (10, 10), (0, 10)):
        pref = torch.tensor([x >> (shift + bits) for x in bounds][: (1 if shift + bits == 32 else nb)], dtype=torch.int64, device="cuda")
        t = timeit(lambda: b.prefix_histogram(keys, n, shift, bits, pref))
        print(f"world={world} prefix_histogram shift={shift} bits={bits} P={pref.numel()}: {t:.3f} ms  {4*n/t/1e6:.0f} GB/s")
    spl = torch.tensor(bounds, dtype=torch.int64, device="cuda")
    t = timeit(lambda: b.class_count(keys, n, spl))
    print(f"world={world} class_count m={nb}: {t:.3f} ms  {4*n/t/1e6:.0f} GB/s")
    k64 = keys.to(torch.int64) & 0xFFFFFFFF
    less = [(k64 < x).sum().item() for x in bounds]; eq = [(k64 == x).sum().item() for x in bounds]
    starts = [0]
    for l, e in zip(less, eq): starts += [l, l + e]
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    st = torch.tensor(starts, dtype=torch.int64, device="cuda")
    t = timeit(lambda: b.partition(keys, n, spl, st, out))
    print(f"world={world} partition m={nb}: {t:.3f} ms  {n/t/1e6:.1f} GKeys/s")

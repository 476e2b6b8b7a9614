"""Randomised parity soak (developer tool): random sizes, distributions, algorithms and kinds against
torch.sort(stable=True) on the device (keys AND payload).  usage: python tools/fuzz_parity.py [seconds] [seed] [min log2 n]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vulkan_radix_sort_b200 import Sorter

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
min_e = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
rng = np.random.default_rng(seed)
g = torch.Generator(device="cuda").manual_seed(seed)
sorters = {a: Sorter(0, algorithm=a) for a in (0, 1, 2)}
t0 = time.time()
runs = fails = 0
kinds = {}


def make(n, dist):
    if dist == "uniform":
        return torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device="cuda", generator=g)
    if dist in ("low16", "low24", "low8"):
        bits = {"low8": 8, "low16": 16, "low24": 24}[dist]
        m = int(rng.integers(1, 60))
        lows = torch.randint(0, 1 << bits, (m,), dtype=torch.int64, device="cuda", generator=g)
        pick = lows[torch.randint(0, m, (n,), device="cuda", generator=g)]
        hi = torch.randint(0, 1 << (32 - bits), (n,), dtype=torch.int64, device="cuda", generator=g) << bits
        return ((hi | pick) & 0xFFFFFFFF).to(torch.int64).sub_((1 << 32) * ((hi | pick) >> 31)).to(torch.int32)
    if dist == "fewbits":
        b = int(rng.integers(1, 20))
        return torch.randint(0, 1 << b, (n,), dtype=torch.int32, device="cuda", generator=g)
    if dist == "sorted":
        return torch.sort(torch.randint(0, (1 << 31) - 1, (n,), dtype=torch.int32, device="cuda", generator=g)).values
    if dist == "const":
        return torch.full((n,), int(rng.integers(0, 1 << 31)), dtype=torch.int32, device="cuda")
    raise ValueError(dist)


while time.time() - t0 < budget:
    e = rng.uniform(min_e, 27.2)
    n = max(1, int(2 ** e) + int(rng.integers(-3, 4)))
    dist = rng.choice(["uniform", "uniform", "low16", "low24", "low8", "fewbits", "sorted", "const"])
    algo = int(rng.choice([0, 0, 1, 2, 2]))
    if algo == 1 and n >= (1 << 27):
        algo = 0
    kv = bool(rng.integers(0, 2))
    k = make(n, dist)
    u = k.view(torch.uint32).to(torch.int64)
    ref_k, ref_i = torch.sort(u, stable=True)
    keys = k.clone()
    s = sorters[algo]
    if kv:
        vals = torch.arange(n, dtype=torch.int32, device="cuda")
        s.sort_key_value(keys, vals)
    else:
        s.sort(keys)
    torch.cuda.synchronize()
    ok = bool((keys.view(torch.uint32).to(torch.int64) == ref_k).all())
    if kv:
        ok = ok and bool((vals.to(torch.int64) == ref_i).all())
    runs += 1
    kinds[(dist, algo, kv)] = kinds.get((dist, algo, kv), 0) + 1
    if not ok:
        fails += 1
        print(f"MISMATCH n={n} dist={dist} algo={algo} kv={kv}", flush=True)
    del k, u, ref_k, ref_i, keys
print(f"fuzz: {runs} sorts in {time.time() - t0:.0f} s, {fails} mismatches, {len(kinds)} (distribution, algorithm, kind) combinations, seed {seed}")
sys.exit(1 if fails else 0)

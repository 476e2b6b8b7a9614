"""Developer timing probe (not the contract bench): sorts resident data through the C-ABI and
prints GKeys/s plus the per-stage split from the 15-slot query pool.

    python tools/quick_bench.py [--log2n 25 28] [--reps 5] [--kv] [--dist uniform]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vulkan_radix_sort_b200 import Sorter, api  # noqa: E402
from vulkan_radix_sort_b200.datagen import make_keys  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, nargs="+", default=[20, 22, 25, 28])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--dist", default="uniform")
    ap.add_argument("--algorithm", type=int, default=0)
    ap.add_argument("--tile-load", type=int, default=0)
    args = ap.parse_args()

    s = Sorter(0, algorithm=args.algorithm, tile_load=args.tile_load)
    p = s.properties
    print(f"device sm={p.smCount} cc={p.ccMajor}.{p.ccMinor} tile keys={p.keysTileSize} kv={p.keyValueTileSize}")
    res, pool = api.vrdxCudaCreateQueryPool(api.cuda_device(0), 15)
    for log2n in args.log2n:
        n = 1 << log2n
        host = torch.from_numpy(make_keys(args.dist, n, 1).view(np.int32))
        src = host.cuda()
        vals_src = torch.arange(n, dtype=torch.int32, device="cuda")
        for kv in (False, True):
            keys = torch.empty_like(src)
            vals = torch.empty_like(src) if kv else None
            storage = s.storage_for(n, kv)
            times, stages = [], None
            for r in range(args.reps + 2):
                keys.copy_(src)
                if kv:
                    vals.copy_(vals_src)
                torch.cuda.synchronize()
                if kv:
                    s.sort_key_value(keys, vals, storage=storage, query_pool=pool)
                else:
                    s.sort(keys, storage=storage, query_pool=pool)
                torch.cuda.synchronize()
                rc, ts = api.vrdxCudaGetQueryPoolResults(pool)
                if r >= 2:
                    times.append(ts[14] - ts[0])
                    stages = ts
            t = float(np.median(times))
            bpk = 68 if kv else 36
            hist = (stages[1] - stages[0]) / 1e3
            passes = [(stages[4 + 3 * i] - stages[3 + 3 * i]) / 1e3 for i in range(4)]
            ups = [(stages[2 + 3 * i] - stages[1 + 3 * i]) / 1e3 for i in range(4)]
            sps = [(stages[3 + 3 * i] - stages[2 + 3 * i]) / 1e3 for i in range(4)]
            print(f"N=2^{log2n} {'kv  ' if kv else 'keys'} {t/1e6:8.3f} ms  {n/t:7.2f} GKeys/s  "
                  f"{bpk*n/t:8.1f} GB/s  hist+reset={hist:.1f}us up={[round(x,1) for x in ups]} sp={[round(x,1) for x in sps]} dn={[round(x,1) for x in passes]}us",
                  flush=True)
            ok = bool((keys[1:].view(torch.uint32).to(torch.int64) >= keys[:-1].view(torch.uint32).to(torch.int64)).all()) \
                if n <= (1 << 26) else True
            if not ok:
                print("   !!! output not sorted")
    api.vrdxCudaDestroyQueryPool(pool)


if __name__ == "__main__":
    main()

"""Developer A/B sweep: every tile shape of the loaded library x both algorithms x keys-only / key-value,
resident data, timed with the 15-slot query pool.  Each configuration is verified once against
torch.sort(stable=True) (keys AND payload, i.e. stability) before it is timed.

    VRDX_LIB=build/ab/libvrdx_<variant>.so python tools/shape_sweep.py [--log2n 25 28] [--shapes 0 1 2 3 4]
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
    ap.add_argument("--log2n", type=int, nargs="+", default=[25, 28])
    ap.add_argument("--shapes", type=int, nargs="+", default=[0, 1, 2, 3, 4])
    ap.add_argument("--algos", type=int, nargs="+", default=[2, 1])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--experiment", type=int, default=0)
    ap.add_argument("--dist", default="uniform")
    ap.add_argument("--pairs-only-shape", action="store_true")
    ap.add_argument("--kinds", nargs="+", default=["keys", "kv"])
    args = ap.parse_args()
    tag = os.path.basename(os.environ.get("VRDX_LIB", "product"))
    res, pool = api.vrdxCudaCreateQueryPool(api.cuda_device(0), 15)
    nmax = 1 << max(args.log2n)
    host = torch.from_numpy(make_keys(args.dist, nmax, 1).view(np.int32))
    src_all = host.cuda()
    vals_all = torch.arange(nmax, dtype=torch.int32, device="cuda")
    nv = 1 << 22
    ref_k, ref_i = torch.sort(src_all[:nv].view(torch.uint32).to(torch.int64), stable=True)
    for shape in args.shapes:
        for algo in args.algos:
            try:
                # (--pairs-only-shape: the index selects a key-value shape, the keys shape stays the default)
                s = Sorter(0, algorithm=algo, reserved=(0 if args.pairs_only_shape else shape + 1, shape + 1, args.experiment))
            except RuntimeError as e:
                print(f"[{tag}] shape {shape} algo {algo}: {e}")
                continue
            for kv in [k == "kv" for k in args.kinds]:
                # correctness first (2^22 + an odd tail), keys and payload
                for n in (nv, nv - 4099):
                    keys = src_all[:n].clone()
                    vals = vals_all[:n].clone()
                    if kv:
                        s.sort_key_value(keys, vals)
                    else:
                        s.sort(keys)
                    torch.cuda.synchronize()
                    rk, ri = (ref_k, ref_i) if n == nv else torch.sort(src_all[:n].view(torch.uint32).to(torch.int64), stable=True)
                    ok = bool((keys.view(torch.uint32).to(torch.int64) == rk).all())
                    if kv:
                        ok = ok and bool((vals.to(torch.int64) == ri).all())
                    if not ok:
                        print(f"[{tag}] shape {shape} algo {algo} kv={kv} n={n}: WRONG RESULT", flush=True)
                for log2n in args.log2n:
                    n = 1 << log2n
                    keys = torch.empty(n, dtype=torch.int32, device="cuda")
                    vals = torch.empty(n, dtype=torch.int32, device="cuda") if kv else None
                    storage = s.storage_for(n, kv)
                    times, stages = [], None
                    for r in range(args.reps + 2):
                        keys.copy_(src_all[:n])
                        if kv:
                            vals.copy_(vals_all[:n])
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
                    dn = [(stages[4 + 3 * i] - stages[3 + 3 * i]) / 1e3 for i in range(4)]
                    up = [(stages[2 + 3 * i] - stages[1 + 3 * i]) / 1e3 for i in range(4)]
                    p = s.properties
                    print(f"[{tag}] shape {shape} tile {p.keyValueTileSize if kv else p.keysTileSize} "
                          f"{'rts' if algo == 2 else 'one'} {'kv  ' if kv else 'keys'} 2^{log2n}: {t/1e6:7.3f} ms "
                          f"{n/t:6.2f} GKeys/s  up={[round(x) for x in up]} dn={[round(x) for x in dn]} us", flush=True)
            s.close()
    api.vrdxCudaDestroyQueryPool(pool)


if __name__ == "__main__":
    main()

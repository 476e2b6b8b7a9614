#!/usr/bin/env python
"""Join two CSVs of bench_cpp/bench (schema of the reference's bench/bench.cc:199) — the b200 backend and
the cuda (CUB) backend of the same sweep — into the table kept as profiles/sweep/summary.txt."""
import csv
import sys


def load(path):
    rows = {}
    for f in csv.reader(open(path)):
        if len(f) == 7 and f[1].isdigit():
            rows[(int(f[1]), f[2])] = float(f[5])
    return rows


def main(b200_csv, cuda_csv):
    a, b = load(b200_csv), load(cuda_csv)
    sizes = sorted({n for n, _ in a} & {n for n, _ in b})
    print("# reference sweep (bench/bench.cc protocol, 128 points 2^18..2^25, median of 3 runs, fresh data per run), GItems/s")
    print("# b200: times from the 15-slot GPU-timestamp query pool (ts[14]-ts[0]); cub: two CUDA events around the call")
    print("# n, b200 keys, cub keys, ratio, b200 kv, cub kv, ratio")
    kinds = sorted({k for _, k in a})
    kk = [k for k in kinds if k.startswith("keys")][0]
    kv = [k for k in kinds if k != kk][0]
    win_k = win_v = 0
    for i, n in enumerate(sizes):
        rk, rv = a[(n, kk)] / b[(n, kk)], a[(n, kv)] / b[(n, kv)]
        win_k += rk > 1.0
        win_v += rv > 1.0
        if i % 8 == 0 or i == len(sizes) - 1:
            print(f"{n:9d}  {a[(n, kk)]:6.2f} {b[(n, kk)]:6.2f}  {rk:4.2f}   {a[(n, kv)]:6.2f} {b[(n, kv)]:6.2f}  {rv:4.2f}")
    print(f"# b200 faster than CUB at {win_k}/{len(sizes)} sweep points keys-only, {win_v}/{len(sizes)} key-value; "
          f"peak b200 keys {max(a[(n, kk)] for n in sizes):.2f}, peak cub {max(b[(n, kk)] for n in sizes):.2f}; "
          f"min ratio keys {min(a[(n, kk)] / b[(n, kk)] for n in sizes):.2f}, kv {min(a[(n, kv)] / b[(n, kv)] for n in sizes):.2f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

"""Small sorts of every flavour for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vulkan_radix_sort_b200 import Sorter, api
from vulkan_radix_sort_b200.datagen import make_keys
flavours = {"onesweep": (1, 1, None), "rts": (2, 1, None), "onesweep_256x16": (1, 1, (2, 2)), "rts_512x16": (2, 1, (4, 4)),
            # VRDX_EXPERIMENTS builds only:
            "onesweep_tma": (1, 2, None), "rts_tma": (2, 2, None), "cluster4": (1, 1, (0, 0, 2)), "ranges": (2, 1, (0, 0, 7))}
only = sys.argv[1:] or list(flavours)
for name in only:
    algo, load, res = flavours[name]
    s = Sorter(0, algorithm=algo, tile_load=load, reserved=res)
    for n in (1, 777, 5120, 6400, 20011, 70001):
        for dist in ("uniform", "bits4", "sorted"):   # sorted: every warp agrees in the low bits (relaxed ranking)
            k = make_keys(dist, n, 3)
            dk = torch.from_numpy(k.view(np.int32)).cuda()
            dv = torch.arange(n, dtype=torch.int32, device="cuda")
            cnt = torch.tensor([max(n - 5, 0)], dtype=torch.int32, device="cuda")
            kk = dk.clone()
            s.sort(kk)
            assert np.array_equal(kk.cpu().numpy().view(np.uint32), np.sort(k)), (name, n, dist, 'keys')
            kk = dk.clone()
            s.sort_indirect(kk, cnt, max_count=n)
            s.sort_key_value_indirect(dk, dv, cnt, max_count=n)
            torch.cuda.synchronize()
            c = int(cnt.item())
            got = dk.cpu().numpy().view(np.uint32)
            assert np.array_equal(got[:c], np.sort(k[:c], kind="stable")), (name, n, dist)
    if algo == 2:
        # block-free tiles (keys-only reduce-then-scan): run VRDX_TWO_RUNS=1 to cover the two-run flavours too;
        # few distinct values below the second / third digit give tiles of one, two and more runs in passes 1 and 2
        rng = np.random.default_rng(5)
        for n in (20011, 70001):
            for m in (2, 3, 9):
                low = rng.choice(1 << 16, size=m, replace=False).astype(np.uint32)
                k = (low[rng.integers(0, m, n)] | (rng.integers(0, 1 << 16, n, dtype=np.uint32) << np.uint32(16))).astype(np.uint32)
                kk = torch.from_numpy(k.view(np.int32)).cuda()
                s.sort(kk)
                assert np.array_equal(kk.cpu().numpy().view(np.uint32), np.sort(k)), (name, n, m, 'block-free')
    s.close()
    print("ok", name, flush=True)

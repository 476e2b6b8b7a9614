"""Developer probe: look-back depth statistics from a -DVRDX_STATS build of the library."""
import os, sys, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vulkan_radix_sort_b200 import build
build.LIB_PATH = os.path.join(build.LIB_DIR, "libvrdx_b200_stats.so")
from vulkan_radix_sort_b200 import Sorter
from vulkan_radix_sort_b200.datagen import make_keys
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
n = 1 << log2n
src = torch.from_numpy(make_keys("uniform", n, 1).view(np.int32)).cuda()
s = Sorter(0)
tile = s.properties.keysTileSize
for rep in range(2):
    k = src.clone()
    st = s.storage_for(n, False)
    s.sort(k, storage=st)
    torch.cuda.synchronize()
    hdr = st[:4144].cpu().numpy().view(np.uint32)
    rounds, cells, notready = hdr[1033], hdr[1034], hdr[1035]   # reserved[0..2] after count[4]+hist[1024]+tickets[4]+done[1]
    tiles = 4 * ((n + tile - 1) // tile - 1)
    print(f"tiles={tiles} per-tile: rounds={rounds/tiles:.2f} cells={cells/tiles:.2f} notready={notready/tiles:.2f}")

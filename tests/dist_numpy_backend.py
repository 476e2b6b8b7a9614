"""NumPy stand-in for the device kernels of the distributed sort — TEST DOUBLE ONLY.

It lets tests/test_dist_gloo.py drive the host logic of vulkan_radix_sort_b200/dist.py (splitter
search, tie splitting, slice bookkeeping, all-to-all-v) under gloo on CPU.  Each method restates
the contract of the C-ABI entry point it replaces (include/vrdx_dist.h) and the local sort uses
the oracle (oracle/lsd_oracle.c)."""
import numpy as np
import torch

from oracle import cpu_oracle


def _u32(t: torch.Tensor, count: int) -> np.ndarray:
    return t.numpy()[:count].view(np.uint32)


class NumpyBackend:
    device = torch.device("cpu")

    def prefix_histogram(self, keys, count, shift, bits, prefixes):
        k = _u32(keys, count).astype(np.uint64)
        bins = 1 << bits
        out = np.zeros((int(prefixes.numel()), bins), dtype=np.int64)
        digit = ((k >> np.uint64(shift)) & np.uint64(bins - 1)).astype(np.int64)
        hi = k >> np.uint64(shift + bits)
        for j, p in enumerate(prefixes.tolist()):
            sel = np.ones(k.size, dtype=bool) if shift + bits >= 32 else hi == np.uint64(p)
            out[j] = np.bincount(digit[sel], minlength=bins)
        return torch.from_numpy(out)

    def class_count(self, keys, count, splitters):
        k64 = _u32(keys, count).astype(np.uint64)
        u = np.array(splitters.tolist(), dtype=np.uint64)
        gt = (k64[:, None] > u[None, :]).sum(axis=1) if u.size else np.zeros(k64.size, dtype=np.int64)
        eq = (k64[:, None] == u[None, :]).any(axis=1) if u.size else np.zeros(k64.size, dtype=bool)
        return torch.from_numpy(np.bincount(2 * gt + eq, minlength=2 * u.size + 1).astype(np.int64))

    def partition(self, keys, count, splitters, class_starts, out):
        k = _u32(keys, count)
        u = np.array(splitters.tolist(), dtype=np.uint64)
        k64 = k.astype(np.uint64)
        gt = (k64[:, None] > u[None, :]).sum(axis=1) if u.size else np.zeros(k.size, dtype=np.int64)
        eq = (k64[:, None] == u[None, :]).any(axis=1) if u.size else np.zeros(k.size, dtype=bool)
        cls = 2 * gt + eq
        order = np.argsort(cls, kind="stable")
        grouped = k[order]
        starts = class_starts.tolist()
        sizes = np.bincount(cls, minlength=len(starts))
        expect = np.concatenate([[0], np.cumsum(sizes)[:-1]])
        assert list(expect) == [int(x) for x in starts], (list(expect), starts)  # the plan's class layout is exact
        out.numpy()[:count].view(np.uint32)[:] = grouped

    def local_sort(self, keys, count, storage=None):
        view = keys.numpy()[:count].view(np.uint32)
        view[:] = cpu_oracle.sort_keys(view)

    def storage_for(self, max_count):
        return None

    def close(self):
        pass

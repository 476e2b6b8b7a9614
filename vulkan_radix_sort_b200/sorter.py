"""Host-side convenience over the C-ABI: a sorter bound to one CUDA device.

This is the Python twin of how the reference's own caller drives the API
(bench/vulkan_benchmark.cc:253-433: create sorter once, query storage requirements, allocate
storage, record vrdxCmdSort* into a command buffer, submit, wait).  PyTorch supplies device
memory and streams only; every sort goes through ``vrdxCmdSort*`` in libvrdx_b200.so.
"""
from __future__ import annotations

import torch

from . import api


def _as_i32(t: torch.Tensor) -> torch.Tensor:
    """uint32 payloads are carried in int32 tensors (same bits; torch's uint32 support is thin)."""
    if t.dtype == torch.int32:
        return t
    if t.dtype == torch.uint32:
        return t.view(torch.int32)
    raise TypeError(f"keys/values must be 32-bit integers, got {t.dtype}")


class Sorter:
    """One ``VrdxSorter`` plus grow-only storage buffers (the caller-owned scratch), one per CUDA stream
    that sorts were enqueued on: the storage holds per-sort state (header, tables, ping-pong halves), so
    two sorts in flight on different streams must not share it (the C API leaves that to the caller, this
    wrapper must not hide a race).  Each buffer is allocated under the stream it serves, which makes the
    caching allocator's reuse of a regrown buffer stream-ordered."""

    def __init__(self, device: int | torch.device = 0, algorithm: int = api.VRDX_CUDA_ALGORITHM_AUTO,
                 tile_load: int = api.VRDX_CUDA_TILE_LOAD_AUTO, reserved=None):
        if not torch.cuda.is_available():
            raise RuntimeError("vulkan_radix_sort_b200 needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        ordinal = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", ordinal)
        info = api.VrdxSorterCreateInfo(api.cuda_device(ordinal), api.cuda_device(ordinal), None)
        res, handle = api.vrdxCudaCreateSorter(info, algorithm, tile_load, reserved)
        if res != api.VK_SUCCESS:
            raise RuntimeError(f"vrdxCreateSorter failed with VkResult {res}")
        self.handle = handle
        self._storage: dict[int, torch.Tensor] = {}

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "handle", None):
            torch.cuda.synchronize(self.device)
            api.vrdxDestroySorter(self.handle)
            self.handle = None
            self._storage = {}

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ storage
    def storage_requirements(self, max_count: int, key_value: bool = False):
        fn = api.vrdxGetSorterKeyValueStorageRequirements if key_value else api.vrdxGetSorterStorageRequirements
        return fn(self.handle, max_count)

    def storage_for(self, max_count: int, key_value: bool, stream=None) -> torch.Tensor:
        return self._scratch(int(self.storage_requirements(max_count, key_value).size), stream)

    def _scratch(self, need: int, stream=None) -> torch.Tensor:
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        cur = self._storage.get(st.cuda_stream)
        if cur is None or cur.numel() < need:
            self._storage.pop(st.cuda_stream, None)
            with torch.cuda.stream(st):
                cur = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._storage[st.cuda_stream] = cur
        return cur

    # ------------------------------------------------------------------ sorts (in place)
    def _stream(self, stream):
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        return s.cuda_stream

    def check(self) -> None:
        err = api.vrdxCudaGetLastError(self.handle)
        if err:
            raise RuntimeError(f"vrdx CUDA error {err}: {api.vrdxCudaGetErrorString(err)}")

    def sort(self, keys: torch.Tensor, count: int | None = None, storage: torch.Tensor | None = None,
             stream=None, query_pool=None, query: int = 0) -> None:
        """vrdxCmdSort: sort keys[0:count] ascending as unsigned 32-bit, in place."""
        k = _as_i32(keys)
        n = k.numel() if count is None else int(count)
        st = storage if storage is not None else self.storage_for(n, False, stream)
        api.vrdxCmdSort(self._stream(stream), self.handle, n, k.data_ptr(), 0, st.data_ptr(), 0,
                        query_pool, query)
        self.check()

    def sort_key_value(self, keys: torch.Tensor, values: torch.Tensor, count: int | None = None,
                       storage: torch.Tensor | None = None, stream=None, query_pool=None,
                       query: int = 0) -> None:
        """vrdxCmdSortKeyValue: stable sort of (key, value) pairs by key, in place."""
        k, v = _as_i32(keys), _as_i32(values)
        n = k.numel() if count is None else int(count)
        st = storage if storage is not None else self.storage_for(n, True, stream)
        api.vrdxCmdSortKeyValue(self._stream(stream), self.handle, n, k.data_ptr(), 0, v.data_ptr(), 0,
                                st.data_ptr(), 0, query_pool, query)
        self.check()

    def sort_indirect(self, keys: torch.Tensor, count_buffer: torch.Tensor, max_count: int | None = None,
                      count_offset: int = 0, storage: torch.Tensor | None = None, stream=None,
                      query_pool=None, query: int = 0) -> None:
        """vrdxCmdSortIndirect: the element count is read from device memory at sort time."""
        k = _as_i32(keys)
        m = k.numel() if max_count is None else int(max_count)
        st = storage if storage is not None else self.storage_for(m, False, stream)
        api.vrdxCmdSortIndirect(self._stream(stream), self.handle, m, count_buffer.data_ptr(), count_offset,
                                k.data_ptr(), 0, st.data_ptr(), 0, query_pool, query)
        self.check()

    def sort_key_value_indirect(self, keys: torch.Tensor, values: torch.Tensor, count_buffer: torch.Tensor,
                                max_count: int | None = None, count_offset: int = 0,
                                storage: torch.Tensor | None = None, stream=None, query_pool=None,
                                query: int = 0) -> None:
        k, v = _as_i32(keys), _as_i32(values)
        m = k.numel() if max_count is None else int(max_count)
        st = storage if storage is not None else self.storage_for(m, True, stream)
        api.vrdxCmdSortKeyValueIndirect(self._stream(stream), self.handle, m, count_buffer.data_ptr(),
                                        count_offset, k.data_ptr(), 0, v.data_ptr(), 0, st.data_ptr(), 0,
                                        query_pool, query)
        self.check()

    def sort_ex(self, keys: torch.Tensor, values: torch.Tensor | None = None, *, key_type: int | None = None,
                descending: bool = False, begin_bit: int = 0, end_bit: int = 32,
                count_buffer: torch.Tensor | None = None, max_count: int | None = None, count_offset: int = 0,
                storage: torch.Tensor | None = None, stream=None, query_pool=None, query: int = 0) -> None:
        """``vrdxCudaCmdSortEx``: typed keys (uint32 / int32 / float32, inferred from the tensor dtype unless
        given), ascending or descending, optional bit sub-range, optional payload, optional device count."""
        if key_type is None:
            key_type = {torch.float32: api.VRDX_CUDA_KEY_TYPE_FLOAT32, torch.int32: api.VRDX_CUDA_KEY_TYPE_INT32,
                        torch.uint32: api.VRDX_CUDA_KEY_TYPE_UINT32}.get(keys.dtype)
            if key_type is None:
                raise TypeError(f"keys must be a 32-bit type, got {keys.dtype}")
        k = keys.view(torch.int32) if keys.dtype != torch.int32 else keys
        v = _as_i32(values) if values is not None else None
        m = k.numel() if max_count is None else int(max_count)
        st = storage if storage is not None else self.storage_for(m, v is not None, stream)
        info = api.make_key_info(key_type, api.VRDX_CUDA_SORT_ORDER_DESCENDING if descending
                                 else api.VRDX_CUDA_SORT_ORDER_ASCENDING, begin_bit, end_bit)
        api.vrdxCudaCmdSortEx(self._stream(stream), self.handle, info, m,
                              count_buffer.data_ptr() if count_buffer is not None else None, count_offset,
                              k.data_ptr(), 0, v.data_ptr() if v is not None else None, 0, st.data_ptr(), 0,
                              query_pool, query)
        self.check()

    def sort_keys64(self, keys: torch.Tensor, *, key_type: int | None = None, descending: bool = False,
                    count_buffer: torch.Tensor | None = None, max_count: int | None = None, count_offset: int = 0,
                    storage: torch.Tensor | None = None, stream=None) -> None:
        """``vrdxCudaCmdSortKeys64``: 64-bit keys (int64 / uint64 / float64 tensor), keys only, in place."""
        if key_type is None:
            key_type = {torch.float64: api.VRDX_CUDA_KEY_TYPE_FLOAT32, torch.int64: api.VRDX_CUDA_KEY_TYPE_INT32,
                        torch.uint64: api.VRDX_CUDA_KEY_TYPE_UINT32}.get(keys.dtype)
        if key_type is None or keys.element_size() != 8:
            raise TypeError(f"keys must be a 64-bit type, got {keys.dtype}")
        m = keys.numel() if max_count is None else int(max_count)
        if storage is None:
            storage = self._scratch(int(api.vrdxCudaGetSorterKeys64StorageRequirements(self.handle, m).size), stream)
        info = api.make_key_info(key_type, api.VRDX_CUDA_SORT_ORDER_DESCENDING if descending
                                 else api.VRDX_CUDA_SORT_ORDER_ASCENDING)
        api.vrdxCudaCmdSortKeys64(self._stream(stream), self.handle, info, m,
                                  count_buffer.data_ptr() if count_buffer is not None else None, count_offset,
                                  keys.data_ptr(), 0, storage.data_ptr(), 0)
        self.check()

    @property
    def properties(self):
        return api.vrdxCudaGetSorterProperties(self.handle)

    @property
    def last_launch_count(self) -> int:
        return api.vrdxCudaGetLastLaunchCount(self.handle)

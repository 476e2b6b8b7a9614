"""ctypes binding of libvrdx_b200.so — the reference's C API, name for name.

Every function here is a 1:1 call into the C-ABI declared in ``include/vk_radix_sort.h`` /
``include/vrdx_cuda.h`` (reference: src/vk_radix_sort.h.in:11-81).  There is NO fallback:
if the CUDA library is missing or fails to load, importing the symbols raises.

Handles are plain integers on this side:
  VkDevice / VkPhysicalDevice -> ``cuda_device(ordinal)``            (ordinal + 1)
  VkCommandBuffer             -> ``torch.cuda.Stream.cuda_stream``    (cudaStream_t)
  VkBuffer                    -> ``tensor.data_ptr()``                (device pointer)
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_uint32, c_uint64, c_void_p, POINTER, Structure, byref

from . import build as _build

VK_SUCCESS = 0
VK_NOT_READY = 1
VK_ERROR_OUT_OF_HOST_MEMORY = -1
VK_ERROR_OUT_OF_DEVICE_MEMORY = -2
VK_ERROR_INITIALIZATION_FAILED = -3
VK_ERROR_FEATURE_NOT_PRESENT = -8
VK_BUFFER_USAGE_TRANSFER_DST_BIT = 0x2
VK_BUFFER_USAGE_STORAGE_BUFFER_BIT = 0x20
VK_NULL_HANDLE = None

VRDX_CUDA_ALGORITHM_AUTO = 0
VRDX_CUDA_ALGORITHM_ONESWEEP = 1
VRDX_CUDA_ALGORITHM_REDUCE_THEN_SCAN = 2
VRDX_CUDA_TILE_LOAD_AUTO = 0
VRDX_CUDA_TILE_LOAD_DIRECT = 1
VRDX_CUDA_TILE_LOAD_TMA = 2

VRDX_CUDA_KEY_TYPE_UINT32 = 0
VRDX_CUDA_KEY_TYPE_INT32 = 1
VRDX_CUDA_KEY_TYPE_FLOAT32 = 2
VRDX_CUDA_SORT_ORDER_ASCENDING = 0
VRDX_CUDA_SORT_ORDER_DESCENDING = 1

QUERY_COUNT = 15  # timestamps written per sort (src/vk_radix_sort.h.in:39-50)


class VrdxSorterCreateInfo(Structure):
    """struct VrdxSorterCreateInfo (src/vk_radix_sort.h.in:18-22)."""
    _fields_ = [("physicalDevice", c_void_p), ("device", c_void_p), ("pipelineCache", c_void_p)]


class VrdxSorterStorageRequirements(Structure):
    """struct VrdxSorterStorageRequirements (src/vk_radix_sort.h.in:28-31)."""
    _fields_ = [("size", c_uint64), ("usage", c_uint32)]


class VrdxCudaSorterOptions(Structure):
    _fields_ = [("structSize", c_uint32), ("algorithm", c_int), ("tileLoad", c_int),
                ("reserved", c_uint32 * 5)]


class VrdxCudaSortKeyInfo(Structure):
    """struct VrdxCudaSortKeyInfo (include/vrdx_cuda.h): key type, order and bit range of vrdxCudaCmdSortEx."""
    _fields_ = [("structSize", c_uint32), ("keyType", c_int), ("order", c_int), ("beginBit", c_uint32),
                ("endBit", c_uint32), ("reserved", c_uint32 * 3)]


class VrdxCudaSorterProperties(Structure):
    _fields_ = [("deviceOrdinal", c_int), ("smCount", c_int), ("ccMajor", c_int), ("ccMinor", c_int),
                ("keysTileSize", c_uint32), ("keyValueTileSize", c_uint32),
                ("offsetAlignment", c_uint32), ("maxOnesweepCount", c_uint32)]


# name -> (restype, argtypes); the eight reference entry points first, then the CUDA extensions.
_SIGNATURES = {
    "vrdxCreateSorter": (c_int, [POINTER(VrdxSorterCreateInfo), POINTER(c_void_p)]),
    "vrdxDestroySorter": (None, [c_void_p]),
    "vrdxGetSorterStorageRequirements": (None, [c_void_p, c_uint32, POINTER(VrdxSorterStorageRequirements)]),
    "vrdxGetSorterKeyValueStorageRequirements": (None, [c_void_p, c_uint32, POINTER(VrdxSorterStorageRequirements)]),
    "vrdxCmdSort": (None, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64, c_void_p, c_uint64, c_void_p, c_uint32]),
    "vrdxCmdSortIndirect": (None, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64, c_void_p, c_uint64,
                                   c_void_p, c_uint64, c_void_p, c_uint32]),
    "vrdxCmdSortKeyValue": (None, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64, c_void_p, c_uint64,
                                   c_void_p, c_uint64, c_void_p, c_uint32]),
    "vrdxCmdSortKeyValueIndirect": (None, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64, c_void_p, c_uint64,
                                           c_void_p, c_uint64, c_void_p, c_uint64, c_void_p, c_uint32]),
    "vrdxCudaCreateSorter": (c_int, [POINTER(VrdxSorterCreateInfo), POINTER(VrdxCudaSorterOptions), POINTER(c_void_p)]),
    "vrdxCudaGetLastError": (c_int, [c_void_p]),
    "vrdxCudaGetErrorString": (ctypes.c_char_p, [c_int]),
    "vrdxCudaGetLastLaunchCount": (c_uint32, [c_void_p]),
    "vrdxCudaCreateQueryPool": (c_int, [c_void_p, c_uint32, POINTER(c_void_p)]),
    "vrdxCudaDestroyQueryPool": (None, [c_void_p]),
    "vrdxCudaGetQueryPoolResults": (c_int, [c_void_p, c_uint32, c_uint32, POINTER(c_uint64)]),
    "vrdxCudaImportMemoryFd": (c_int, [c_void_p, c_int, c_uint64, c_int, POINTER(c_void_p)]),
    "vrdxCudaImportedMemoryBuffer": (c_void_p, [c_void_p, c_uint64]),
    "vrdxCudaReleaseImportedMemory": (None, [c_void_p]),
    "vrdxCudaGetSorterProperties": (None, [c_void_p, POINTER(VrdxCudaSorterProperties)]),
    "vrdxCudaCmdSortEx": (None, [c_void_p, c_void_p, POINTER(VrdxCudaSortKeyInfo), c_uint32, c_void_p, c_uint64,
                                 c_void_p, c_uint64, c_void_p, c_uint64, c_void_p, c_uint64, c_void_p, c_uint32]),
    "vrdxCudaGetSorterKeys64StorageRequirements": (None, [c_void_p, c_uint32, POINTER(VrdxSorterStorageRequirements)]),
    "vrdxCudaCmdSortKeys64": (None, [c_void_p, c_void_p, POINTER(VrdxCudaSortKeyInfo), c_uint32, c_void_p, c_uint64,
                                     c_void_p, c_uint64, c_void_p, c_uint64]),
    "vrdxCudaImportSemaphoreFd": (c_int, [c_void_p, c_int, c_int, POINTER(c_void_p)]),
    "vrdxCudaCmdWaitSemaphore": (c_int, [c_void_p, c_void_p, c_uint64]),
    "vrdxCudaCmdSignalSemaphore": (c_int, [c_void_p, c_void_p, c_uint64]),
    "vrdxCudaReleaseImportedSemaphore": (None, [c_void_p]),
    # include/vrdx_dist.h
    "vrdxDistCmdPrefixHistogram": (None, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64, c_uint32, c_uint32, c_uint32,
                                          c_void_p, c_uint64, c_void_p, c_uint64]),
    "vrdxDistCmdClassCount": (None, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64, c_uint32, c_void_p, c_uint64,
                                     c_void_p, c_uint64]),
    "vrdxDistCmdPartitionScatter": (None, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64, c_uint32, c_void_p, c_uint64,
                                           c_void_p, c_uint64, c_uint32, c_void_p, c_uint64]),
    "vrdxDistAllocShared": (c_int, [c_void_p, c_uint64, POINTER(c_void_p), ctypes.c_char_p]),
    "vrdxDistFreeShared": (None, [c_void_p, c_void_p]),
    "vrdxDistOpenShared": (c_int, [c_void_p, ctypes.c_char_p, POINTER(c_void_p)]),
    "vrdxDistCloseShared": (None, [c_void_p, c_void_p]),
    "vrdxDistCmdPartition": (None, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint64, c_uint32, c_void_p, c_uint64,
                                    c_void_p, c_uint64, c_void_p, c_uint64]),
}

REFERENCE_ENTRY_POINTS = tuple(list(_SIGNATURES)[:8])  # the eight functions of src/vk_radix_sort.h.in:24-81
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def load_library() -> ctypes.CDLL:
    """Load libvrdx_b200.so (building it in-tree if absent). Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    # (re)build if the library is missing or older than csrc/ or include/ (no silent stale library);
    # a VRDX_LIB override is loaded as it is
    path = _build.LIB_PATH if os.environ.get("VRDX_LIB") else _build.build()
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def cuda_device(ordinal: int) -> int:
    """vrdxCudaDevice(ordinal) / vrdxCudaPhysicalDevice(ordinal) (include/vrdx_cuda.h)."""
    return int(ordinal) + 1


# ------------------------------------------------------------------ the reference's functions

def vrdxCreateSorter(create_info: VrdxSorterCreateInfo):
    """-> (VkResult, VrdxSorter). Reference: vrdxCreateSorter, src/vk_radix_sort.h.in:24."""
    sorter = c_void_p()
    res = load_library().vrdxCreateSorter(byref(create_info), byref(sorter))
    return res, (sorter.value if res == VK_SUCCESS else None)


def vrdxCudaCreateSorter(create_info: VrdxSorterCreateInfo, algorithm=VRDX_CUDA_ALGORITHM_AUTO,
                         tile_load=VRDX_CUDA_TILE_LOAD_AUTO, reserved=None):
    """``reserved``: optional tile-shape selectors for A/B runs (index + 1; 0 = default):
    [0] keys-only shape, [1] key-value shape (both algorithms), [2] experiment id (libraries built with
    -DVRDX_EXPERIMENTS only; the product library answers VK_ERROR_FEATURE_NOT_PRESENT)."""
    opts = VrdxCudaSorterOptions()
    opts.structSize = ctypes.sizeof(VrdxCudaSorterOptions)
    opts.algorithm = algorithm
    opts.tileLoad = tile_load
    for i, v in enumerate(reserved or ()):
        opts.reserved[i] = int(v)
    sorter = c_void_p()
    res = load_library().vrdxCudaCreateSorter(byref(create_info), byref(opts), byref(sorter))
    return res, (sorter.value if res == VK_SUCCESS else None)


def vrdxDestroySorter(sorter) -> None:
    load_library().vrdxDestroySorter(sorter)


def vrdxGetSorterStorageRequirements(sorter, max_element_count: int) -> VrdxSorterStorageRequirements:
    req = VrdxSorterStorageRequirements()
    load_library().vrdxGetSorterStorageRequirements(sorter, max_element_count, byref(req))
    return req


def vrdxGetSorterKeyValueStorageRequirements(sorter, max_element_count: int) -> VrdxSorterStorageRequirements:
    req = VrdxSorterStorageRequirements()
    load_library().vrdxGetSorterKeyValueStorageRequirements(sorter, max_element_count, byref(req))
    return req


def vrdxCmdSort(command_buffer, sorter, element_count, keys_buffer, keys_offset, storage_buffer,
                storage_offset, query_pool=None, query=0) -> None:
    load_library().vrdxCmdSort(command_buffer, sorter, element_count, keys_buffer, keys_offset,
                               storage_buffer, storage_offset, query_pool, query)


def vrdxCmdSortIndirect(command_buffer, sorter, max_element_count, indirect_buffer, indirect_offset,
                        keys_buffer, keys_offset, storage_buffer, storage_offset, query_pool=None,
                        query=0) -> None:
    load_library().vrdxCmdSortIndirect(command_buffer, sorter, max_element_count, indirect_buffer,
                                       indirect_offset, keys_buffer, keys_offset, storage_buffer,
                                       storage_offset, query_pool, query)


def vrdxCmdSortKeyValue(command_buffer, sorter, element_count, keys_buffer, keys_offset, values_buffer,
                        values_offset, storage_buffer, storage_offset, query_pool=None, query=0) -> None:
    load_library().vrdxCmdSortKeyValue(command_buffer, sorter, element_count, keys_buffer, keys_offset,
                                       values_buffer, values_offset, storage_buffer, storage_offset,
                                       query_pool, query)


def vrdxCmdSortKeyValueIndirect(command_buffer, sorter, max_element_count, indirect_buffer,
                                indirect_offset, keys_buffer, keys_offset, values_buffer, values_offset,
                                storage_buffer, storage_offset, query_pool=None, query=0) -> None:
    load_library().vrdxCmdSortKeyValueIndirect(command_buffer, sorter, max_element_count, indirect_buffer,
                                               indirect_offset, keys_buffer, keys_offset, values_buffer,
                                               values_offset, storage_buffer, storage_offset, query_pool,
                                               query)


# ------------------------------------------------------------------ CUDA extensions

def vrdxCudaGetLastError(sorter) -> int:
    return load_library().vrdxCudaGetLastError(sorter)


def vrdxCudaGetErrorString(code: int) -> str:
    return load_library().vrdxCudaGetErrorString(code).decode()


def vrdxCudaGetLastLaunchCount(sorter) -> int:
    return load_library().vrdxCudaGetLastLaunchCount(sorter)


def vrdxCudaCreateQueryPool(device, query_count: int = QUERY_COUNT):
    pool = c_void_p()
    res = load_library().vrdxCudaCreateQueryPool(device, query_count, byref(pool))
    return res, (pool.value if res == VK_SUCCESS else None)


def vrdxCudaDestroyQueryPool(pool) -> None:
    load_library().vrdxCudaDestroyQueryPool(pool)


def vrdxCudaGetQueryPoolResults(pool, first_query: int = 0, query_count: int = QUERY_COUNT):
    out = (c_uint64 * query_count)()
    res = load_library().vrdxCudaGetQueryPoolResults(pool, first_query, query_count, out)
    return res, list(out)


def make_key_info(key_type: int = VRDX_CUDA_KEY_TYPE_UINT32, order: int = VRDX_CUDA_SORT_ORDER_ASCENDING,
                  begin_bit: int = 0, end_bit: int = 32) -> VrdxCudaSortKeyInfo:
    info = VrdxCudaSortKeyInfo()
    info.structSize = ctypes.sizeof(VrdxCudaSortKeyInfo)
    info.keyType, info.order, info.beginBit, info.endBit = key_type, order, begin_bit, end_bit
    return info


def vrdxCudaCmdSortEx(command_buffer, sorter, key_info, element_count, indirect_buffer, indirect_offset,
                      keys_buffer, keys_offset, values_buffer, values_offset, storage_buffer, storage_offset,
                      query_pool=None, query=0) -> None:
    load_library().vrdxCudaCmdSortEx(command_buffer, sorter, byref(key_info) if key_info is not None else None,
                                     element_count, indirect_buffer, indirect_offset, keys_buffer, keys_offset,
                                     values_buffer, values_offset, storage_buffer, storage_offset, query_pool,
                                     query)


def vrdxCudaGetSorterKeys64StorageRequirements(sorter, max_element_count: int) -> VrdxSorterStorageRequirements:
    req = VrdxSorterStorageRequirements()
    load_library().vrdxCudaGetSorterKeys64StorageRequirements(sorter, max_element_count, byref(req))
    return req


def vrdxCudaCmdSortKeys64(command_buffer, sorter, key_info, element_count, indirect_buffer, indirect_offset,
                          keys_buffer, keys_offset, storage_buffer, storage_offset) -> None:
    load_library().vrdxCudaCmdSortKeys64(command_buffer, sorter, byref(key_info) if key_info is not None else None,
                                         element_count, indirect_buffer, indirect_offset, keys_buffer, keys_offset,
                                         storage_buffer, storage_offset)


def vrdxCudaGetSorterProperties(sorter) -> VrdxCudaSorterProperties:
    p = VrdxCudaSorterProperties()
    load_library().vrdxCudaGetSorterProperties(sorter, byref(p))
    return p

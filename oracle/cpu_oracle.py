"""oracle/cpu_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end to the two CPU checkers:

* ``liboracle.so``  — our C restatement of the reference's algorithm (oracle/lsd_oracle.c);
* ``_ref/libvrdx_ref.so`` — the reference's OWN ``CpuBenchmark`` / ``DataGenerator``
  (bench/cpu_benchmark.cc:19-53, bench/data_generator.cc:8-27) compiled unmodified.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``vulkan_radix_sort_b200``) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libvrdx_ref.so")

_u32p = ctypes.POINTER(ctypes.c_uint32)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(verbose: bool = False) -> None:
    """Compile the checkers (C restatement always; _ref only where /root/reference exists)."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout, out.stderr)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed")


def _ptr(a: np.ndarray):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u32p)


_oracle = None
_ref = None


def oracle_lib() -> ctypes.CDLL:
    global _oracle
    if _oracle is None:
        if not os.path.exists(_ORACLE_SO):
            build()
        lib = ctypes.CDLL(_ORACLE_SO)
        lib.vrdx_oracle_sort_keys.argtypes = [_u32p, ctypes.c_uint64, _u32p]
        lib.vrdx_oracle_sort_keys.restype = ctypes.c_int
        lib.vrdx_oracle_sort_key_value.argtypes = [_u32p, _u32p, ctypes.c_uint64, _u32p, _u32p]
        lib.vrdx_oracle_sort_key_value.restype = ctypes.c_int
        lib.vrdx_oracle_sort_partitioned.argtypes = [_u32p, _u32p, ctypes.c_uint32, ctypes.c_uint32,
                                                     _u32p, _u32p]
        lib.vrdx_oracle_sort_partitioned.restype = ctypes.c_int
        lib.vrdx_oracle_is_sorted.argtypes = [_u32p, ctypes.c_uint64]
        lib.vrdx_oracle_is_sorted.restype = ctypes.c_int
        lib.vrdx_oracle_multiset_fingerprint.argtypes = [_u32p, _u32p, ctypes.c_uint64, _u64p]
        lib.vrdx_oracle_multiset_fingerprint.restype = None
        lib.vrdx_oracle_check_stable_permutation.argtypes = [_u32p, _u32p, _u32p, ctypes.c_uint64]
        lib.vrdx_oracle_check_stable_permutation.restype = ctypes.c_int
        _oracle = lib
    return _oracle


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def ref_lib() -> ctypes.CDLL:
    global _ref
    if _ref is None:
        lib = ctypes.CDLL(_REF_SO)
        lib.vrdx_ref_generate.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, _u32p, _u32p]
        lib.vrdx_ref_sort_keys.argtypes = [_u32p, ctypes.c_uint32, _u32p, _u64p]
        lib.vrdx_ref_sort_key_value.argtypes = [_u32p, _u32p, ctypes.c_uint32, _u32p, _u32p, _u64p]
        _ref = lib
    return _ref


# ----------------------------------------------------------------- C restatement

def sort_keys(keys: np.ndarray) -> np.ndarray:
    """Net-effect restatement: 4 stable counting passes (lsd_oracle.c: vrdx_oracle_sort_keys)."""
    k = np.ascontiguousarray(keys, dtype=np.uint32).copy()
    scratch = np.empty_like(k)
    rc = oracle_lib().vrdx_oracle_sort_keys(_ptr(k), k.size, _ptr(scratch))
    assert rc == 0
    return k


def sort_key_value(keys: np.ndarray, values: np.ndarray):
    k = np.ascontiguousarray(keys, dtype=np.uint32).copy()
    v = np.ascontiguousarray(values, dtype=np.uint32).copy()
    assert k.size == v.size
    sk, sv = np.empty_like(k), np.empty_like(v)
    rc = oracle_lib().vrdx_oracle_sort_key_value(_ptr(k), _ptr(v), k.size, _ptr(sk), _ptr(sv))
    assert rc == 0
    return k, v


def sort_partitioned(keys: np.ndarray, values, count: int):
    """Structural restatement (upsweep/spine/downsweep, 4096-key partitions) with the
    indirect contract: arrays hold max_count elements, only [0,count) is sorted."""
    k = np.ascontiguousarray(keys, dtype=np.uint32).copy()
    sk = np.empty_like(k)
    if values is not None:
        v = np.ascontiguousarray(values, dtype=np.uint32).copy()
        sv = np.empty_like(v)
        rc = oracle_lib().vrdx_oracle_sort_partitioned(_ptr(k), _ptr(v), count, k.size, _ptr(sk), _ptr(sv))
        assert rc == 0
        return k, v
    rc = oracle_lib().vrdx_oracle_sort_partitioned(_ptr(k), None, count, k.size, _ptr(sk), None)
    assert rc == 0
    return k, None


# ----------------------------------------------------------------- key-type / order / bit-range extension

KEY_UINT32, KEY_INT32, KEY_FLOAT32 = 0, 1, 2


def sortable_image(bits: np.ndarray, key_type: int = KEY_UINT32, descending: bool = False) -> np.ndarray:
    """Order-preserving unsigned image of 32-bit keys given as raw uint32 bit patterns.

    No counterpart in the reference; this is the definition the CUDA path
    (include/vrdx_cuda.h: VrdxCudaSortKeyInfo) is checked against, and tests/test_oracle.py pins it
    to NumPy's own typed sorts: int32 flips the sign bit, float32 flips all bits of negatives and the
    sign bit of the rest (IEEE-754 total order), descending complements the image."""
    b = np.ascontiguousarray(bits).view(np.uint32)
    if key_type == KEY_INT32:
        t = b ^ np.uint32(0x80000000)
    elif key_type == KEY_FLOAT32:
        neg = (b >> np.uint32(31)).astype(bool)
        t = np.where(neg, ~b, b ^ np.uint32(0x80000000)).astype(np.uint32)
    else:
        t = b.copy()
    return (~t).astype(np.uint32) if descending else t


def sort_ex(bits: np.ndarray, values=None, key_type: int = KEY_UINT32, descending: bool = False,
            begin_bit: int = 0, end_bit: int = 32):
    """Stable sort by bits [begin_bit, end_bit) of the sortable image; returns (keys, values) as uint32."""
    b = np.ascontiguousarray(bits).view(np.uint32)
    assert 0 <= begin_bit <= end_bit <= 32
    field = sortable_image(b, key_type, descending).astype(np.uint64) >> np.uint64(begin_bit)
    field &= np.uint64((1 << (end_bit - begin_bit)) - 1)
    perm = np.argsort(field, kind="stable")
    return b[perm], (None if values is None else np.ascontiguousarray(values).view(np.uint32)[perm])


def sort_keys64(bits: np.ndarray, key_type: int = KEY_UINT32, descending: bool = False) -> np.ndarray:
    """64-bit counterpart of sort_ex (keys only): sort by the order-preserving unsigned image of the 64-bit
    pattern (KEY_INT32 / KEY_FLOAT32 stand for int64 / float64).  Returns the raw uint64 patterns in order."""
    b = np.ascontiguousarray(bits).view(np.uint64)
    top = np.uint64(1) << np.uint64(63)
    if key_type == KEY_INT32:
        t = b ^ top
    elif key_type == KEY_FLOAT32:
        t = np.where((b >> np.uint64(63)).astype(bool), ~b, b ^ top).astype(np.uint64)
    else:
        t = b.copy()
    if descending:
        t = ~t
    return b[np.argsort(t, kind="stable")]


def is_sorted(keys: np.ndarray) -> bool:
    k = np.ascontiguousarray(keys, dtype=np.uint32)
    return bool(oracle_lib().vrdx_oracle_is_sorted(_ptr(k), k.size))


def multiset_fingerprint(keys: np.ndarray, values=None):
    k = np.ascontiguousarray(keys, dtype=np.uint32)
    out = (ctypes.c_uint64 * 2)()
    vp = None
    if values is not None:
        v = np.ascontiguousarray(values, dtype=np.uint32)
        vp = _ptr(v)
    oracle_lib().vrdx_oracle_multiset_fingerprint(_ptr(k), vp, k.size, out)
    return int(out[0]), int(out[1])


def check_stable_permutation(keys_in, keys_sorted, values_sorted) -> bool:
    a = np.ascontiguousarray(keys_in, dtype=np.uint32)
    b = np.ascontiguousarray(keys_sorted, dtype=np.uint32)
    c = np.ascontiguousarray(values_sorted, dtype=np.uint32)
    return bool(oracle_lib().vrdx_oracle_check_stable_permutation(_ptr(a), _ptr(b), _ptr(c), a.size))


# ----------------------------------------------------------------- the reference itself

def ref_generate(seed: int, size: int, bits: int = 32):
    k = np.empty(size, dtype=np.uint32)
    v = np.empty(size, dtype=np.uint32)
    ref_lib().vrdx_ref_generate(seed, size, bits, _ptr(k), _ptr(v))
    return k, v


def ref_sort_keys(keys: np.ndarray):
    """Reference CpuBenchmark::Sort. Returns (sorted keys, ns of its timed region)."""
    k = np.ascontiguousarray(keys, dtype=np.uint32)
    out = np.empty_like(k)
    ns = ctypes.c_uint64(0)
    ref_lib().vrdx_ref_sort_keys(_ptr(k), k.size, _ptr(out), ctypes.byref(ns))
    return out, int(ns.value)


def ref_sort_key_value(keys: np.ndarray, values: np.ndarray):
    k = np.ascontiguousarray(keys, dtype=np.uint32)
    v = np.ascontiguousarray(values, dtype=np.uint32)
    ok, ov = np.empty_like(k), np.empty_like(v)
    ns = ctypes.c_uint64(0)
    ref_lib().vrdx_ref_sort_key_value(_ptr(k), _ptr(v), k.size, _ptr(ok), _ptr(ov), ctypes.byref(ns))
    return ok, ov, int(ns.value)

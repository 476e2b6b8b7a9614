"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/*.h declares; struct layouts match the reference's; pure host functions behave.
No compute calls here (no GPU in this container)."""
import ctypes
import os
import re
import subprocess

import pytest

from vulkan_radix_sort_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")


@pytest.fixture(scope="module")
def lib():
    build.build()
    return api.load_library()


def _declared_functions(header):
    text = open(os.path.join(INCLUDE, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(vrdx[A-Za-z0-9_]+)\s*\(", text))
    inline = set(re.findall(r"static\s+inline\s+\w+\s+(vrdx[A-Za-z0-9_]+)\s*\(", text))
    return names - inline


def test_library_exports_every_declared_symbol(lib):
    declared = (_declared_functions("vk_radix_sort.h") | _declared_functions("vrdx_cuda.h")
                | _declared_functions("vrdx_dist.h"))
    assert {"vrdxCreateSorter", "vrdxDestroySorter", "vrdxGetSorterStorageRequirements",
            "vrdxGetSorterKeyValueStorageRequirements", "vrdxCmdSort", "vrdxCmdSortIndirect",
            "vrdxCmdSortKeyValue", "vrdxCmdSortKeyValueIndirect"} <= declared
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert set(api.EXPORTED_SYMBOLS) <= declared


def test_symbols_have_c_linkage():
    out = subprocess.run(["nm", "-D", "--defined-only", build.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for name in api.REFERENCE_ENTRY_POINTS:
        assert name in exported  # unmangled


def test_struct_layouts_match_reference():
    # VrdxSorterStorageRequirements {VkDeviceSize size; VkBufferUsageFlags usage;} -> 16 bytes (h.in:28-31)
    assert ctypes.sizeof(api.VrdxSorterStorageRequirements) == 16
    assert api.VrdxSorterStorageRequirements.usage.offset == 8
    # VrdxSorterCreateInfo: three handles (h.in:18-22)
    assert ctypes.sizeof(api.VrdxSorterCreateInfo) == 24


def test_headers_compile_as_c_and_cxx(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#define VRDX_IMPLEMENTATION\n#include "vrdx_cuda.h"\n'
                   "int main(void){VrdxSorterStorageRequirements r; r.size=0; (void)r;"
                   " return (int)(uintptr_t)vrdxCudaDevice(0) - 1 + (VRDX_VERSION != (4<<12));}\n")
    for cc, std in (("gcc", "-std=c11"), ("g++", "-std=c++17")):
        exe = tmp_path / ("t_" + cc)
        flags = ["-x", "c++"] if cc == "g++" else []
        subprocess.run([cc, std, "-Wall", "-Werror", "-I", INCLUDE, *flags, str(src), "-o", str(exe)], check=True)
        assert subprocess.run([str(exe)]).returncode == 0


def test_storage_requirements_are_pure_host_arithmetic(lib):
    # legal without a GPU and with a NULL sorter: pure function of maxElementCount
    prev_k = prev_kv = 0
    for n in (0, 1, 4096, 4097, 1 << 18, (1 << 25) + 3, 1 << 28, (1 << 30) + 5, (1 << 32) - 1):
        k = api.vrdxGetSorterStorageRequirements(None, n)
        kv = api.vrdxGetSorterKeyValueStorageRequirements(None, n)
        assert k.usage == kv.usage == (api.VK_BUFFER_USAGE_STORAGE_BUFFER_BIT | api.VK_BUFFER_USAGE_TRANSFER_DST_BIT)
        assert k.size % 16 == 0 and kv.size % 16 == 0
        assert k.size >= 4 * n and kv.size >= 8 * n          # alt keys (+ alt values) fit; 64-bit math
        assert kv.size >= k.size                                 # (the tables of a key-value sort are the smaller ones:
                                                                 # its tiles are larger)
        assert k.size >= prev_k and kv.size >= prev_kv         # monotone in N
        prev_k, prev_kv = k.size, kv.size
    # not more than the reference's own scratch (h.in:279-308) at the BASELINE sizes:
    # 2^25: 142,610,464 / 276,828,192 B   2^28: 1,140,854,816 / 2,214,596,640 B   
    assert api.vrdxGetSorterStorageRequirements(None, 1 << 28).size <= 1_140_854_816
    assert api.vrdxGetSorterKeyValueStorageRequirements(None, 1 << 28).size <= 2_214_596_640
    assert api.vrdxGetSorterStorageRequirements(None, 1 << 25).size <= 142_610_464 * 1.05
    # monotone across the AUTO crossovers too (storage sized for max must serve every smaller count)
    prev = 0
    prev_kv = 0
    for n in list(range((1 << 25) - 20000, (1 << 25) + 20000, 1021)) + list(range((1 << 27) - 20000, (1 << 27) + 20000, 1021)):
        size = api.vrdxGetSorterStorageRequirements(None, n).size
        size_kv = api.vrdxGetSorterKeyValueStorageRequirements(None, n).size
        assert size >= prev and size_kv >= prev_kv
        prev, prev_kv = size, size_kv


def test_create_sorter_error_paths(lib):
    import torch
    res, h = api.vrdxCreateSorter(api.VrdxSorterCreateInfo(None, None, None))   # NULL device
    assert res == api.VK_ERROR_INITIALIZATION_FAILED and h is None
    res, h = api.vrdxCreateSorter(api.VrdxSorterCreateInfo(api.cuda_device(4096), api.cuda_device(4096), None))
    assert res == api.VK_ERROR_INITIALIZATION_FAILED and h is None
    assert lib.vrdxCreateSorter(None, None) == api.VK_ERROR_INITIALIZATION_FAILED
    api.vrdxDestroySorter(None)  # NULL-safe like the reference (h.in:268)
    if not torch.cuda.is_available():
        # no device here: creation must FAIL, never fall back to a CPU path
        res, h = api.vrdxCreateSorter(api.VrdxSorterCreateInfo(api.cuda_device(0), api.cuda_device(0), None))
        assert res != api.VK_SUCCESS and h is None


def test_no_cpu_fallback_in_product_package():
    pkg = os.path.join(ROOT, "vulkan_radix_sort_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                text = open(os.path.join(dirpath, f)).read()
                assert "cpu_oracle" not in text and "liboracle" not in text and "lsd_oracle" not in text, f


def _prototypes(header):
    """name -> list of C parameter type strings, parsed from a header of this repo."""
    text = open(os.path.join(INCLUDE, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = {}
    for m in re.finditer(r"\b(vrdx[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        parts = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        protos[name] = parts
    return protos


def _ctype_kind(c_param: str) -> str:
    """Coarse ABI class of a C parameter: 'ptr', 'u64', 'u32' or 'i32'."""
    t = c_param.rsplit(" ", 1)[0] if " " in c_param else c_param
    if "*" in c_param or "[" in c_param:
        return "ptr"
    if t in ("VkBuffer", "VkCommandBuffer", "VkDevice", "VkPhysicalDevice", "VkQueryPool", "VkPipelineCache",
             "VrdxSorter", "VrdxCudaImportedMemory", "VrdxCudaImportedSemaphore"):
        return "ptr"   # dispatchable and non-dispatchable handles are pointers on 64-bit
    if t in ("VkDeviceSize", "uint64_t", "size_t"):
        return "u64"
    if t in ("uint32_t",):
        return "u32"
    if t in ("int",):
        return "i32"
    raise AssertionError(f"unclassified C parameter type: {c_param!r}")


def test_ctypes_signatures_match_the_c_prototypes():
    """Every exported function: same parameter count and ABI class in include/*.h and in api._SIGNATURES."""
    kinds = {ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr", ctypes.c_uint64: "u64", ctypes.c_uint32: "u32",
             ctypes.c_int: "i32"}
    protos = {}
    for h in ("vk_radix_sort.h", "vrdx_cuda.h", "vrdx_dist.h"):
        protos.update(_prototypes(h))
    for name, (restype, argtypes) in api._SIGNATURES.items():
        assert name in protos, name
        c_params = protos[name]
        assert len(c_params) == len(argtypes), (name, c_params, argtypes)
        for c_param, a in zip(c_params, argtypes):
            want = _ctype_kind(c_param)
            got = "ptr" if hasattr(a, "contents") or a in (ctypes.c_void_p, ctypes.c_char_p) else kinds[a]
            assert want == got, (name, c_param, a)

"""Vulkan-interop entry points (contract: reference README.md:121-157): every
argument-validation branch of the seven exported functions runs here, on the GPU box, and a
run-time probe says whether a Vulkan loader / ICD exists to do the real export -> import round
trip (there is none in this image: the probe result is printed and recorded by bench.py)."""
import ctypes
import os

import pytest
import torch

from vulkan_radix_sort_b200 import api

pytestmark = pytest.mark.gpu

OK = api.VK_SUCCESS
INIT_FAILED = api.VK_ERROR_INITIALIZATION_FAILED


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    return api.load_library()


def _closed_fd():
    r, w = os.pipe()
    os.close(r)
    os.close(w)
    return r


def test_import_memory_fd_rejects_bad_arguments(lib):
    dev = api.cuda_device(0)
    out = ctypes.c_void_p()
    assert lib.vrdxCudaImportMemoryFd(dev, -1, 4096, 0, ctypes.byref(out)) == INIT_FAILED      # fd < 0
    assert lib.vrdxCudaImportMemoryFd(dev, 3, 0, 0, ctypes.byref(out)) == INIT_FAILED           # size 0
    assert lib.vrdxCudaImportMemoryFd(dev, 3, 4096, 0, None) == INIT_FAILED                     # NULL out-pointer
    assert lib.vrdxCudaImportMemoryFd(api.cuda_device(99), 3, 4096, 0, ctypes.byref(out)) == INIT_FAILED  # no such device
    assert lib.vrdxCudaImportMemoryFd(0, 3, 4096, 0, ctypes.byref(out)) == INIT_FAILED          # VK_NULL_HANDLE device
    assert out.value is None                                                                     # never written on failure


def test_import_memory_fd_with_a_dead_or_foreign_fd_fails_cleanly_and_leaks_nothing(lib):
    dev = api.cuda_device(0)
    out = ctypes.c_void_p()
    free0, _ = torch.cuda.mem_get_info()
    for dedicated in (0, 1):
        assert lib.vrdxCudaImportMemoryFd(dev, _closed_fd(), 1 << 20, dedicated, ctypes.byref(out)) == INIT_FAILED
        fd = os.open("/dev/null", os.O_RDWR)          # open, but not an exported VkDeviceMemory
        try:
            assert lib.vrdxCudaImportMemoryFd(dev, fd, 1 << 20, dedicated, ctypes.byref(out)) == INIT_FAILED
        finally:
            try:
                os.close(fd)
            except OSError:
                pass
    assert out.value is None
    assert torch.cuda.is_available() and int(torch.zeros(4, device="cuda").sum().item()) == 0   # context still healthy
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < (8 << 20)                  # nothing mapped
    assert cudaGetLastErrorIsClear(lib)


def cudaGetLastErrorIsClear(lib) -> bool:
    """A failed import must not leave a sticky CUDA error behind: a sort right after it works."""
    from vulkan_radix_sort_b200 import Sorter
    s = Sorter(0)
    k = torch.randint(-2**31, 2**31 - 1, (10007,), dtype=torch.int32, device="cuda")
    ref = torch.sort(k.view(torch.uint32).to(torch.int64)).values
    s.sort(k)
    torch.cuda.synchronize()
    ok = bool((k.view(torch.uint32).to(torch.int64) == ref).all())
    s.close()
    return ok


def test_imported_memory_buffer_and_release_are_null_safe(lib):
    assert lib.vrdxCudaImportedMemoryBuffer(None, 0) is None          # VK_NULL_HANDLE memory -> VK_NULL_HANDLE buffer
    assert lib.vrdxCudaImportedMemoryBuffer(None, 1 << 40) is None
    lib.vrdxCudaReleaseImportedMemory(None)                           # no-op, like vkFreeMemory(NULL)


def test_import_semaphore_fd_rejects_bad_arguments(lib):
    dev = api.cuda_device(0)
    out = ctypes.c_void_p()
    for timeline in (0, 1):
        assert lib.vrdxCudaImportSemaphoreFd(dev, -1, timeline, ctypes.byref(out)) == INIT_FAILED
        assert lib.vrdxCudaImportSemaphoreFd(dev, 3, timeline, None) == INIT_FAILED
        assert lib.vrdxCudaImportSemaphoreFd(api.cuda_device(99), 3, timeline, ctypes.byref(out)) == INIT_FAILED
        assert lib.vrdxCudaImportSemaphoreFd(dev, _closed_fd(), timeline, ctypes.byref(out)) == INIT_FAILED
    assert out.value is None


def test_semaphore_commands_reject_a_null_semaphore(lib):
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.vrdxCudaCmdWaitSemaphore(stream, None, 1) == INIT_FAILED
    assert lib.vrdxCudaCmdSignalSemaphore(stream, None, 1) == INIT_FAILED
    lib.vrdxCudaReleaseImportedSemaphore(None)
    torch.cuda.synchronize()


def test_vulkan_probe_reports_what_the_box_has():
    """Not an assertion about the box: records whether the real round trip could run here."""
    import bench
    p = bench.probe_vulkan()
    print("vulkan probe:", p)
    assert set(p) >= {"loader", "icd_files", "reference_vulkan_path"}
    if p["loader_loads"] and p["nvidia_icd"]:
        pytest.xfail("a Vulkan loader and an NVIDIA ICD are present: the export -> import -> sort round trip should be added")

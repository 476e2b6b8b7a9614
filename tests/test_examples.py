"""The C-API example (examples/c_api_example.cc) is the INTEGRATION.md usage as a real program: it must build against
include/*.h + libvrdx_b200.so with a plain host compiler (CPU test) and produce correct results on a B200 (GPU test)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = os.path.join(ROOT, "examples")
EXE = os.path.join(EXAMPLES, "c_api_example")


def _build():
    from vulkan_radix_sort_b200 import build
    build.build()
    out = subprocess.run(["make", "-C", EXAMPLES, "c_api_example"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert os.path.exists(EXE)


def test_c_api_example_builds_with_a_host_compiler():
    _build()
    # it binds the eight reference entry points it uses by name: no unresolved vrdx symbol may remain
    syms = subprocess.run(["nm", "-u", EXE], capture_output=True, text=True).stdout
    for name in ("vrdxCreateSorter", "vrdxGetSorterKeyValueStorageRequirements", "vrdxCmdSortKeyValueIndirect",
                 "vrdxCudaCmdSortEx", "vrdxDestroySorter"):
        assert name in syms, name   # undefined in the executable = resolved from libvrdx_b200.so at load time
    ldd = subprocess.run(["ldd", EXE], capture_output=True, text=True).stdout
    assert "libvrdx_b200.so" in ldd and "not found" not in ldd.split("libvrdx_b200.so")[1].splitlines()[0]


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1000003, 40000001])
def test_c_api_example_runs_correctly(n):
    _build()
    out = subprocess.run([EXE, str(n)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 mismatches" in out.stdout

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU checkers (oracle/): C restatement + the reference's own CpuBenchmark when built."""
    from oracle import cpu_oracle
    cpu_oracle.oracle_lib()
    return cpu_oracle


@pytest.fixture(scope="session")
def sorter():
    """A vrdx sorter on cuda:0 through the C-ABI. No fallback: fails if the library is missing."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vulkan_radix_sort_b200 import Sorter
    s = Sorter(0)
    yield s
    s.close()


FLAVOURS = {
    # name: (algorithm, tile_load, reserved selectors: [0] keys shape + 1, [1] key-value shape + 1, [2] experiment)
    "onesweep": ("ONESWEEP", "DIRECT", None),
    "reduce_then_scan": ("REDUCE_THEN_SCAN", "DIRECT", None),
    "auto": ("AUTO", "AUTO", None),
    # alternate tile shapes compiled into the product library (kKeysShapes / kPairShapes in csrc/vrdx_api.cu)
    "onesweep_256x16": ("ONESWEEP", "DIRECT", (2, 2)),
    "reduce_then_scan_512x16": ("REDUCE_THEN_SCAN", "DIRECT", (4, 4)),
    # losing variants, only in libraries built with VRDX_EXPERIMENTS=1 (skipped on the product library)
    "x_onesweep_tma_persistent": ("ONESWEEP", "TMA", None),
    "x_reduce_then_scan_tma_persistent": ("REDUCE_THEN_SCAN", "TMA", None),
    "x_onesweep_cluster4_lookback": ("ONESWEEP", "DIRECT", (0, 0, 2)),
    "x_round1_tile_kernel": ("AUTO", "DIRECT", (0, 0, 1)),
}


@pytest.fixture(scope="session", params=list(FLAVOURS))
def any_sorter(request):
    """Every algorithm / tile-load flavour the library ships, each through the same C-ABI."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vulkan_radix_sort_b200 import Sorter, api
    algo, load, reserved = FLAVOURS[request.param]
    try:
        s = Sorter(0, algorithm=getattr(api, "VRDX_CUDA_ALGORITHM_" + algo),
                   tile_load=getattr(api, "VRDX_CUDA_TILE_LOAD_" + load), reserved=reserved)
    except RuntimeError as e:
        if request.param.startswith("x_") and str(api.VK_ERROR_FEATURE_NOT_PRESENT) in str(e):
            pytest.skip("experimental variant: not in the product library (build with VRDX_EXPERIMENTS=1)")
        raise
    s.kind = request.param
    yield s
    s.close()

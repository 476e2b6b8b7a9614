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

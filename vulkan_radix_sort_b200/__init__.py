"""vulkan_radix_sort_b200 — a B200-native (sm_100a) 32-bit LSD radix sort behind the
vk_radix_sort C API (jaesung-cs/vulkan_radix_sort v0.4.0).

The product is ``lib/libvrdx_b200.so`` (C-ABI, see ``include/``).  This package only holds
the in-tree build script, the ctypes binding of that ABI and thin host-side helpers.
There is no CPU fallback anywhere in the package.
"""
from . import api, build, datagen  # noqa: F401

__all__ = ["api", "build", "datagen", "Sorter"]
__version__ = "0.1.0"


def __getattr__(name):  # torch is imported only when device memory is actually needed
    if name == "Sorter":
        from .sorter import Sorter
        return Sorter
    raise AttributeError(name)

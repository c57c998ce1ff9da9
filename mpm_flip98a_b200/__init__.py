"""B200-native MLS-MPM substep engine: Python host side over the C-ABI (include/mpm.h).

The product is libmpm.so (hand-written sm_100a CUDA, built from csrc/).  This package only binds it
with ctypes -- the same way ``exec.py`` of the reference would -- and fails
loudly when the library or a CUDA device is missing: there is no CPU fallback.
"""
from .engine import Engine, Group, MpmError, Config, Material, load_library, LIB_PATH  # noqa: F401
from . import scenes  # noqa: F401

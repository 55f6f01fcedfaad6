"""ataxxzero_b200 -- B200-native self-play hot path of AtaxxZero behind a C ABI.

Python here is a thin host layer (ctypes) over libataxxzero.so; every compute call runs
hand-written sm_100a CUDA kernels and fails loudly when the library or the GPU is missing.
"""
from ._native import AzError, Context, Position, lib  # noqa: F401

__version__ = "0.1.0"

"""recom_b200 — B200-native fused embedding-column stage behind RECom's op surface.

The package holds only what the hot path needs: csrc/ (hand-written sm_100a kernels
and the C++ host side behind the C ABI of include/recom_b200.h) and thin Python
plumbing (ctypes binding, column schema, synthetic configs, sharding). Importing it
loads librecom_b200.so and raises if the library has not been built: there is no CPU
or PyTorch fallback.
"""
from . import _lib  # noqa: F401  (loads the shared library; raises if it is missing)
from . import schema, synth  # noqa: F401

__all__ = ["_lib", "schema", "synth"]

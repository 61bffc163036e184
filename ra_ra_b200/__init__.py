"""ra_ra_b200: B200-native execution back end for ra-ra's traversal hot path (`ra::ply`).

  abi   ctypes mirror of include/racu.h (the C ABI)
  ra    host-side mirror of ra-ra's array surface (views, slicing, expression trees -> racu_desc)
  cuda  the executor over the CUDA library ra_ra_b200/libracu.so; importing it fails loudly if the library is missing
"""
from . import abi, ra  # noqa: F401

__all__ = ["abi", "ra"]

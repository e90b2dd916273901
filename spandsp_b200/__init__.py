"""spandsp_b200 - B200-native multi-channel tone-detect / demod engine behind spandsp's C API.

The product is the sm_100a shared library ``libspandsp_b200.so`` (sources in ``csrc/``, C ABI in
``include/``); this package only holds its build script and a ctypes binding used by the tests
and the benchmark.
"""
from . import build as _build  # noqa: F401

__all__ = ["engine", "build"]

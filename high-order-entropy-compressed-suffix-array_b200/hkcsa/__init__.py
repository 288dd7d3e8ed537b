"""hkcsa -- B200-native hot path of the H_k-CSA index (suffix array -> BWT ->
wavelet tree with rank/select directories; batched FM backward search).

Python host layer over libhkcsa.so (hand-written sm_100a CUDA behind a C-ABI,
include/hkcsa.h).  Importing this package never touches the GPU; the first
call into the engine loads the library and raises if it has not been built.
"""
from . import _lib
from ._lib import HkcsaError  # noqa: F401

__all__ = ["_lib", "HkcsaError", "engine"]


def __getattr__(name):
    if name == "engine":
        import importlib
        return importlib.import_module(".engine", __name__)
    raise AttributeError(name)

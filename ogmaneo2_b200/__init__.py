"""ogmaneo2_b200 — B200-native implementation of OgmaNeo2's per-step Sparse Predictive Hierarchy update.

The product is libogn_b200.so (hand-written sm_100a CUDA kernels + C++ host classes that mirror the reference's
ogmaneo:: API, see include/). This package is the Python host side: a ctypes binding of the library's flat C ABI.
There is no CPU or PyTorch fallback: `lib()` raises if the library is not built, and every compute call fails loudly if
no sm_100 device is present.
"""
import ctypes
import os

from .flat_api import FlatHierarchy, FlatLib, InputType, LayerDesc  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libogn_b200.so")
_lib = None


class ExtensionMissing(RuntimeError):
    pass


def cdll() -> ctypes.CDLL:
    """The raw shared library (built in-tree by `python -m ogmaneo2_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ExtensionMissing(
                f"{LIB_PATH} is not built: run `python -m ogmaneo2_b200.build`. The CUDA extension is the only "
                "implementation of the step path; there is no CPU fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def lib() -> FlatLib:
    """The product's export (prefix ogn_) of the flat hierarchy ABI."""
    return FlatLib(cdll(), "ogn_")


from .hierarchy import Hierarchy, CreateOptions  # noqa: E402,F401

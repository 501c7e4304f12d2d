"""ctypes binding of the ImageEncoder C ABI (include/ogn_image_encoder_api.h): ogmaneo::ImageEncoder
(ImageEncoder.h:18-154) behind a (library, prefix) pair like flat_api.FlatLib — prefix ogn_ is the CUDA product, oport_ /
oref_ the CPU checkers under oracle/ (test infrastructure). No computation happens here."""
import ctypes as C
from typing import Sequence

import numpy as np

_VP = C.c_void_p
_IP = C.POINTER(C.c_int)
_FP = C.POINTER(C.c_float)


class _CVisibleDesc(C.Structure):
    _fields_ = [("size_x", C.c_int), ("size_y", C.c_int), ("size_z", C.c_int), ("radius", C.c_int)]


_SIGS = {
    "ienc_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_CVisibleDesc), C.c_uint32, C.POINTER(_VP)]),
    "ienc_destroy": (None, [_VP]),
    "ienc_last_error": (C.c_char_p, []),
    "ienc_step": (C.c_int, [_VP, C.POINTER(_FP), C.c_int]),
    "ienc_reconstruct": (C.c_int, [_VP, _IP]),
    "ienc_hidden_cs": (C.c_int, [_VP, _IP]),
    "ienc_hidden_resources": (C.c_int, [_VP, _FP]),
    "ienc_reconstruction": (C.c_int, [_VP, C.c_int, _FP]),
    "ienc_weights": (C.c_int64, [_VP, C.c_int, _FP, C.c_int64]),
    "ienc_set_params": (C.c_int, [_VP, C.c_float, C.c_float]),
    "ienc_save": (C.c_int64, [_VP, C.c_char_p]),
    "ienc_load": (C.c_int, [C.c_char_p, C.POINTER(_VP)]),
}
IENC_API_NAMES = tuple(_SIGS)


class IEncLib:
    def __init__(self, lib: C.CDLL, prefix: str):
        self.lib, self.prefix = lib, prefix
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, prefix + name)
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    def error(self) -> str:
        m = self.ienc_last_error()
        return m.decode() if m else ""


class ImageEncoder:
    """ogmaneo::ImageEncoder: initRandom (constructor), step, reconstruct, getters, save / load."""

    def __init__(self, lib: IEncLib, hiddenSize: Sequence[int] = None, visibleLayers: Sequence[tuple] = None, seed: int = 1234, _handle=None):
        """visibleLayers: [((size_x, size_y, size_z), radius), ...] — ImageEncoder::VisibleLayerDesc."""
        self.f = lib
        self.h = _VP()
        if _handle is not None:
            self.h = _handle
            return
        self.hiddenSize = tuple(int(v) for v in hiddenSize)
        self.visible = [(tuple(int(v) for v in s), int(r)) for s, r in visibleLayers]
        d = (_CVisibleDesc * len(self.visible))()
        for c, (s, r) in zip(d, self.visible):
            c.size_x, c.size_y, c.size_z, c.radius = s[0], s[1], s[2], r
        rc = lib.ienc_create(*self.hiddenSize, len(self.visible), d, seed & 0xFFFFFFFF, C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"{lib.prefix}ienc_create failed ({rc}): {lib.error()}")

    @classmethod
    def createOn(cls, lib: IEncLib, hierarchy, hiddenSize, visibleLayers, seed: int = 1234):
        """CUDA product only: an encoder on the device and stream of `hierarchy` (a `ogmaneo2_b200.Hierarchy`), so that
        stepDevice() / hiddenCsDevice() feed hierarchy.stepDevice() without a host round trip."""
        fn = getattr(lib.lib, lib.prefix + "ienc_create_on")
        fn.restype = C.c_int
        fn.argtypes = [_VP, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_CVisibleDesc), C.c_uint32, C.POINTER(_VP)]
        vis = [(tuple(int(v) for v in s), int(r)) for s, r in visibleLayers]
        d = (_CVisibleDesc * len(vis))()
        for c, (s, r) in zip(d, vis):
            c.size_x, c.size_y, c.size_z, c.radius = s[0], s[1], s[2], r
        h = _VP()
        rc = fn(hierarchy.h, *[int(v) for v in hiddenSize], len(vis), d, seed & 0xFFFFFFFF, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"{lib.prefix}ienc_create_on failed ({rc}): {lib.error()}")
        obj = cls(lib, _handle=h)
        obj.hiddenSize, obj.visible = tuple(int(v) for v in hiddenSize), vis
        obj._keep = hierarchy  # the hierarchy owns the stream: keep it alive
        return obj

    def stepDevice(self, inputPtrs: Sequence[int], learnEnabled: bool = True):
        """Dense inputs already in device memory (one pointer per visible layer); asynchronous."""
        fn = getattr(self.f.lib, self.f.prefix + "ienc_step_device")
        fn.restype, fn.argtypes = C.c_int, [_VP, C.POINTER(C.c_void_p), C.c_int]
        ptrs = (C.c_void_p * len(inputPtrs))(*[int(p) for p in inputPtrs])
        self._chk(fn(self.h, ptrs, int(learnEnabled)), "ienc_step_device")

    def hiddenCsDevice(self) -> int:
        fn = getattr(self.f.lib, self.f.prefix + "ienc_hidden_cs_device")
        fn.restype, fn.argtypes = C.c_void_p, [_VP]
        return int(fn(self.h))

    def synchronize(self):
        fn = getattr(self.f.lib, self.f.prefix + "ienc_synchronize")
        fn.restype, fn.argtypes = C.c_int, [_VP]
        self._chk(fn(self.h), "ienc_synchronize")

    @classmethod
    def load(cls, lib: IEncLib, path: str, hiddenSize, visibleLayers):
        h = _VP()
        rc = lib.ienc_load(path.encode(), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"{lib.prefix}ienc_load failed ({rc}): {lib.error()}")
        obj = cls(lib, _handle=h)
        obj.hiddenSize = tuple(hiddenSize)
        obj.visible = [(tuple(s), int(r)) for s, r in visibleLayers]
        return obj

    def close(self):
        if self.h:
            self.f.ienc_destroy(self.h)
            self.h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{self.f.prefix}{what} failed ({rc}): {self.f.error()}")

    def step(self, inputActs: Sequence[np.ndarray], learnEnabled: bool = True):
        arrs = [np.ascontiguousarray(a, dtype=np.float32).reshape(-1) for a in inputActs]
        for a, (s, _) in zip(arrs, self.visible):
            if a.size != s[0] * s[1] * s[2]:
                raise ValueError("dense input size does not match the visible layer")
        ptrs = (_FP * len(arrs))(*[a.ctypes.data_as(_FP) for a in arrs])
        self._chk(self.f.ienc_step(self.h, ptrs, int(learnEnabled)), "ienc_step")

    def reconstruct(self, hiddenCs: np.ndarray):
        a = np.ascontiguousarray(hiddenCs, dtype=np.int32).reshape(-1)
        if a.size != self.hiddenSize[0] * self.hiddenSize[1]:
            raise ValueError("hidden CSDR size does not match the hidden grid")
        self._chk(self.f.ienc_reconstruct(self.h, a.ctypes.data_as(_IP)), "ienc_reconstruct")

    def getHiddenCs(self) -> np.ndarray:
        out = np.empty(self.f.ienc_hidden_cs(self.h, None), dtype=np.int32)
        self.f.ienc_hidden_cs(self.h, out.ctypes.data_as(_IP))
        return out

    def getHiddenResources(self) -> np.ndarray:
        out = np.empty(self.f.ienc_hidden_resources(self.h, None), dtype=np.float32)
        self.f.ienc_hidden_resources(self.h, out.ctypes.data_as(_FP))
        return out

    def getReconstruction(self, vli: int) -> np.ndarray:
        out = np.empty(self.f.ienc_reconstruction(self.h, vli, None), dtype=np.float32)
        self.f.ienc_reconstruction(self.h, vli, out.ctypes.data_as(_FP))
        return out

    def getWeights(self, vli: int) -> np.ndarray:
        n = self.f.ienc_weights(self.h, vli, None, 0)
        out = np.empty(n, dtype=np.float32)
        self.f.ienc_weights(self.h, vli, out.ctypes.data_as(_FP), n)
        return out

    def setParams(self, alpha: float = 0.01, gamma: float = 0.01):
        self._chk(self.f.ienc_set_params(self.h, alpha, gamma), "ienc_set_params")

    def save(self, path: str) -> int:
        n = self.f.ienc_save(self.h, path.encode())
        if n < 0:
            raise RuntimeError(f"ienc_save failed: {self.f.error()}")
        return n

    def snapshot(self) -> dict:
        s = {"hiddenCs": self.getHiddenCs(), "resources": self.getHiddenResources()}
        for v in range(len(self.visible)):
            s[f"w{v}"] = self.getWeights(v)
            s[f"recon{v}"] = self.getReconstruction(v)
        return s

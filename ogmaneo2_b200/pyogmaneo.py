"""pyogmaneo-style names over the B200 path.

The reference repository ships no Python binding; user scripts written for the OgmaNeo2 SWIG binding use a small
vocabulary — `ComputeSystem`, `Int3`, `LayerDesc`, `Hierarchy(cs, inputSizes, inputTypes, layerDescs)`,
`h.step(cs, inputCs, learnEnabled, reward, mimic)`, `h.getPredictionCs(i)`, the `inputType*` constants — that maps one to
one onto the C++ classes this package re-implements (include/ogmaneo/Hierarchy.h, ComputeSystem.h, Helpers.h). This
module offers those names on top of `ogmaneo2_b200.Hierarchy`, so that such a script runs on the CUDA path by changing
its import:

    from ogmaneo2_b200 import pyogmaneo
    cs = pyogmaneo.ComputeSystem()
    lds = [pyogmaneo.LayerDesc() for _ in range(4)]
    for ld in lds: ld.hiddenSize = pyogmaneo.Int3(4, 4, 16)
    h = pyogmaneo.Hierarchy(cs, [pyogmaneo.Int3(1, 1, 32)], [pyogmaneo.inputTypePrediction], lds)
    h.step(cs, [[7]], True)
    print(h.getPredictionCs(0))

It is a naming layer only: every call goes through the flat C ABI of libogn_b200.so (no CPU fallback).
"""
import dataclasses
from typing import List, Sequence

import numpy as np

from . import hierarchy as _hier
from .flat_api import InputType
from .flat_api import LayerDesc as _FlatLayerDesc

inputTypeNone = int(InputType.none)
inputTypePrediction = int(InputType.prediction)
inputTypeAction = int(InputType.action)


@dataclasses.dataclass
class Int3:
    """ogmaneo::Int3 (Helpers.h:38-53)."""
    x: int = 0
    y: int = 0
    z: int = 0

    def __iter__(self):
        return iter((self.x, self.y, self.z))


class ComputeSystem:
    """ogmaneo::ComputeSystem (ComputeSystem.h:17-41): the seed of its rng is what the hierarchy's initial weights depend
    on; thread counts are accepted and ignored (the work runs on the GPU)."""

    def __init__(self, seed: int = 1234, device: int = -1):
        self.seed = int(seed)
        self.device = int(device)

    @staticmethod
    def setNumThreads(n: int):
        from . import lib
        lib().set_num_threads(int(n))

    @staticmethod
    def getNumThreads() -> int:
        return 1


class LayerDesc:
    """ogmaneo::Hierarchy::LayerDesc (Hierarchy.h:42-66), same member names and defaults."""

    def __init__(self, hiddenSize: Int3 = None, ffRadius: int = 2, pRadius: int = 2, ticksPerUpdate: int = 2, temporalHorizon: int = 4,
                 aRadius: int = 2, historyCapacity: int = 32):
        self.hiddenSize = hiddenSize if hiddenSize is not None else Int3(4, 4, 16)
        self.ffRadius, self.pRadius, self.ticksPerUpdate = ffRadius, pRadius, ticksPerUpdate
        self.temporalHorizon, self.aRadius, self.historyCapacity = temporalHorizon, aRadius, historyCapacity

    def _flat(self) -> _FlatLayerDesc:
        return _FlatLayerDesc(hiddenSize=tuple(self.hiddenSize), ffRadius=self.ffRadius, pRadius=self.pRadius,
                              ticksPerUpdate=self.ticksPerUpdate, temporalHorizon=self.temporalHorizon, aRadius=self.aRadius,
                              historyCapacity=self.historyCapacity)


class _SCLayer:
    """What `h.getSCLayer(l)` hands out: ogmaneo::SparseCoder's public members and getters (SparseCoder.h:83-147) as live
    properties of the device-resident layer (reads and writes go through the flat ABI)."""

    def __init__(self, h, l):
        self._h, self._l = h, l

    @property
    def alpha(self) -> float:
        return self._h.getSCAlpha(self._l)

    @alpha.setter
    def alpha(self, v: float):
        self._h.setSCAlpha(self._l, float(v))

    def getHiddenCs(self) -> List[int]: return self._h.scHiddenCs(self._l).tolist()
    def getHiddenCsPrev(self) -> List[int]: return self._h.scHiddenCsPrev(self._l).tolist()
    def getHiddenSize(self) -> Int3: return Int3(*self._h.layerHiddenSize(0, self._l))
    def getNumVisibleLayers(self) -> int: return self._h.scNumVisible(self._l)

    def getVisibleLayerDesc(self, v: int) -> "VisibleLayerDesc":
        d = self._h.layerVisibleDesc(0, self._l, v)
        return VisibleLayerDesc(Int3(*d[:3]), d[3])

    def getWeights(self, v: int) -> np.ndarray:
        """getVisibleLayer(v).weights.nonZeroValues (reference CSR order)."""
        return self._h.scWeights(self._l, v)


class _PLayer:
    """An element of `h.getPLayers(l)`: ogmaneo::Predictor (Predictor.h:82-146)."""

    def __init__(self, h, l, p):
        self._h, self._l, self._p = h, l, p

    @property
    def alpha(self) -> float:
        return self._h.getPredAlpha(self._l, self._p)

    @alpha.setter
    def alpha(self, v: float):
        self._h.setPredAlpha(self._l, self._p, float(v))

    def getHiddenCs(self) -> List[int]: return self._h.predHiddenCs(self._l, self._p).tolist()
    def getHiddenSize(self) -> Int3: return Int3(*self._h.layerHiddenSize(1 + self._p, self._l))
    def getNumVisibleLayers(self) -> int: return self._h.predNumVisible(self._l, self._p)

    def getVisibleLayerDesc(self, v: int) -> "VisibleLayerDesc":
        d = self._h.layerVisibleDesc(1 + self._p, self._l, v)
        return VisibleLayerDesc(Int3(*d[:3]), d[3])

    def getWeights(self, v: int) -> np.ndarray: return self._h.predWeights(self._l, self._p, v)


class _ALayer:
    """An element of `h.getALayers()`: ogmaneo::Actor (Actor.h:108-179); alpha, beta, gamma, minSteps and historyIters
    are live properties."""

    def __init__(self, h, p):
        object.__setattr__(self, "_h", h)
        object.__setattr__(self, "_p", p)

    _PARAMS = ("alpha", "beta", "gamma", "minSteps", "historyIters")

    def __getattr__(self, name):
        if name in _ALayer._PARAMS:
            return self._h.getActorParams(self._p)[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name not in _ALayer._PARAMS:
            raise AttributeError(name)
        cur = self._h.getActorParams(self._p)
        cur[name] = value
        self._h.setActorParams(self._p, float(cur["alpha"]), float(cur["beta"]), float(cur["gamma"]), int(cur["minSteps"]),
                               int(cur["historyIters"]))

    def getHiddenCs(self) -> List[int]: return self._h.actorHiddenCs(self._p).tolist()
    def getHiddenSize(self) -> Int3: return Int3(*self._h.layerHiddenSize(-1 - self._p))
    def getNumVisibleLayers(self) -> int: return self._h.actorNumVisible(self._p)

    def getVisibleLayerDesc(self, v: int) -> "VisibleLayerDesc":
        d = self._h.layerVisibleDesc(-1 - self._p, 0, v)
        return VisibleLayerDesc(Int3(*d[:3]), d[3])


@dataclasses.dataclass
class VisibleLayerDesc:
    """SparseCoder / Predictor / Actor::VisibleLayerDesc (SparseCoder.h:18-29)."""
    size: Int3
    radius: int


class State:
    """ogmaneo::State (Hierarchy.h:26-36): what getState() returns and setState() takes."""

    def __init__(self, opaque):
        self._s = opaque


class Hierarchy:
    """ogmaneo::Hierarchy (Hierarchy.h:90-215). Construct with (cs, inputSizes, inputTypes, layerDescs) — initRandom — or
    with (cs, fileName, inputSizes=...) — readFromStream."""

    def __init__(self, cs: ComputeSystem, inputSizes_or_file, inputTypes: Sequence[int] = None, layerDescs: Sequence[LayerDesc] = None,
                 inputSizes: Sequence[Int3] = None):
        self._cs = cs
        if isinstance(inputSizes_or_file, str):
            from . import lib
            sizes = [tuple(s) for s in (inputSizes or [])]
            flat = _hier._bind_product(lib())
            self._h = _hier.FlatHierarchy.load(flat, inputSizes_or_file, sizes, cs.seed)
            self._sizes = sizes
        else:
            self._sizes = [tuple(s) for s in inputSizes_or_file]
            self._h = _hier.Hierarchy(self._sizes, [int(t) for t in inputTypes], [ld._flat() for ld in layerDescs], seed=cs.seed,
                                      options=_hier.CreateOptions(device=cs.device))

    def step(self, cs: ComputeSystem, inputCs: Sequence[Sequence[int]], learnEnabled: bool = True, reward: float = 0.0, mimic: bool = False):
        self._h.step([np.asarray(c, dtype=np.int32) for c in inputCs], learnEnabled, reward, mimic)

    def getPredictionCs(self, i: int) -> List[int]:
        return self._h.getPredictionCs(i).tolist()

    def getNumLayers(self) -> int:
        return self._h.getNumLayers()

    def getHiddenCs(self, l: int) -> List[int]:
        return self._h.scHiddenCs(l).tolist()

    def getUpdate(self, l: int) -> bool:
        return self._h.getUpdate(l)

    def getTicks(self, l: int) -> int:
        return self._h.getTicks(l)

    def getTicksPerUpdate(self, l: int) -> int:
        return self._h.getTicksPerUpdate(l)

    def getInputSizes(self) -> List[Int3]:
        return [Int3(*s) for s in self._sizes]

    def save(self, fileName: str) -> int:
        """Hierarchy::writeToStream into a file (same bytes as the reference writes)."""
        return self._h.save(fileName)

    # --- layers (Hierarchy.h:180-215): proxies whose attributes read and write the device-resident layer -----------------
    def getSCLayer(self, l: int) -> _SCLayer:
        return _SCLayer(self._h, l)

    def getPLayers(self, l: int) -> list:
        """One entry per predictor slot of layer l, None where the reference holds a null pointer."""
        return [_PLayer(self._h, l, p) if self._h.predExists(l, p) else None for p in range(self._h.predCount(l))]

    def getALayers(self) -> list:
        return [_ALayer(self._h, p) if self._h.actorExists(p) else None for p in range(len(self._sizes))]

    def getState(self) -> State:
        return State(self._h.getState())

    def setState(self, state: State):
        self._h.setState(state._s)

    # --- the flat accessors of the SWIG binding, for scripts that use those ----------------------------------------------
    def getHiddenSize(self, l: int) -> Int3: return self.getSCLayer(l).getHiddenSize()
    def getNumSCVisibleLayers(self, l: int) -> int: return self._h.scNumVisible(l)
    def pLayerExists(self, l: int, v: int) -> bool: return self._h.predExists(l, v)
    def aLayerExists(self, v: int) -> bool: return self._h.actorExists(v)
    def getSCAlpha(self, l: int) -> float: return self._h.getSCAlpha(l)
    def setSCAlpha(self, l: int, alpha: float): self._h.setSCAlpha(l, float(alpha))
    def getPAlpha(self, l: int, v: int) -> float: return self._h.getPredAlpha(l, v)
    def setPAlpha(self, l: int, v: int, alpha: float): self._h.setPredAlpha(l, v, float(alpha))
    def getAAlpha(self, v: int) -> float: return self._h.getActorParams(v)["alpha"]
    def setAAlpha(self, v: int, x: float): self.getALayers()[v].alpha = x
    def getABeta(self, v: int) -> float: return self._h.getActorParams(v)["beta"]
    def setABeta(self, v: int, x: float): self.getALayers()[v].beta = x
    def getAGamma(self, v: int) -> float: return self._h.getActorParams(v)["gamma"]
    def setAGamma(self, v: int, x: float): self.getALayers()[v].gamma = x
    def getAMinSteps(self, v: int) -> int: return self._h.getActorParams(v)["minSteps"]
    def setAMinSteps(self, v: int, x: int): self.getALayers()[v].minSteps = x
    def getAHistoryIters(self, v: int) -> int: return self._h.getActorParams(v)["historyIters"]
    def setAHistoryIters(self, v: int, x: int): self.getALayers()[v].historyIters = x


class ImageEncoderVisibleLayerDesc:
    """ogmaneo::ImageEncoder::VisibleLayerDesc (ImageEncoder.h:21-33), same member names and defaults."""

    def __init__(self, size: Int3 = None, radius: int = 2):
        self.size = size if size is not None else Int3(4, 4, 16)
        self.radius = radius


class ImageEncoder:
    """ogmaneo::ImageEncoder (ImageEncoder.h:18-154) on the CUDA path: construct with (cs, hiddenSize, visibleLayerDescs) —
    initRandom — or with (cs, fileName, hiddenSize=..., visibleLayerDescs=...) — readFromStream. `alpha` and `gamma` are the
    public members of the reference (defaults 0.01). With `hierarchy=` the encoder runs on that hierarchy's stream, so its
    hidden CSDR can feed `hierarchy.step` without leaving the device (stepDevice / hiddenCsDevice of the wrapped object)."""

    def __init__(self, cs: ComputeSystem, hiddenSize_or_file, visibleLayerDescs: Sequence[ImageEncoderVisibleLayerDesc] = None,
                 hiddenSize: Int3 = None, hierarchy: "Hierarchy" = None):
        from . import cdll
        from . import image_encoder as _ienc
        lib = _ienc.IEncLib(cdll(), "ogn_")
        vis = [(tuple(d.size), int(d.radius)) for d in (visibleLayerDescs or [])]
        self._alpha, self._gamma = 0.01, 0.01
        if isinstance(hiddenSize_or_file, str):
            self._e = _ienc.ImageEncoder.load(lib, hiddenSize_or_file, tuple(hiddenSize), vis)
        elif hierarchy is not None:
            self._e = _ienc.ImageEncoder.createOn(lib, hierarchy._h, tuple(hiddenSize_or_file), vis, seed=cs.seed)
        else:
            self._e = _ienc.ImageEncoder(lib, tuple(hiddenSize_or_file), vis, seed=cs.seed)

    alpha = property(lambda self: self._alpha, lambda self, v: self._set(alpha=v))
    gamma = property(lambda self: self._gamma, lambda self, v: self._set(gamma=v))

    def _set(self, alpha=None, gamma=None):
        self._alpha = self._alpha if alpha is None else float(alpha)
        self._gamma = self._gamma if gamma is None else float(gamma)
        self._e.setParams(self._alpha, self._gamma)

    def step(self, cs: ComputeSystem, inputActs: Sequence[Sequence[float]], learnEnabled: bool = True):
        self._e.step([np.asarray(a, dtype=np.float32) for a in inputActs], learnEnabled)

    def reconstruct(self, cs: ComputeSystem, hiddenCs: Sequence[int]):
        self._e.reconstruct(np.asarray(hiddenCs, dtype=np.int32))

    def getHiddenCs(self) -> List[int]:
        return self._e.getHiddenCs().tolist()

    def getReconstruction(self, i: int) -> List[float]:
        return self._e.getReconstruction(i).tolist()

    def getHiddenSize(self) -> Int3:
        return Int3(*self._e.hiddenSize)

    def getNumVisibleLayers(self) -> int:
        return len(self._e.visible)

    def getVisibleLayerDesc(self, i: int) -> ImageEncoderVisibleLayerDesc:
        s, r = self._e.visible[i]
        return ImageEncoderVisibleLayerDesc(Int3(*s), r)

    def save(self, fileName: str) -> int:
        """ImageEncoder::writeToStream into a file (same bytes as the reference writes)."""
        return self._e.save(fileName)

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

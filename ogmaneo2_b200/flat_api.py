"""ctypes binding of the flat hierarchy C ABI declared in include/ogn_hierarchy_api.h.

The ABI is exported with three prefixes (ogn_ = product, oport_/oref_ = CPU checkers under oracle/); this module
binds whichever (library, prefix) pair it is given and contains no computation of its own. The Python-facing
classes mirror the reference's C++ interface for the step path (names, argument meaning, defaults):

    LayerDesc            <-> ogmaneo::Hierarchy::LayerDesc          (Hierarchy.h:42-66)
    InputType            <-> ogmaneo::InputType                     (Hierarchy.h:19-23)
    FlatHierarchy.step   <-> Hierarchy::step                        (Hierarchy.cpp:192-285)
    .getPredictionCs     <-> Hierarchy::getPredictionCs             (Hierarchy.h:144-151)
"""
import ctypes as C
import dataclasses
import enum
from typing import List, Sequence

import numpy as np


class InputType(enum.IntEnum):
    none = 0
    prediction = 1
    action = 2


@dataclasses.dataclass
class LayerDesc:
    hiddenSize: tuple = (4, 4, 16)
    ffRadius: int = 2
    pRadius: int = 2
    ticksPerUpdate: int = 2
    temporalHorizon: int = 4
    aRadius: int = 2
    historyCapacity: int = 32


class _CLayerDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "hidden_x", "hidden_y", "hidden_z", "ff_radius", "p_radius", "ticks_per_update", "temporal_horizon",
        "a_radius", "history_capacity")]


class _CHierarchyDesc(C.Structure):
    _fields_ = [("num_inputs", C.c_int), ("input_sizes", C.POINTER(C.c_int)), ("input_types", C.POINTER(C.c_int)),
                ("num_layers", C.c_int), ("layers", C.POINTER(_CLayerDesc))]


_IP = C.POINTER(C.c_int)
_FP = C.POINTER(C.c_float)
_VP = C.c_void_p

# name -> (restype, argtypes)
_SIGS = {
    "create": (C.c_int, [C.POINTER(_CHierarchyDesc), C.c_uint32, C.POINTER(_VP)]),
    "destroy": (None, [_VP]),
    "last_error": (C.c_char_p, []),
    "set_num_threads": (C.c_int, [C.c_int]),
    "step": (C.c_int, [_VP, C.POINTER(_IP), C.c_int, C.c_float, C.c_int]),
    "num_layers": (C.c_int, [_VP]),
    "get_update": (C.c_int, [_VP, C.c_int]),
    "get_ticks": (C.c_int, [_VP, C.c_int]),
    "get_ticks_per_update": (C.c_int, [_VP, C.c_int]),
    "get_prediction_cs": (C.c_int, [_VP, C.c_int, _IP]),
    "sc_hidden_cs": (C.c_int, [_VP, C.c_int, _IP]),
    "sc_hidden_cs_prev": (C.c_int, [_VP, C.c_int, _IP]),
    "sc_num_visible": (C.c_int, [_VP, C.c_int]),
    "sc_weights": (C.c_int64, [_VP, C.c_int, C.c_int, _FP, C.c_int64]),
    "sc_reconstructions": (C.c_int, [_VP, C.c_int, C.c_int, _FP]),
    "pred_count": (C.c_int, [_VP, C.c_int]),
    "pred_exists": (C.c_int, [_VP, C.c_int, C.c_int]),
    "pred_hidden_cs": (C.c_int, [_VP, C.c_int, C.c_int, _IP]),
    "pred_activations": (C.c_int, [_VP, C.c_int, C.c_int, _FP]),
    "pred_num_visible": (C.c_int, [_VP, C.c_int, C.c_int]),
    "pred_weights": (C.c_int64, [_VP, C.c_int, C.c_int, C.c_int, _FP, C.c_int64]),
    "pred_input_cs_prev": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, _IP]),
    "actor_exists": (C.c_int, [_VP, C.c_int]),
    "actor_hidden_cs": (C.c_int, [_VP, C.c_int, _IP]),
    "actor_values": (C.c_int, [_VP, C.c_int, _FP]),
    "actor_num_visible": (C.c_int, [_VP, C.c_int]),
    "actor_weights": (C.c_int64, [_VP, C.c_int, C.c_int, C.c_int, _FP, C.c_int64]),
    "set_sc_alpha": (C.c_int, [_VP, C.c_int, C.c_float]),
    "set_pred_alpha": (C.c_int, [_VP, C.c_int, C.c_int, C.c_float]),
    "set_actor_params": (C.c_int, [_VP, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]),
    "get_sc_alpha": (C.c_float, [_VP, C.c_int]),
    "get_pred_alpha": (C.c_float, [_VP, C.c_int, C.c_int]),
    "get_actor_params": (C.c_int, [_VP, C.c_int, _FP, _FP, _FP, _IP, _IP]),
    "layer_hidden_size": (C.c_int, [_VP, C.c_int, C.c_int, _IP]),
    "layer_visible_desc": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, _IP]),
    "save": (C.c_int64, [_VP, C.c_char_p]),
    "load": (C.c_int, [C.c_char_p, C.c_uint32, C.POINTER(_VP)]),
    "get_state": (C.c_int, [_VP, C.POINTER(_VP)]),
    "set_state": (C.c_int, [_VP, _VP]),
    "free_state": (None, [_VP]),
    "rng_peek": (C.c_uint32, [_VP]),
}

FLAT_API_NAMES = tuple(_SIGS)


class FlatLib:
    """One (shared library, prefix) export of the flat hierarchy ABI."""

    def __init__(self, lib: C.CDLL, prefix: str):
        self.lib, self.prefix = lib, prefix
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, prefix + name)  # AttributeError if a declared symbol is missing
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    def error(self) -> str:
        m = self.last_error()
        return m.decode() if m else ""


def _ip(a: np.ndarray):
    return a.ctypes.data_as(_IP)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_FP)


class HierarchyState:
    """Opaque ogmaneo::State (Hierarchy.h:26-36) owned by Python; freed with the object."""

    def __init__(self, flat: FlatLib, s):
        self.f, self.s = flat, s

    def __del__(self):
        try:
            if self.s:
                self.f.free_state(self.s)
                self.s = None
        except Exception:
            pass


class FlatHierarchy:
    """A hierarchy handle behind the flat ABI. Mirrors ogmaneo::Hierarchy for the step path."""

    def __init__(self, flat: FlatLib, inputSizes: Sequence[tuple] = None, inputTypes: Sequence[int] = None,
                 layerDescs: Sequence[LayerDesc] = None, seed: int = 1234, _handle=None):
        self.f = flat
        self.h = _VP()
        if _handle is not None:
            self.h = _handle
            return
        self.inputSizes = [tuple(int(v) for v in s) for s in inputSizes]
        sizes = (C.c_int * (3 * len(inputSizes)))(*[int(v) for s in inputSizes for v in s])
        types = (C.c_int * len(inputTypes))(*[int(t) for t in inputTypes])
        lds = (_CLayerDesc * len(layerDescs))()
        for c, d in zip(lds, layerDescs):
            c.hidden_x, c.hidden_y, c.hidden_z = (int(v) for v in d.hiddenSize)
            c.ff_radius, c.p_radius, c.ticks_per_update = d.ffRadius, d.pRadius, d.ticksPerUpdate
            c.temporal_horizon, c.a_radius, c.history_capacity = d.temporalHorizon, d.aRadius, d.historyCapacity
        desc = _CHierarchyDesc(len(inputSizes), sizes, types, len(layerDescs), lds)
        rc = flat.create(C.byref(desc), seed & 0xFFFFFFFF, C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"{flat.prefix}create failed ({rc}): {flat.error()}")
        self.layerDescs = list(layerDescs)

    @classmethod
    def load(cls, flat: FlatLib, path: str, inputSizes: Sequence[tuple], seed: int = 1234):
        h = _VP()
        rc = flat.load(path.encode(), seed & 0xFFFFFFFF, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"{flat.prefix}load failed ({rc}): {flat.error()}")
        obj = cls(flat, _handle=h)
        obj.inputSizes = [tuple(s) for s in inputSizes]
        return obj

    def close(self):
        if self.h:
            self.f.destroy(self.h)
            self.h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{self.f.prefix}{what} failed ({rc}): {self.f.error()}")

    # --- Hierarchy::step -------------------------------------------------------------------------------------
    def step(self, inputCs: Sequence[np.ndarray], learnEnabled: bool = True, reward: float = 0.0, mimic: bool = False):
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in inputCs]
        for a, s in zip(arrs, self.inputSizes):
            if a.size != s[0] * s[1]:
                raise ValueError("input CSDR size does not match the input's column count")
        ptrs = (_IP * len(arrs))(*[_ip(a) for a in arrs])
        self._chk(self.f.step(self.h, ptrs, int(learnEnabled), float(reward), int(mimic)), "step")

    # --- getters ---------------------------------------------------------------------------------------------
    def _ints(self, fn, *idx):
        n = fn(self.h, *idx, None)
        if n < 0:
            raise RuntimeError(f"getter failed: {self.f.error()}")
        out = np.empty(n, dtype=np.int32)
        fn(self.h, *idx, _ip(out))
        return out

    def _floats(self, fn, *idx):
        n = fn(self.h, *idx, None)
        if n < 0:
            raise RuntimeError(f"getter failed: {self.f.error()}")
        out = np.empty(n, dtype=np.float32)
        fn(self.h, *idx, _fp(out))
        return out

    def _weights(self, fn, *idx):
        n = fn(self.h, *idx, None, 0)
        if n < 0:
            raise RuntimeError(f"getter failed: {self.f.error()}")
        out = np.empty(n, dtype=np.float32)
        fn(self.h, *idx, _fp(out), n)
        return out

    def getNumLayers(self): return self.f.num_layers(self.h)
    def getUpdate(self, l): return bool(self.f.get_update(self.h, l))
    def getTicks(self, l): return self.f.get_ticks(self.h, l)
    def getTicksPerUpdate(self, l): return self.f.get_ticks_per_update(self.h, l)
    def getPredictionCs(self, i): return self._ints(self.f.get_prediction_cs, i)
    def scHiddenCs(self, l): return self._ints(self.f.sc_hidden_cs, l)
    def scHiddenCsPrev(self, l): return self._ints(self.f.sc_hidden_cs_prev, l)
    def scNumVisible(self, l): return self.f.sc_num_visible(self.h, l)
    def scWeights(self, l, vli): return self._weights(self.f.sc_weights, l, vli)
    def scReconstructions(self, l, vli): return self._floats(self.f.sc_reconstructions, l, vli)
    def predCount(self, l): return self.f.pred_count(self.h, l)
    def predExists(self, l, p): return bool(self.f.pred_exists(self.h, l, p))
    def predHiddenCs(self, l, p): return self._ints(self.f.pred_hidden_cs, l, p)
    def predActivations(self, l, p): return self._floats(self.f.pred_activations, l, p)
    def predNumVisible(self, l, p): return self.f.pred_num_visible(self.h, l, p)
    def predWeights(self, l, p, vli): return self._weights(self.f.pred_weights, l, p, vli)
    def predInputCsPrev(self, l, p, vli): return self._ints(self.f.pred_input_cs_prev, l, p, vli)
    def actorExists(self, p): return bool(self.f.actor_exists(self.h, p))
    def actorHiddenCs(self, p): return self._ints(self.f.actor_hidden_cs, p)
    def actorValues(self, p): return self._floats(self.f.actor_values, p)
    def actorNumVisible(self, p): return self.f.actor_num_visible(self.h, p)
    def actorWeights(self, p, vli, which): return self._weights(self.f.actor_weights, p, vli, which)
    def setSCAlpha(self, l, a): self._chk(self.f.set_sc_alpha(self.h, l, a), "set_sc_alpha")
    def setPredAlpha(self, l, p, a): self._chk(self.f.set_pred_alpha(self.h, l, p, a), "set_pred_alpha")

    def setActorParams(self, p, alpha=0.02, beta=0.02, gamma=0.99, minSteps=8, historyIters=8):
        self._chk(self.f.set_actor_params(self.h, p, alpha, beta, gamma, minSteps, historyIters), "set_actor_params")

    def getSCAlpha(self, l): return float(self.f.get_sc_alpha(self.h, l))
    def getPredAlpha(self, l, p): return float(self.f.get_pred_alpha(self.h, l, p))

    def getActorParams(self, p) -> dict:
        a, b, g = C.c_float(), C.c_float(), C.c_float()
        ms, hi = C.c_int(), C.c_int()
        self._chk(self.f.get_actor_params(self.h, p, C.byref(a), C.byref(b), C.byref(g), C.byref(ms), C.byref(hi)), "get_actor_params")
        return dict(alpha=a.value, beta=b.value, gamma=g.value, minSteps=ms.value, historyIters=hi.value)

    def layerHiddenSize(self, kind: int, l: int = 0) -> tuple:
        """kind: 0 SparseCoder of layer l, 1 + p Predictor p of layer l, -1 - p Actor of input p."""
        out = (C.c_int * 3)()
        self._chk(self.f.layer_hidden_size(self.h, kind, l, out), "layer_hidden_size")
        return tuple(out)

    def layerVisibleDesc(self, kind: int, l: int, vli: int) -> tuple:
        """(size.x, size.y, size.z, radius) of visible layer vli; kind as in layerHiddenSize."""
        out = (C.c_int * 4)()
        self._chk(self.f.layer_visible_desc(self.h, kind, l, vli, out), "layer_visible_desc")
        return tuple(out)

    def save(self, path: str) -> int:
        n = self.f.save(self.h, path.encode())
        if n < 0:
            raise RuntimeError(f"save failed: {self.f.error()}")
        return n

    def rngPeek(self) -> int: return int(self.f.rng_peek(self.h))

    def getState(self) -> "HierarchyState":
        """Hierarchy::getState (Hierarchy.cpp:434-466): a snapshot of every CSDR, history ring, tick and update flag."""
        s = _VP()
        self._chk(self.f.get_state(self.h, C.byref(s)), "get_state")
        return HierarchyState(self.f, s)

    def setState(self, state: "HierarchyState"):
        """Hierarchy::setState (Hierarchy.cpp:468-493): rewind to a snapshot taken from this (or an identical) hierarchy."""
        self._chk(self.f.set_state(self.h, state.s), "set_state")

    # --- everything the parity tests compare, as one dict ------------------------------------------------------
    def snapshot(self, weights: bool = False) -> dict:
        s = {}
        L = self.getNumLayers()
        for l in range(L):
            s[f"update{l}"] = np.array([self.getUpdate(l)], dtype=np.int32)
            s[f"ticks{l}"] = np.array([self.getTicks(l)], dtype=np.int32)
            s[f"sc{l}.hiddenCs"] = self.scHiddenCs(l)
            s[f"sc{l}.hiddenCsPrev"] = self.scHiddenCsPrev(l)
            if weights:
                for v in range(self.scNumVisible(l)):
                    s[f"sc{l}.w{v}"] = self.scWeights(l, v)
            for p in range(self.predCount(l)):
                if not self.predExists(l, p):
                    continue
                s[f"p{l}.{p}.hiddenCs"] = self.predHiddenCs(l, p)
                s[f"p{l}.{p}.act"] = self.predActivations(l, p)
                for v in range(self.predNumVisible(l, p)):
                    s[f"p{l}.{p}.inputCsPrev{v}"] = self.predInputCsPrev(l, p, v)
                    if weights:
                        s[f"p{l}.{p}.w{v}"] = self.predWeights(l, p, v)
        for i in range(len(self.inputSizes)):
            if self.actorExists(i):
                s[f"a{i}.hiddenCs"] = self.actorHiddenCs(i)
                s[f"a{i}.values"] = self.actorValues(i)
                if weights:
                    for v in range(self.actorNumVisible(i)):
                        s[f"a{i}.wv{v}"] = self.actorWeights(i, v, 0)
                        s[f"a{i}.wa{v}"] = self.actorWeights(i, v, 1)
        return s

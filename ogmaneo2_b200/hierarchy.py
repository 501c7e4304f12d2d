"""Python host mirror of ogmaneo::Hierarchy for the step path, bound to libogn_b200.so (include/ogn_b200.h).

    h = Hierarchy(inputSizes, inputTypes, layerDescs, seed=1234)        # ComputeSystem.rng.seed + initRandom
    h.step([csdr0, ...], learnEnabled=True, reward=0.0, mimic=False)    # Hierarchy::step (Hierarchy.cpp:192-285)
    pred = h.getPredictionCs(0)                                         # Hierarchy::getPredictionCs

plus the device-resident extensions (stepDevice, ensembles, sharding) that have no counterpart in the reference.
"""
import ctypes as C
import dataclasses
from typing import Callable, Optional, Sequence

import numpy as np

from . import flat_api
from .flat_api import FlatHierarchy, FlatLib, _CHierarchyDesc, _CLayerDesc, _IP, _VP

ALL_GATHER_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)


class _COptions(C.Structure):
    _fields_ = [("device", C.c_int), ("device_init", C.c_int), ("replay_rng", C.c_int), ("shard_rank", C.c_int),
                ("shard_world", C.c_int), ("all_gather", ALL_GATHER_FN), ("all_gather_user", C.c_void_p),
                ("ensemble", C.c_int), ("ensemble_seeds", C.POINTER(C.c_uint32)), ("peer_exchange", C.c_int)]


@dataclasses.dataclass
class CreateOptions:
    device: int = -1
    deviceInit: bool = False      # counter-hash init on the device instead of the reference's mt19937 order
    replayRng: bool = True        # keep ComputeSystem::rng bit-identical to the reference (cheap for small layers)
    shardRank: int = 0
    shardWorld: int = 1
    allGather: Optional[Callable] = None  # f(buffer_ptr: int, count_per_rank: int, stream_ptr: int)
    ensemble: int = 1
    ensembleSeeds: Optional[Sequence[int]] = None
    peerExchange: bool = False    # shardWorld > 1, one process per GPU: the library's own peer-memory all-gather kernel


_PRODUCT_SIGS = {
    "create_ex": (C.c_int, [C.POINTER(_CHierarchyDesc), C.c_uint32, C.POINTER(_COptions), C.POINTER(_VP)]),
    "step_device": (C.c_int, [_VP, C.POINTER(C.c_void_p), C.c_int, C.c_float, C.c_int]),
    "prediction_cs_device": (C.c_void_p, [_VP, C.c_int]),
    "stream": (C.c_void_p, [_VP]),
    "prediction_readback_queue": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int]),
    "prediction_readback_finish": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_int)]),
    "synchronize": (C.c_int, [_VP]),
    "ensemble_size": (C.c_int, [_VP]),
    "take_sc_update_count": (C.c_uint64, [_VP, C.c_int]),
    "device_count": (C.c_int, []),
    "checkpoint_save": (C.c_int64, [_VP, C.c_char_p]),
    "checkpoint_load": (C.c_int, [C.c_char_p, C.POINTER(_COptions), C.POINTER(_VP)]),
    "shard_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "peer_export": (C.c_int, [_VP, C.c_void_p]),
    "peer_attach": (C.c_int, [_VP, C.c_void_p, C.c_int]),
    "peer_take_wait_ns": (C.c_uint64, [_VP]),
    "profile_enable": (C.c_int, [_VP, C.c_int]),
    "profile_read": (C.c_int, [_VP, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
}
PRODUCT_API_NAMES = tuple(_PRODUCT_SIGS)


def _bind_product(flat: FlatLib):
    for name, (res, args) in _PRODUCT_SIGS.items():
        fn = getattr(flat.lib, "ogn_" + name)
        fn.restype, fn.argtypes = res, args
        setattr(flat, name, fn)
    return flat


class Hierarchy(FlatHierarchy):
    """A device-resident hierarchy. Inherits step()/getPredictionCs()/getters from the flat binding."""

    def __init__(self, inputSizes, inputTypes, layerDescs, seed: int = 1234, options: CreateOptions = None):
        from . import lib
        flat = _bind_product(lib())
        opt = options or CreateOptions()
        self.options = opt
        self.inputSizes = [tuple(int(v) for v in s) for s in inputSizes]
        sizes = (C.c_int * (3 * len(inputSizes)))(*[int(v) for s in inputSizes for v in s])
        types = (C.c_int * len(inputTypes))(*[int(t) for t in inputTypes])
        lds = (_CLayerDesc * len(layerDescs))()
        for c, d in zip(lds, layerDescs):
            c.hidden_x, c.hidden_y, c.hidden_z = (int(v) for v in d.hiddenSize)
            c.ff_radius, c.p_radius, c.ticks_per_update = d.ffRadius, d.pRadius, d.ticksPerUpdate
            c.temporal_horizon, c.a_radius, c.history_capacity = d.temporalHorizon, d.aRadius, d.historyCapacity
        desc = _CHierarchyDesc(len(inputSizes), sizes, types, len(layerDescs), lds)
        co = _COptions()
        co.device, co.device_init, co.replay_rng = opt.device, int(opt.deviceInit), int(opt.replayRng)
        co.shard_rank, co.shard_world, co.ensemble = opt.shardRank, opt.shardWorld, opt.ensemble
        co.peer_exchange = int(opt.peerExchange)
        self._cb = None
        if opt.allGather is not None:
            user_fn = opt.allGather

            def _tramp(_user, buf, count, stream):
                user_fn(buf, count, stream)
            self._cb = ALL_GATHER_FN(_tramp)  # keep alive for the life of the handle
            co.all_gather = self._cb
        self._seeds = None
        if opt.ensembleSeeds is not None:
            self._seeds = (C.c_uint32 * len(opt.ensembleSeeds))(*[int(s) & 0xFFFFFFFF for s in opt.ensembleSeeds])
            co.ensemble_seeds = self._seeds
            co.ensemble = len(opt.ensembleSeeds)
        h = _VP()
        rc = flat.create_ex(C.byref(desc), seed & 0xFFFFFFFF, C.byref(co), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"ogn_create_ex failed ({rc}): {flat.error()}")
        FlatHierarchy.__init__(self, flat, _handle=h)
        self.inputSizes = [tuple(int(v) for v in s) for s in inputSizes]
        self.layerDescs = list(layerDescs)
        self.ensemble = flat.ensemble_size(h)

    def step(self, inputCs, learnEnabled=True, reward=0.0, mimic=False):
        arrs = [np.ascontiguousarray(a, dtype=np.int32).reshape(-1) for a in inputCs]
        for a, s in zip(arrs, self.inputSizes):
            if a.size != s[0] * s[1] * self.ensemble:
                raise ValueError("input CSDR size does not match columns x ensemble size")
        ptrs = (_IP * len(arrs))(*[a.ctypes.data_as(_IP) for a in arrs])
        self._chk(self.f.step(self.h, ptrs, int(learnEnabled), float(reward), int(mimic)), "step")

    def stepDevice(self, inputPtrs: Sequence[int], learnEnabled=True, reward=0.0, mimic=False):
        """inputPtrs: device addresses (ints) of int32 [member][columns] buffers. Asynchronous."""
        ptrs = (C.c_void_p * len(inputPtrs))(*[C.c_void_p(int(p)) for p in inputPtrs])
        self._chk(self.f.step_device(self.h, ptrs, int(learnEnabled), float(reward), int(mimic)), "step_device")

    def checkpointSave(self, path: str) -> int:
        """Hierarchy::writeCheckpoint: everything, weights in their device layouts, 64-bit counts. Returns bytes written."""
        n = int(self.f.checkpoint_save(self.h, path.encode()))
        if n < 0:
            raise RuntimeError(f"ogn_checkpoint_save failed: {self.f.error()}")
        return n

    @classmethod
    def checkpointLoad(cls, path: str, inputSizes, options: "CreateOptions" = None):
        from . import lib
        flat = _bind_product(lib())
        opt = options or CreateOptions()
        co = _COptions()
        co.device, co.device_init, co.replay_rng = opt.device, 1, int(opt.replayRng)
        co.shard_rank, co.shard_world, co.ensemble = opt.shardRank, opt.shardWorld, 1
        co.peer_exchange = int(opt.peerExchange)
        h = _VP()
        rc = flat.checkpoint_load(path.encode(), C.byref(co), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"ogn_checkpoint_load failed ({rc}): {flat.error()}")
        obj = cls.__new__(cls)
        FlatHierarchy.__init__(obj, flat, _handle=h)
        obj.options, obj._cb, obj._seeds = opt, None, None
        obj.inputSizes = [tuple(int(v) for v in s) for s in inputSizes]
        obj.ensemble = flat.ensemble_size(h)
        return obj

    def peerExport(self) -> bytes:
        """The 64-byte CUDA IPC handle of this rank's exchange arena (CreateOptions.peerExchange)."""
        buf = C.create_string_buffer(64)
        self._chk(self.f.peer_export(self.h, buf), "peer_export")
        return buf.raw

    def peerAttach(self, handles: Sequence[bytes]):
        """handles: every rank's peerExport() in rank order (gathered by the caller, e.g. all_gather_object)."""
        blob = b"".join(bytes(h) for h in handles)
        if len(blob) != 64 * len(handles):
            raise ValueError("peer handles must be 64 bytes each")
        self._chk(self.f.peer_attach(self.h, C.c_char_p(blob), len(handles)), "peer_attach")

    def peerTakeWaitNs(self) -> int:
        return int(self.f.peer_take_wait_ns(self.h))

    def predictionCsDevice(self, i: int) -> int:
        return int(self.f.prediction_cs_device(self.h, i) or 0)

    def predictionReadbackQueue(self, i: int, x0: int = 0, x1: int = 0) -> int:
        """Queue an asynchronous copy of x rows [x0, x1) (default: all) of the current prediction of input i; returns a ticket."""
        t = int(self.f.prediction_readback_queue(self.h, i, x0, x1))
        if t < 0:
            raise RuntimeError(f"ogn_prediction_readback_queue failed: {self.f.error()}")
        return t

    def predictionReadbackFinish(self, ticket: int, out: np.ndarray) -> int:
        n = int(self.f.prediction_readback_finish(self.h, ticket, out.ctypes.data_as(C.POINTER(C.c_int))))
        if n < 0:
            raise RuntimeError(f"ogn_prediction_readback_finish failed: {self.f.error()}")
        return n

    def stream(self) -> int:
        return int(self.f.stream(self.h) or 0)

    def synchronize(self):
        self._chk(self.f.synchronize(self.h), "synchronize")

    KERNEL_KINDS = ("sc_forward", "sc_learn", "pred_step", "actor_forward", "actor_learn", "sc_update", "hier_small", "exchange")

    def profileEnable(self, on=True, period: int = 1):
        """Count launches; time (CUDA events) the launches of every period-th step."""
        self._chk(self.f.profile_enable(self.h, (max(int(period), 1) if on else 0)), "profile_enable")

    def profileRead(self, timed: bool = False) -> dict:
        """{kind: (device ms over the timed launches, launches[, timed launches])} accumulated since profileEnable()."""
        ms = (C.c_double * len(self.KERNEL_KINDS))()
        n = (C.c_int64 * len(self.KERNEL_KINDS))()
        nt = (C.c_int64 * len(self.KERNEL_KINDS))()
        self._chk(self.f.profile_read(self.h, ms, n, nt), "profile_read")
        if timed:
            return {k: (ms[i], n[i], nt[i]) for i, k in enumerate(self.KERNEL_KINDS)}
        return {k: (ms[i] * (n[i] / nt[i]) if nt[i] else 0.0, n[i]) for i, k in enumerate(self.KERNEL_KINDS)}

    def takeSCUpdateCount(self, l: int) -> int:
        return int(self.f.take_sc_update_count(self.h, l))

"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, compiled from /root/reference by
oracle/Makefile). Run in the build container:  python tests/golden/make_golden.py

Each fixture holds, for one small configuration, the inputs' generator parameters, the per-step prediction / hidden
CSDRs and the final state (every CSDR, activation and a checksum + sample of every weight matrix) of the reference run
with ONE OpenMP thread (the reference's RNG use races with more). The reference ships no golden vectors of
its own; these pin the oracle port and the CUDA path where /root/reference is absent.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import make_inputs, reward_at  # noqa: E402
from ogmaneo2_b200 import configs  # noqa: E402
from ogmaneo2_b200.flat_api import FlatHierarchy, InputType, LayerDesc  # noqa: E402
from oracle import loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_configs():
    odd = dict(name="odd_5x7x6_to_3x4x5", inputSizes=[(5, 7, 6)], inputTypes=[InputType.prediction],
               layerDescs=[LayerDesc(hiddenSize=(3, 4, 5), ffRadius=1, pRadius=1, temporalHorizon=2),
                           LayerDesc(hiddenSize=(3, 4, 5), ffRadius=1, pRadius=1, temporalHorizon=2)],
               streams=["iid"], seed=99)
    return {
        "c1_sine": (configs.c1_sine(), 300),
        "c2_video_8": (configs.c2_video(8, 3), 40),
        "c3_actor_8": (configs.c3_actor(8), 120),
        "c4_large_16": (configs.c4_large(16), 24),
        "c5_unit_8": (configs.c5_unit(8), 40),
        "odd_geometry": (odd, 30),
    }


def summarize(snapshot):
    """Small fixture form of a full snapshot: CSDRs / activations verbatim, weights as (crc32, first 64 values)."""
    out = {}
    for k, v in snapshot.items():
        name = k.split(".")[-1]
        if name.startswith("w") and v.dtype == np.float32 and v.size > 256:
            out[k + ".crc"] = np.array([zlib.crc32(v.tobytes())], dtype=np.uint32)
            out[k + ".head"] = v[:64].copy()
        else:
            out[k] = v
    return out


def run(flat, cfg, steps):
    """The trajectory a fixture stores; also used by the tests to replay a fixture on another implementation."""
    h = FlatHierarchy(flat, cfg["inputSizes"], cfg["inputTypes"], cfg["layerDescs"], cfg["seed"])
    preds, hidden = [], []
    action = None
    pidx = [i for i, t in enumerate(cfg["inputTypes"]) if t != InputType.none]
    for t in range(steps):
        h.step(make_inputs(cfg, t, action), True, reward_at(t), False)
        if "action" in cfg["streams"]:
            action = h.getPredictionCs(cfg["streams"].index("action"))
        preds.append(np.concatenate([h.getPredictionCs(i) for i in pidx]))
        hidden.append(np.concatenate([h.scHiddenCs(l) for l in range(h.getNumLayers())]))
    out = {"trajectory.pred": np.stack(preds), "trajectory.hidden": np.stack(hidden), "rng_peek": np.array([h.rngPeek()], dtype=np.uint32)}
    out.update(summarize(h.snapshot(True)))
    h.close()
    return out


if __name__ == "__main__":
    ref = loader.ref()
    if ref is None:
        raise SystemExit("oracle/_ref is not built: the reference sources are needed to (re)generate the fixtures")
    ref.set_num_threads(1)
    for name, (cfg, steps) in golden_configs().items():
        data = run(ref, cfg, steps)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, steps=np.array([steps]), **data)
        print(f"{name}: {steps} steps, {len(data)} arrays, {os.path.getsize(path)} bytes")

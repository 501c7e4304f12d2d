#!/usr/bin/env python
"""Stage timeline of the tile-staged ImageEncoder kernel (ogn_debug_ienc_trace): mean cycles between the stamps of a block,
and how the blocks of one launch spread over time. Debug aid for profiles/, not a benchmark.

    python tools/trace_image_encoder.py [--size 128]
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ogmaneo2_b200.image_encoder import IEncLib, ImageEncoder  # noqa: E402

STAGES = ["copies queued", "patch staged", "block landed", "distances", "ranked", "updated", "stored"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    args = ap.parse_args()
    import ogmaneo2_b200 as og
    lib = og.cdll()
    n = args.size
    hidden, visible = (n // 2, n // 2, 32), [((n, n, 3), 4)]
    rng = np.random.RandomState(5)
    enc = ImageEncoder(IEncLib(lib, "ogn_"), hidden, visible, seed=77)
    for t in range(4):
        enc.step([rng.rand(n * n * 3).astype(np.float32)], True)
    blocks = hidden[0] * hidden[1]
    lib.ogn_debug_ienc_trace.argtypes = [ctypes.c_int, ctypes.c_void_p]
    assert lib.ogn_debug_ienc_trace(blocks, None) == 0
    enc.step([rng.rand(n * n * 3).astype(np.float32)], True)
    out = np.zeros((blocks, 12), np.int64)
    assert lib.ogn_debug_ienc_trace(blocks, out.ctypes.data) == 0
    d = np.diff(out[:, :8], axis=1)
    res = {"blocks": blocks, "mean_cycles_per_stage": {s: float(d[:, i].mean()) for i, s in enumerate(STAGES)},
           "p90_cycles_per_stage": {s: float(np.percentile(d[:, i], 90)) for i, s in enumerate(STAGES)},
           "mean_block_cycles": float((out[:, 7] - out[:, 0]).mean()),
           "launch_span_us": float((out[:, 11].max() - out[:, 11].min()) / 1e3),
           "sms_used": int(len(set(out[:, 10].tolist()))),
           "block_start_us_percentiles": [float(x) for x in np.percentile((out[:, 11] - out[:, 11].min()) / 1e3, [10, 25, 50, 75, 90, 100])]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()

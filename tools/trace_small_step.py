#!/usr/bin/env python
"""Phase timeline of the single-launch step (ogn_debug_small_trace) on BASELINE config 2: per phase of one launch, when it
started, how long its work and its barrier took (microseconds, block 0 and the last block). Debug aid.

    OGN_SMALL_KERNEL=1 python tools/trace_small_step.py [--workload c2]
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    args = ap.parse_args()
    os.environ.setdefault("OGN_SMALL_KERNEL", "1")
    import bench
    import ogmaneo2_b200 as og
    from bench import streams
    cfg = bench.workload(args.workload, 1)
    h = og.Hierarchy(cfg["inputSizes"], cfg["inputTypes"], cfg["layerDescs"], seed=cfg["seed"])
    ins = lambda t: [streams.make(k, t, s) for s, k in zip(cfg["inputSizes"], cfg["streams"])]
    lib = og.cdll()
    lib.ogn_debug_small_trace.argtypes = [ctypes.c_int, ctypes.c_void_p]
    kinds = {0: "fwd", 1: "recon", 2: "update", 3: "pred"}
    for t in range(64):
        h.step(ins(t), True)
    out = []
    for t in range(64, 72):
        assert lib.ogn_debug_small_trace(1, None) == 0
        h.step(ins(t), True)
        h.synchronize()
        buf = np.zeros(1536, np.uint64)
        assert lib.ogn_debug_small_trace(0, buf.ctypes.data) == 0
        tr = buf.reshape(2, 256, 3).astype(np.int64)
        n = int((tr[0, :, 0] > 0).sum())
        t0 = tr[0, 0, 0]
        rows = [{"phase": i, "start_us": round((tr[0, i, 0] - t0) / 1e3, 2), "work_us": round((tr[0, i, 1] - tr[0, i, 0]) / 1e3, 2),
                 "barrier_us": round((tr[0, i, 2] - tr[0, i, 1]) / 1e3, 2),
                 "last_block_work_us": round((tr[1, i, 1] - tr[1, i, 0]) / 1e3, 2) if tr[1, i, 0] > 0 else None} for i in range(n)]
        out.append({"step": t, "phases": n, "total_us": round((tr[0, n - 1, 2] - t0) / 1e3, 2), "rows": rows})
    print(json.dumps(out))


if __name__ == "__main__":
    main()

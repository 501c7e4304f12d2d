#!/usr/bin/env python
"""Measurement of the ImageEncoder pre-encoder next to the reference's CPU path.

    python tools/bench_image_encoder.py [--steps K] [--warmup W] [--size N]

One JSON line: encoder steps/s with learning on through the host API (`ogn_ienc_step`: dense fp32 image in host memory,
H2D inside the timed region, synchronous like the reference's call), the algorithmic bytes of a step (every weight is read
once and written once: the Hebbian pass updates all of them), the achieved fraction of the measured HBM peak, the compiled
reference (or the oracle port) timed on the host cores on the same inputs, and a bit-equality check of the hidden CSDRs
and resources of both after the run.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ogmaneo2_b200.image_encoder import IEncLib, ImageEncoder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--size", type=int, default=128, help="input image edge (RGB); hidden grid is size/2 x size/2 x 32, radius 4")
    args = ap.parse_args()
    import ogmaneo2_b200 as og
    from oracle import loader
    n = args.size
    hidden, visible = (n // 2, n // 2, 32), [((n, n, 3), 4)]
    total = args.warmup + args.steps
    rng = np.random.RandomState(5)
    base = rng.rand(n * n * 3).astype(np.float32)
    frames = [np.clip(0.7 * np.roll(base, 3 * 7 * t) + 0.3 * rng.rand(n * n * 3).astype(np.float32), 0, 1).astype(np.float32) for t in range(total)]

    dev = ImageEncoder(IEncLib(og.cdll(), "ogn_"), hidden, visible, seed=77)
    for t in range(args.warmup):
        dev.step([frames[t]], True)
    t0 = time.perf_counter()
    for t in range(args.warmup, total):
        dev.step([frames[t]], True)
    gpu_s = time.perf_counter() - t0

    kind = "reference" if loader.ref() is not None else "port"
    loader.port()
    lib = IEncLib(ctypes.CDLL(loader.REF_SO), "oref_") if kind == "reference" else IEncLib(ctypes.CDLL(loader.PORT_SO), "oport_")
    cpu = ImageEncoder(lib, hidden, visible, seed=77)
    cpu_steps = min(total, 40)
    t0 = time.perf_counter()
    for t in range(cpu_steps):
        cpu.step([frames[t]], True)
    cpu_s = time.perf_counter() - t0
    # parity on the steps both have run: replay the CPU's steps on a fresh device encoder
    chk = ImageEncoder(IEncLib(og.cdll(), "ogn_"), hidden, visible, seed=77)
    for t in range(cpu_steps):
        chk.step([frames[t]], True)
    same = bool(np.array_equal(chk.getHiddenCs(), cpu.getHiddenCs()) and
                np.array_equal(chk.getHiddenResources().view(np.int32), cpu.getHiddenResources().view(np.int32)))

    weights = dev.getWeights(0).size
    alg = 8 * weights  # every weight read once and written once per learning step (the tile-staged kernel's DRAM traffic)
    peak = 6549.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    value = args.steps / gpu_s
    print(json.dumps({"metric": "image_encoder_steps_per_s_learn_on", "value": value, "unit": "steps/s", "ms_per_step": 1e3 * gpu_s / args.steps,
                      "config": {"workload": f"ImageEncoder: {n}x{n}x3 dense input -> {hidden[0]}x{hidden[1]}x{hidden[2]} hidden columns, radius 4",
                                 "weights": int(weights), "api": "ogn_ienc_step(host image), synchronous; H2D inside the timed region"},
                      "roofline": {"bound": "hbm", "alg_bytes_per_step": alg, "achieved": alg * value / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": alg * value / 1e9 / peak},
                      "cpu_baseline": {"kind": kind, "value": cpu_steps / cpu_s, "unit": "steps/s", "cores": os.cpu_count(),
                                       "sample": f"{cpu_steps} steps in {cpu_s:.2f}s"},
                      "parity": {"steps": cpu_steps, "hidden_cs_and_resources_bit_identical": same}}))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""image -> ImageEncoder -> Hierarchy::step, device-resident against the host-buffer path (the caller in
front of the hierarchy). One JSON line.

    python tools/bench_image_pipeline.py [--size 256] [--steps 200] [--warmup 20]

Config: size x size x 3 dense frames -> encoder size/2 x size/2 x 32 (radius 4) -> one hierarchy layer of the same grid
(radius 4, temporalHorizon 1), learning on everywhere. "device": frames already in HBM, ogn_ienc_step_device +
ogn_step_device queued back to back on one stream, one synchronise at the end of the timed region. "host": the reference's
calling sequence through host buffers (ogn_ienc_step, ogn_ienc_hidden_cs, ogn_step with a host CSDR), every call synchronous.
Both produce the same prediction CSDR (checked).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    args = ap.parse_args()
    import torch
    import ogmaneo2_b200 as og
    from ogmaneo2_b200.flat_api import InputType, LayerDesc
    from ogmaneo2_b200.image_encoder import IEncLib, ImageEncoder
    n = args.size
    hidden, visible = (n // 2, n // 2, 32), [((n, n, 3), 4)]
    lds = [LayerDesc(hiddenSize=hidden, ffRadius=4, pRadius=4, ticksPerUpdate=2, temporalHorizon=1)]
    total = args.warmup + args.steps
    rng = np.random.RandomState(5)
    base = rng.rand(n * n * 3).astype(np.float32)
    nframes = 32  # cycled
    frames = [np.clip(0.7 * np.roll(base, 3 * 7 * t) + 0.3 * rng.rand(n * n * 3).astype(np.float32), 0, 1).astype(np.float32) for t in range(nframes)]
    lib = IEncLib(og.cdll(), "ogn_")

    def make():
        h = og.Hierarchy([hidden], [InputType.prediction], lds, seed=11)
        e = ImageEncoder.createOn(lib, h, hidden, visible, seed=77)
        return h, e

    h, e = make()
    dframes = [torch.from_numpy(f).cuda() for f in frames]
    torch.cuda.synchronize()
    for t in range(args.warmup):
        e.stepDevice([dframes[t % nframes].data_ptr()], True)
        h.stepDevice([e.hiddenCsDevice()], True)
    h.synchronize()
    t0 = time.perf_counter()
    for t in range(args.warmup, total):
        e.stepDevice([dframes[t % nframes].data_ptr()], True)
        h.stepDevice([e.hiddenCsDevice()], True)
    h.synchronize()
    dev_s = time.perf_counter() - t0
    pred_dev = h.getPredictionCs(0).copy()

    h2, e2 = make()
    for t in range(args.warmup):
        e2.step([frames[t % nframes]], True)
        h2.step([e2.getHiddenCs()], True)
    t0 = time.perf_counter()
    for t in range(args.warmup, total):
        e2.step([frames[t % nframes]], True)
        h2.step([e2.getHiddenCs()], True)
    pred_host = h2.getPredictionCs(0).copy()
    host_s = time.perf_counter() - t0
    print(json.dumps({"metric": "image_to_prediction_steps_per_s_learn_on", "unit": "steps/s",
                      "config": {"workload": f"{n}x{n}x3 frames -> ImageEncoder {hidden[0]}x{hidden[1]}x32 r=4 -> 1 hierarchy layer {hidden[0]}x{hidden[1]}x32 r=4, TH=1"},
                      "device_resident": {"value": args.steps / dev_s, "ms_per_step": 1e3 * dev_s / args.steps,
                                          "api": "ogn_ienc_step_device + ogn_step_device on one stream, frames in HBM"},
                      "host_buffers": {"value": args.steps / host_s, "ms_per_step": 1e3 * host_s / args.steps,
                                       "api": "ogn_ienc_step + ogn_ienc_hidden_cs + ogn_step (host CSDR), synchronous calls"},
                      "same_prediction": bool(np.array_equal(pred_dev, pred_host)), "steps": args.steps, "warmup": args.warmup}))


if __name__ == "__main__":
    main()

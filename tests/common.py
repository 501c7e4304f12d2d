"""Shared helpers of the parity tests: drive two implementations of the flat hierarchy ABI on the same seeded inputs
and compare everything observable, bit for bit."""
import numpy as np

from ogmaneo2_b200 import streams
from ogmaneo2_b200.flat_api import FlatHierarchy


def make_inputs(cfg, t, action=None):
    ins = []
    for size, kind in zip(cfg["inputSizes"], cfg["streams"]):
        if kind == "action":
            ins.append(np.zeros(size[0] * size[1], np.int32) if action is None else action.copy())
        else:
            ins.append(streams.make(kind, t, size, seed=cfg.get("noise_seed", 7), shift=cfg.get("shift", 0)))
    return ins


def reward_at(t):
    return float(np.float32(np.sin(np.float32(0.1 * t))))


def bits(a):
    return np.ascontiguousarray(a).view(np.int32) if a.dtype == np.float32 else a


def diff_snapshots(a: dict, b: dict):
    """Returns a list of (key, first differing indices) — empty when bit-identical."""
    out = []
    if a.keys() != b.keys():
        out.append(("keys", sorted(set(a) ^ set(b))))
        return out
    for k in a:
        if a[k].shape != b[k].shape:
            out.append((k, ("shape", a[k].shape, b[k].shape)))
        elif not np.array_equal(bits(a[k]), bits(b[k])):
            idx = np.flatnonzero(bits(a[k]) != bits(b[k]))
            out.append((k, (int(idx.size), idx[:4].tolist(), a[k][idx[:4]].tolist(), b[k][idx[:4]].tolist())))
    return out


def run_pair(cfg, steps, mk_a, mk_b, check_every=1, weights_at_end=True, learn=True, compare_rng=False):
    """Steps implementation A (device under test) and B (checker) side by side. Returns None or a failure description."""
    ha, hb = mk_a(cfg), mk_b(cfg)
    d = diff_snapshots(ha.snapshot(True), hb.snapshot(True))
    if d:
        return f"{cfg['name']}: initial state differs: {d[:3]}"
    action = None
    for t in range(steps):
        ins = make_inputs(cfg, t, action)
        r = reward_at(t)
        ha.step(ins, learn, r, False)
        hb.step(ins, learn, r, False)
        if "action" in cfg["streams"]:
            i = cfg["streams"].index("action")
            action = hb.getPredictionCs(i)
            pa = ha.getPredictionCs(i)
            if not np.array_equal(action, pa):
                return f"{cfg['name']}: action differs at step {t}: {pa.tolist()} vs {action.tolist()}"
        last = t == steps - 1
        if t % check_every == 0 or last:
            d = diff_snapshots(ha.snapshot(last and weights_at_end), hb.snapshot(last and weights_at_end))
            if d:
                return f"{cfg['name']}: step {t}: {d[:3]}"
            if compare_rng and ha.rngPeek() != hb.rngPeek():
                return f"{cfg['name']}: step {t}: ComputeSystem::rng state differs"
    ha.close()
    hb.close()
    return None


def maker(flat):
    def mk(cfg):
        return FlatHierarchy(flat, cfg["inputSizes"], cfg["inputTypes"], cfg["layerDescs"], cfg["seed"])
    return mk


def c1_last1000_errors(h, cfg, steps=10000):
    """run `steps` steps of the sine stream with learning on and count, over the last 1000, the steps
    whose prediction differs from the input that actually came next. A functional known-answer (the hierarchy has learned the
    wave) on top of the bit-level parity tests."""
    from ogmaneo2_b200 import streams
    ins = lambda t: [streams.make(k, t, s) for s, k in zip(cfg["inputSizes"], cfg["streams"])]
    errs = 0
    for t in range(steps):
        h.step(ins(t), True)
        if t >= steps - 1000:
            errs += int((h.getPredictionCs(0) != ins(t + 1)[0]).sum())
    return errs

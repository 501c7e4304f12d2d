"""CPU tests (no GPU): the oracle port (oracle/sph_oracle.cpp) is pinned against
  (a) the committed golden fixtures generated from the unmodified reference (tests/golden/make_golden.py), and
  (b) the compiled reference itself (oracle/_ref), side by side, wherever /root/reference was available to build it:
      CSDRs, activations, weights, ComputeSystem RNG state and save files, bit for bit.
"""
import os
import sys
import zlib

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

import make_golden  # noqa: E402
from common import diff_snapshots, make_inputs, maker, run_pair  # noqa: E402
from ogmaneo2_b200 import configs  # noqa: E402
from ogmaneo2_b200.flat_api import FlatHierarchy  # noqa: E402
from oracle import loader  # noqa: E402

GOLDEN = make_golden.golden_configs()
HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return {k: z[k] for k in z.files}


def check_against_golden(flat, name):
    cfg, steps = GOLDEN[name]
    gold = load_golden(name)
    assert int(gold.pop("steps")[0]) == steps
    got = make_golden.run(flat, cfg, steps)
    assert got.keys() == gold.keys()
    for k in gold:
        a, b = got[k], gold[k]
        assert a.shape == b.shape, k
        if a.dtype == np.float32:
            a, b = a.view(np.int32), b.view(np.int32)
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)[0]
            pytest.fail(f"{name}: '{k}' differs from the reference fixture first at index {bad.tolist()}")


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_port_matches_golden(name):
    port = loader.port()
    port.set_num_threads(0)
    check_against_golden(port, name)


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_reference_still_matches_golden(name):
    """Guards the fixtures themselves: regenerating them here must give the same bytes (same libm, same libstdc++)."""
    ref = loader.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference sources on this box)")
    ref.set_num_threads(1)
    check_against_golden(ref, name)


REF_CASES = [("c1", configs.c1_sine(), 500), ("c2", configs.c2_video(12, 3), 50), ("c3", configs.c3_actor(8), 250),
             ("c4", configs.c4_large(24), 25), ("c5", configs.c5_unit(16), 30)]


@pytest.mark.parametrize("name,cfg,steps", REF_CASES, ids=[c[0] for c in REF_CASES])
def test_port_vs_reference_side_by_side(name, cfg, steps):
    ref = loader.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference sources on this box)")
    ref.set_num_threads(1)
    port = loader.port()
    port.set_num_threads(0)
    err = run_pair(cfg, steps, maker(port), maker(ref), check_every=1 if steps <= 60 else 7, compare_rng=True)
    assert err is None, err


def test_c1_learns_the_sine_wave():
    """Config 1 at its contract length (10 000 steps): the reference mispredicts 19 of the last 1000 steps; the port must
    give the same count (and the CUDA path too: tests/test_gpu_parity.py::test_c1_contract_length_known_answer)."""
    from common import c1_last1000_errors
    cfg = configs.c1_sine()
    port = loader.port()
    port.set_num_threads(1)
    got = c1_last1000_errors(maker(port)(cfg), cfg)
    ref = loader.ref()
    if ref is not None:
        ref.set_num_threads(1)
        assert got == c1_last1000_errors(maker(ref)(cfg), cfg)
    assert got == 19


def _mask_pad(b: bytes) -> bytes:
    return b  # Int3.pad is zeroed by both writers here (the reference leaves it uninitialised: SURVEY.md D1)


@pytest.mark.parametrize("cfg", [configs.c1_sine(), configs.c3_actor(8), configs.c2_video(6, 2)], ids=["c1", "c3_actor", "c2"])
def test_save_files_identical_and_cross_loadable(cfg, tmp_path):
    """Save after N steps: same bytes from port and reference; each loads the other's file and continues identically."""
    ref = loader.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference sources on this box)")
    ref.set_num_threads(1)
    port = loader.port()
    a, b = maker(port)(cfg), maker(ref)(cfg)
    action = None
    for t in range(23):
        ins = make_inputs(cfg, t, action)
        a.step(ins, True, 0.25)
        b.step(ins, True, 0.25)
        if "action" in cfg["streams"]:
            action = b.getPredictionCs(cfg["streams"].index("action"))
    fa, fb = str(tmp_path / "port.ohr"), str(tmp_path / "ref.ohr")
    na, nb = a.save(fa), b.save(fb)
    assert na == nb
    assert zlib.crc32(open(fa, "rb").read()) == zlib.crc32(open(fb, "rb").read()), "save files differ"
    # cross-load and continue (RNG reseeded identically on both sides)
    a2 = FlatHierarchy.load(port, fb, cfg["inputSizes"], seed=77)
    b2 = FlatHierarchy.load(ref, fa, cfg["inputSizes"], seed=77)
    assert not diff_snapshots(a2.snapshot(True), b2.snapshot(True))
    for t in range(23, 40):
        ins = make_inputs(cfg, t, action)
        a2.step(ins, True, 0.5)
        b2.step(ins, True, 0.5)
        if "action" in cfg["streams"]:
            action = b2.getPredictionCs(cfg["streams"].index("action"))
    d = diff_snapshots(a2.snapshot(True), b2.snapshot(True))
    assert not d, d[:3]


def test_port_save_load_roundtrip(tmp_path):
    port = loader.port()
    cfg = configs.c2_video(6, 2)
    a = maker(port)(cfg)
    for t in range(9):
        a.step(make_inputs(cfg, t), True)
    f = str(tmp_path / "x.ohr")
    a.save(f)
    b = FlatHierarchy.load(port, f, cfg["inputSizes"])
    assert not diff_snapshots(a.snapshot(True), b.snapshot(True))
    f2 = str(tmp_path / "y.ohr")
    b.save(f2)
    assert open(f, "rb").read() == open(f2, "rb").read()


def test_port_is_thread_count_independent():
    port = loader.port()
    cfg = configs.c3_actor(8)
    outs = []
    for n in (1, 3, 0):
        port.set_num_threads(n)
        h = maker(port)(cfg)
        action = None
        for t in range(60):
            h.step(make_inputs(cfg, t, action), True, 0.1 * t)
            action = h.getPredictionCs(1)
        outs.append(h.snapshot(True))
        outs[-1]["rng"] = np.array([h.rngPeek()], dtype=np.uint32)
    port.set_num_threads(0)
    assert not diff_snapshots(outs[0], outs[1]) and not diff_snapshots(outs[0], outs[2])


@pytest.mark.parametrize("seed", range(16))
def test_port_vs_reference_on_random_shapes(seed):
    """Seeded random geometries (non-square, Z off the powers of two, radius 0 and radius wider than the grid, several
    inputs, temporal horizons above the tick rate): the port must track the compiled reference bit for bit."""
    from random_cfgs import random_cfg
    ref = loader.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference sources on this box)")
    ref.set_num_threads(1)
    port = loader.port()
    port.set_num_threads(0)
    err = run_pair(random_cfg(seed), 24, maker(port), maker(ref), check_every=4, compare_rng=True)
    assert err is None, err


def state_rewind_check(mk_a, mk_b, cfg, label):
    """Hierarchy::getState / setState (Hierarchy.cpp:434-493) through the flat ABI on two implementations side by side:
    snapshot after some learning, run on without learning, rewind both, replay: the replay must reproduce the first pass,
    and learning steps after the rewind must keep both implementations bit-identical (weights included)."""
    a, b = mk_a(cfg), mk_b(cfg)
    for t in range(17):
        ins = make_inputs(cfg, t)
        a.step(ins, True)
        b.step(ins, True)
    sa, sb = a.getState(), b.getState()
    first = []
    for t in range(17, 30):
        ins = make_inputs(cfg, t)
        a.step(ins, False)
        b.step(ins, False)
        first.append((a.getPredictionCs(0), [a.scHiddenCs(l) for l in range(a.getNumLayers())]))
        assert np.array_equal(first[-1][0], b.getPredictionCs(0)), f"{label}: prediction differs at step {t}"
    a.setState(sa)
    b.setState(sb)
    d = diff_snapshots(a.snapshot(False), b.snapshot(False))
    assert not d, f"{label}: state differs right after setState: {d[:3]}"
    for k, t in enumerate(range(17, 30)):
        ins = make_inputs(cfg, t)
        a.step(ins, False)
        b.step(ins, False)
        assert np.array_equal(a.getPredictionCs(0), first[k][0]), f"{label}: replay after setState differs at step {t}"
        for l in range(a.getNumLayers()):
            assert np.array_equal(a.scHiddenCs(l), first[k][1][l]), f"{label}: replayed hidden CSDR {l} differs at step {t}"
    a.setState(sa)
    b.setState(sb)
    for t in range(30, 42):  # learning after a rewind: both sides learn from the same (stale) activations
        ins = make_inputs(cfg, t)
        a.step(ins, True)
        b.step(ins, True)
    d = diff_snapshots(a.snapshot(True), b.snapshot(True))
    assert not d, f"{label}: {d[:3]}"


def test_port_get_set_state_like_the_reference():
    ref = loader.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference sources on this box)")
    ref.set_num_threads(1)
    state_rewind_check(maker(loader.port()), maker(ref), configs.c2_video(8, 3), "port vs reference")


def test_port_layer_getters_like_the_reference():
    """The shape / hyper-parameter getters of the flat ABI (layer_hidden_size, layer_visible_desc, get_*_alpha,
    get_actor_params) answer the same on the port and on the compiled reference, before and after setters."""
    ref = loader.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference sources on this box)")
    cfg = configs.c3_actor(8)
    a, b = maker(loader.port())(cfg), maker(ref)(cfg)
    for h in (a, b):
        h.setSCAlpha(1, 0.25)
        h.setPredAlpha(1, 0, 0.125)
        h.setActorParams(1, 0.03, 0.015, 0.97, 5, 6)
    def facts(h):
        out = []
        for l in range(h.getNumLayers()):
            out.append((h.getSCAlpha(l), h.layerHiddenSize(0, l), [h.layerVisibleDesc(0, l, v) for v in range(h.scNumVisible(l))]))
            for p in range(h.predCount(l)):
                if h.predExists(l, p):
                    out.append((h.getPredAlpha(l, p), h.layerHiddenSize(1 + p, l),
                                [h.layerVisibleDesc(1 + p, l, v) for v in range(h.predNumVisible(l, p))]))
        out.append((sorted(h.getActorParams(1).items()), h.layerHiddenSize(-2), [h.layerVisibleDesc(-2, 0, v) for v in range(h.actorNumVisible(1))]))
        return out
    fa, fb = facts(a), facts(b)
    assert fa == fb
    assert fa[1][0] == 0.25 and dict(fa[-1][0])["historyIters"] == 6 and fa[-1][1] == (2, 2, 4)

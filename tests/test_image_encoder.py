"""ImageEncoder: the oracle port against the compiled reference on the CPU, and the CUDA path against
both on the GPU, bit for bit: hidden CSDRs, resources, weights, reconstructions and save files."""
import ctypes
import os
import zlib

import numpy as np
import pytest

from common import diff_snapshots
from ogmaneo2_b200.image_encoder import IEncLib, ImageEncoder
from oracle import loader

CASES = {
    "rgb": dict(hidden=(6, 5, 16), visible=[((12, 10, 3), 2)]),                      # downsampling image patch encoder
    "two_layers": dict(hidden=(5, 5, 8), visible=[((5, 5, 4), 1), ((9, 7, 2), 3)]),  # two inputs, clamped fields
    "z32": dict(hidden=(4, 6, 32), visible=[((8, 6, 1), 2)]),                        # 32 cells per column (std::sort past its insertion-sort size)
    "odd": dict(hidden=(3, 4, 5), visible=[((7, 3, 2), 1)]),                         # Z off the powers of two
}


def dense_input(case, t, v):
    s = case["visible"][v][0]
    n = s[0] * s[1] * s[2]
    rng = np.random.RandomState(1000 * t + 17 * v + 3)
    base = 0.5 + 0.5 * np.sin(0.37 * np.arange(n, dtype=np.float32) + 0.11 * t).astype(np.float32)
    return (0.8 * base + 0.2 * rng.rand(n).astype(np.float32)).astype(np.float32)


def port_lib():
    loader.port()
    return IEncLib(ctypes.CDLL(loader.PORT_SO), "oport_")


def ref_lib():
    if loader.ref() is None:
        return None
    return IEncLib(ctypes.CDLL(loader.REF_SO), "oref_")


def run_pair(case, la, lb, steps=12, tmp_path=None):
    a = ImageEncoder(la, case["hidden"], case["visible"], seed=77)
    b = ImageEncoder(lb, case["hidden"], case["visible"], seed=77)
    d = diff_snapshots(a.snapshot(), b.snapshot())
    assert not d, f"initial state differs: {d[:3]}"
    for t in range(steps):
        ins = [dense_input(case, t, v) for v in range(len(case["visible"]))]
        learn = t % 4 != 3
        a.step(ins, learn)
        b.step(ins, learn)
        if t % 3 == 2:
            hc = b.getHiddenCs()
            a.reconstruct(hc)
            b.reconstruct(hc)
        d = diff_snapshots(a.snapshot(), b.snapshot())
        assert not d, f"step {t}: {d[:3]}"
    if tmp_path is not None:
        fa, fb = str(tmp_path / "a.ienc"), str(tmp_path / "b.ienc")
        assert a.save(fa) == b.save(fb)
        assert zlib.crc32(open(fa, "rb").read()) == zlib.crc32(open(fb, "rb").read()), "save files differ"
        a2 = ImageEncoder.load(la, fb, case["hidden"], case["visible"])
        b2 = ImageEncoder.load(lb, fa, case["hidden"], case["visible"])
        for t in range(steps, steps + 4):
            ins = [dense_input(case, t, v) for v in range(len(case["visible"]))]
            a2.step(ins, True)
            b2.step(ins, True)
        d = diff_snapshots(a2.snapshot(), b2.snapshot())
        assert not d, f"after cross-load: {d[:3]}"


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_reference(name, tmp_path):
    ref = ref_lib()
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference sources on this box)")
    run_pair(CASES[name], port_lib(), ref, tmp_path=tmp_path)


def test_port_matches_golden():
    """tests/golden/image_encoder.npz: trajectory of the compiled reference on the `rgb` case (make_golden.py)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "image_encoder.npz")
    gold = np.load(path)
    got = golden_run(port_lib())
    for k in gold.files:
        a, b = got[k], gold[k]
        if a.dtype == np.float32:
            a, b = a.view(np.int32), b.view(np.int32)
        assert np.array_equal(a, b), f"'{k}' differs from the reference fixture"


def golden_run(lib):
    case = CASES["rgb"]
    e = ImageEncoder(lib, case["hidden"], case["visible"], seed=77)
    cs = []
    for t in range(10):
        e.step([dense_input(case, t, 0)], True)
        cs.append(e.getHiddenCs())
    e.reconstruct(cs[-1])
    out = e.snapshot()
    out["hiddenCs_per_step"] = np.stack(cs)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_matches_checker(product, name, tmp_path):
    dev = IEncLib(product.cdll(), "ogn_")
    chk = ref_lib() or port_lib()
    run_pair(CASES[name], dev, chk, tmp_path=tmp_path)


@pytest.mark.gpu
def test_device_matches_golden(product):
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "image_encoder.npz")
    gold = np.load(path)
    got = golden_run(IEncLib(product.cdll(), "ogn_"))
    for k in gold.files:
        a, b = got[k], gold[k]
        if a.dtype == np.float32:
            a, b = a.view(np.int32), b.view(np.int32)
        assert np.array_equal(a, b), f"'{k}' differs from the reference fixture"

"""The glibc expf/tanhf replay (ogmaneo2_b200/csrc/ogn_libm.h) against the box's libm. CPU: every 13th float bit pattern
(3.3e8 inputs) by default, all 2^32 with OGN_LIBM_EXHAUSTIVE=1. GPU (-m gpu): the same header compiled for the device
against host libm on 2^24 samples spread over all exponents."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_replay_matches_libm_on_cpu(tmp_path):
    exe = str(tmp_path / "libm_check")
    subprocess.check_call(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-mfma",
                           "-o", exe, os.path.join(HERE, "csrc", "libm_check.c"), "-lm"])
    stride = "1" if os.environ.get("OGN_LIBM_EXHAUSTIVE") else "13"
    be, bt, n = (int(v) for v in subprocess.check_output([exe, stride]).split())
    assert n >= (1 << 32) // 13
    assert be == 0, f"expf replay differs from libm on {be} of {n} inputs"
    assert bt == 0, f"tanhf replay differs from libm on {bt} of {n} inputs"


@pytest.mark.gpu
def test_replay_matches_libm_on_device(product):
    import torch
    lib = product.cdll()
    n = 1 << 24
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    # make sure the ranges the learn rules actually hit are dense: |x| < 32
    bits[: n // 2] = (rng.uniform(-32, 32, size=n // 2).astype(np.float32)).view(np.uint32)
    x = bits.view(np.float32)
    xd = torch.from_numpy(x.copy()).cuda()
    libm = ctypes.CDLL("libm.so.6")
    for name, fn in (("expf", lib.ogn_test_expf), ("tanhf", lib.ogn_test_tanhf)):
        out = torch.empty_like(xd)
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        fn.restype = ctypes.c_int
        assert fn(xd.data_ptr(), out.data_ptr(), n, None) == 0
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        with np.errstate(all="ignore"):
            f = getattr(libm, name)
            f.restype, f.argtypes = ctypes.c_float, [ctypes.c_float]
            # libm through numpy: np.exp / np.tanh on float32 do not call libm's scalar routines, so sample via ctypes
            idx = rng.integers(0, n, size=200000)
            want = np.array([f(float(v)) for v in x[idx]], dtype=np.float32)
        g = got[idx]
        same = (g.view(np.uint32) == want.view(np.uint32)) | (np.isnan(g) & np.isnan(want))
        assert same.all(), f"device {name} differs from libm at x={x[idx][~same][:5]}"

"""Regenerates tests/golden/image_encoder.npz from the UNMODIFIED reference (oracle/_ref, compiled from /root/reference by
oracle/Makefile): the ImageEncoder trajectory tests/test_image_encoder.py::golden_run produces on the `rgb` case.

    python tests/golden/make_golden_ienc.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import test_image_encoder as t  # noqa: E402

if __name__ == "__main__":
    ref = t.ref_lib()
    if ref is None:
        raise SystemExit("oracle/_ref is not built: the fixtures can only be generated where /root/reference is present")
    out = t.golden_run(ref)
    np.savez_compressed(os.path.join(HERE, "image_encoder.npz"), **out)
    print({k: v.shape for k, v in out.items()})

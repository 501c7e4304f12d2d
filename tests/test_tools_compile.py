"""The stand-alone CUDA tool under tools/ cross-compiles for sm_100a (no GPU needed): tools/microbench_patterns.cu is what
profiles/r2_microbench_patterns.jsonl was produced with."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not installed")
def test_microbench_cross_compiles(tmp_path):
    out = str(tmp_path / "mb.o")
    r = subprocess.run([NVCC, "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-c", os.path.join(ROOT, "tools", "microbench_patterns.cu"), "-o", out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and os.path.getsize(out) > 0, r.stderr[-2000:]

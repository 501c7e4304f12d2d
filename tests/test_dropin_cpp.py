"""The C++ drop-in surface (include/ogmaneo/): a program written only against the reference's public API
(tests/csrc/dropin_demo.cpp) compiles with -std=c++14 against our headers, links libogn_b200.so, and on a GPU prints
exactly what the same source prints when built against the unmodified reference (tests/golden/dropin_demo.txt)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "csrc", "dropin_demo.cpp")
GOLDEN = os.path.join(ROOT, "tests", "golden", "dropin_demo.txt")


def build_product_demo(product, tmp_path):
    product.cdll()  # makes sure the library exists
    exe = str(tmp_path / "dropin_demo")
    libdir = os.path.join(ROOT, "ogmaneo2_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-Wall", f"-I{ROOT}/include", SRC, f"-L{libdir}", "-logn_b200",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    return exe


def test_reference_style_program_compiles_and_links(product, tmp_path):
    assert os.path.exists(build_product_demo(product, tmp_path))


def test_golden_is_what_the_reference_prints(tmp_path):
    ref = os.environ.get("OGN_REF_DIR", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "source", "ogmaneo")):
        pytest.skip("reference sources not present on this box")
    exe = str(tmp_path / "ref_demo")
    srcs = [os.path.join(ref, "source", "ogmaneo", f) for f in sorted(os.listdir(os.path.join(ref, "source", "ogmaneo"))) if f.endswith(".cpp")]
    subprocess.check_call(["g++", "-O3", "-std=c++14", "-fopenmp", "-ffp-contract=off", f"-I{ref}/source", SRC] + srcs + ["-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert out.returncode == 0
    assert out.stdout == open(GOLDEN).read()


@pytest.mark.gpu
def test_product_build_prints_the_reference_output(product, tmp_path):
    exe = build_product_demo(product, tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    want, got = open(GOLDEN).read().splitlines(), out.stdout.splitlines()
    assert got == want, "\n".join(f"{a!r} != {b!r}" for a, b in zip(got, want) if a != b)[:2000]

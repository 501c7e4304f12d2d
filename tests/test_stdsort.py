"""csrc/ogn_stdsort.h (the replay of libstdc++'s std::sort that the ImageEncoder kernel uses for tied activations) against
the real std::sort, on the host: tests/csrc/stdsort_check.cpp is compiled with g++ and run on tie-heavy random arrays of
every length up to 64 plus adversarial inputs built with McIlroy's killer comparator (they reach the depth limit and the heap-sort fallback)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_replay_matches_std_sort(tmp_path):
    exe = str(tmp_path / "stdsort_check")
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(HERE, "csrc", "stdsort_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, "400000"], capture_output=True, text=True, timeout=300)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0 and out.stdout.startswith("ok "), out.stdout[-400:]
    assert int(out.stdout.split()[-1]) > 0, "no input reached the heap-sort fallback"

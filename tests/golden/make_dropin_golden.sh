#!/bin/sh
# Regenerates tests/golden/dropin_demo.txt: tests/csrc/dropin_demo.cpp compiled against the UNMODIFIED reference sources
# where they lie (/root/reference), run on the CPU. The product build of the same source must print the same lines.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${OGN_REF_DIR:-/root/reference}
OUT=${TMPDIR:-/tmp}/ogn_dropin_ref_demo
g++ -O3 -std=c++14 -fopenmp -ffp-contract=off -I"$REF/source" "$HERE/../csrc/dropin_demo.cpp" "$REF"/source/ogmaneo/*.cpp -o "$OUT"
"$OUT" > "$HERE/dropin_demo.txt"
echo "wrote $HERE/dropin_demo.txt"

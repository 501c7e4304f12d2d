// dropin_demo.cpp — a program written ONLY against the reference's public C++ API (<ogmaneo/Hierarchy.h>): the same
// source is compiled (a) against /root/reference/source with the reference's own .cpp files (CPU, OpenMP) to produce
// tests/golden/dropin_demo.txt, and (b) against include/ + libogn_b200.so (the product). Both must print the same lines.
// Covers: ComputeSystem + rng seeding, LayerDesc defaults, initRandom, step with and without learning, getPredictionCs,
// getSCLayer / getPLayers getters, Hierarchy copy construction and assignment, getState / setState, writeToStream /
// readFromStream.
#include <ogmaneo/Hierarchy.h>

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <vector>

using namespace ogmaneo;

static uint64_t fnv(const void *p, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char *b = static_cast<const unsigned char *>(p);
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
static uint64_t hashCs(const IntBuffer &b) { return fnv(b.data(), b.size() * sizeof(int)); }

static IntBuffer input(int t, int X, int Y, int Z) {
    IntBuffer b((size_t)X * Y);
    for (int x = 0; x < X; x++)
        for (int y = 0; y < Y; y++) b[y + x * Y] = ((3 * (x + t) + 5 * (y + 2 * t)) / 4 + (t % 7 == 3 ? x * y : 0)) % Z;
    return b;
}

static uint64_t hashHierarchy(const Hierarchy &h) {
    uint64_t v = 0;
    for (int l = 0; l < h.getNumLayers(); l++) {
        v = fnv(h.getSCLayer(l).getHiddenCs().data(), h.getSCLayer(l).getHiddenCs().size() * sizeof(int), v ? v : 1469598103934665603ull);
        for (size_t p = 0; p < h.getPLayers(l).size(); p++)
            if (h.getPLayers(l)[p]) v = fnv(h.getPLayers(l)[p]->getHiddenCs().data(), h.getPLayers(l)[p]->getHiddenCs().size() * sizeof(int), v);
    }
    return v;
}

int main() {
    ComputeSystem::setNumThreads(2);
    ComputeSystem cs;
    cs.rng.seed(4321);
    const int X = 10, Y = 7, Z = 16;
    std::vector<Hierarchy::LayerDesc> lds(3);
    lds[0].hiddenSize = Int3(6, 6, 32); lds[0].ffRadius = 3; lds[0].pRadius = 3;
    lds[1].hiddenSize = Int3(5, 5, 16);                       // defaults: radius 2, ticksPerUpdate 2, temporalHorizon 4
    lds[2].hiddenSize = Int3(4, 4, 16); lds[2].temporalHorizon = 2;
    Hierarchy h;
    h.initRandom(cs, {Int3(X, Y, Z)}, {InputType::prediction}, lds);
    std::printf("layers %d ticksPerUpdate %d %d %d\n", h.getNumLayers(), h.getTicksPerUpdate(0), h.getTicksPerUpdate(1), h.getTicksPerUpdate(2));

    for (int t = 0; t < 40; t++) {
        IntBuffer in = input(t, X, Y, Z);
        h.step(cs, {&in}, t % 9 != 4);
        if (t % 4 == 3)
            std::printf("t %2d pred %016llx state %016llx ticks %d %d upd %d%d%d\n", t, (unsigned long long)hashCs(h.getPredictionCs(0)),
                        (unsigned long long)hashHierarchy(h), h.getTicks(1), h.getTicks(2), (int)h.getUpdate(0), (int)h.getUpdate(1),
                        (int)h.getUpdate(2));
    }

    // deep copy, then both continue identically; assignment over an existing hierarchy as well
    Hierarchy c(h);
    Hierarchy a;
    a = h;
    State st;
    h.getState(st);
    for (int t = 40; t < 52; t++) {
        IntBuffer in = input(t, X, Y, Z);
        h.step(cs, {&in}, true);
        c.step(cs, {&in}, true);
        a.step(cs, {&in}, true);
    }
    std::printf("copy %d assign %d\n", (int)(hashHierarchy(c) == hashHierarchy(h)), (int)(hashHierarchy(a) == hashHierarchy(h)));
    std::printf("t 51 pred %016llx state %016llx\n", (unsigned long long)hashCs(h.getPredictionCs(0)), (unsigned long long)hashHierarchy(h));

    // setState rewinds the CSDR state (weights keep what they learned since): hash after replaying the same 12 inputs
    h.setState(st);
    for (int t = 40; t < 52; t++) {
        IntBuffer in = input(t, X, Y, Z);
        h.step(cs, {&in}, false);
    }
    std::printf("rewound pred %016llx state %016llx\n", (unsigned long long)hashCs(h.getPredictionCs(0)), (unsigned long long)hashHierarchy(h));

    // save / load through streams: same bytes, and the loaded hierarchy continues like the saved one
    std::stringstream ss(std::ios::in | std::ios::out | std::ios::binary);
    h.writeToStream(ss);
    const std::string bytes = ss.str();
    // Int3 carries 4 uninitialised pad bytes per instance in the reference: hash the size only
    std::printf("file %zu bytes\n", bytes.size());
    Hierarchy r;
    r.readFromStream(ss);
    for (int t = 52; t < 60; t++) {
        IntBuffer in = input(t, X, Y, Z);
        h.step(cs, {&in}, true);
        r.step(cs, {&in}, true);
    }
    std::printf("reloaded %d pred %016llx\n", (int)(hashHierarchy(r) == hashHierarchy(h)), (unsigned long long)hashCs(r.getPredictionCs(0)));
    const SparseCoder::VisibleLayer &vl = h.getSCLayer(0).getVisibleLayer(0);
    std::printf("weights %zu hash %016llx\n", vl.weights.nonZeroValues.size(),
                (unsigned long long)fnv(vl.weights.nonZeroValues.data(), vl.weights.nonZeroValues.size() * sizeof(float)));
    return 0;
}

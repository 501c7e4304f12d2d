// oracle/sph_oracle.cpp — TEST INFRASTRUCTURE ONLY. Not part of the product; nothing under ogmaneo2_b200/ may
// include, link or load it. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs use it, and only as the checker / the timed CPU baseline.
//
// A CPU restatement of the reference's per-step Sparse Predictive Hierarchy update (ogmacorp/OgmaNeo2,
// source/ogmaneo/{Hierarchy,SparseCoder,Predictor,Actor,Helpers,SparseMatrix}.cpp), exported through the flat
// hierarchy C ABI (include/ogn_hierarchy_api.h) with the prefix oport_. Each function cites the reference
// file:line it follows. It is written from the algorithm, not from the code: the reference's CSR index arrays
// (columnIndices, rowRanges, the transpose index) are replaced by closed-form receptive-field geometry and are
// only regenerated when a save file is written.
//
// Language: C++17 rather than plain C because bit-identical initial weights and Actor sampling require the very
// libstdc++ <random> classes the reference calls (std::mt19937, std::uniform_real_distribution<float>,
// std::uniform_int_distribution<int>; SURVEY.md §8c: "do not re-implement the distributions").
//
// PARITY PINNING: the reference ships no tests, golden vectors or fixtures for this path, so the
// oracle is pinned against outputs of the reference itself: tests/test_oracle.py steps this port and
// oracle/_ref (the unmodified reference sources compiled here) side by side on all BASELINE configs (reduced
// sizes) and demands bit-equal CSDRs, activations, weights, RNG state and save files; tests/golden/*.npz holds
// reference-generated trajectories (made by tests/golden/make_golden.py) so the pin also holds where
// /root/reference is absent.
//
// Determinism: every runKernelN call site of the reference draws one seed per tile from cs.rng
// (Helpers.cpp:22-31, 46-61). This port draws them serially in tile order, which equals the reference at
// OMP_NUM_THREADS=1 (with more threads the reference itself races on cs.rng), and parallelises the
// per-column bodies with OpenMP, which is race-free because the bodies never write shared state.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <random>
#include <string>
#include <vector>
#include <omp.h>

#include "../include/ogn_hierarchy_api.h"

namespace {

// The port keeps its own thread count: omp_set_num_threads() is process-global and would also change the thread
// count of oracle/_ref when both are loaded (the reference must run single-threaded whenever RNG state is compared).
int g_portThreads = 0;
inline int portThreads() { return g_portThreads > 0 ? g_portThreads : omp_get_num_procs(); }

typedef std::vector<int> IntBuf;
typedef std::vector<float> FloatBuf;

// ---------------------------------------------------------------------------------------------------------------
// Receptive-field geometry: the pattern initSMLocalRF builds (Helpers.cpp:236-313) in closed form.
// Output ("hidden") grid (Ox,Oy,Oz) looks onto input ("visible") grid (Ix,Iy,Iz) with radius r.
// Value order (== the reference's nonZeroValues order): column (ox,oy) -> oz -> ix -> iy -> iz.
struct Geom {
    int Ox = 0, Oy = 0, Oz = 0, Ix = 0, Iy = 0, Iz = 0, r = 0;
    std::vector<int> lox, nx, loy, ny;         // clamped field start / extent per output x / y
    std::vector<int> rlox, rhix, rloy, rhiy;   // first / last output x (y) whose field covers input x (y); empty: lo > hi
    std::vector<int64_t> colBase;              // offset of row (ox,oy,0) in the value array
    int64_t nnz = 0;

    void init(int ix, int iy, int iz, int ox, int oy, int oz, int radius) {
        Ix = ix; Iy = iy; Iz = iz; Ox = ox; Oy = oy; Oz = oz; r = radius;
        // project(), Helpers.h:223-228: (int)(pos * (in/out as float) + 0.5f)
        float sx = (float)Ix / (float)Ox, sy = (float)Iy / (float)Oy;
        lox.resize(Ox); nx.resize(Ox); loy.resize(Oy); ny.resize(Oy);
        for (int x = 0; x < Ox; x++) {
            int c = (int)(x * sx + 0.5f);
            int lo = std::max(0, c - r), hi = std::min(Ix - 1, c + r);
            lox[x] = lo; nx[x] = std::max(0, hi - lo + 1);
        }
        for (int y = 0; y < Oy; y++) {
            int c = (int)(y * sy + 0.5f);
            int lo = std::max(0, c - r), hi = std::min(Iy - 1, c + r);
            loy[y] = lo; ny[y] = std::max(0, hi - lo + 1);
        }
        colBase.resize((size_t)Ox * Oy);
        int64_t off = 0;
        for (int x = 0; x < Ox; x++)
            for (int y = 0; y < Oy; y++) {
                colBase[(size_t)y + (size_t)x * Oy] = off;
                off += (int64_t)nx[x] * ny[y] * Iz * Oz;
            }
        nnz = off;
        rlox.assign(Ix, Ox); rhix.assign(Ix, -1); rloy.assign(Iy, Oy); rhiy.assign(Iy, -1);
        for (int x = 0; x < Ox; x++)
            for (int i = lox[x]; i < lox[x] + nx[x]; i++) { rlox[i] = std::min(rlox[i], x); rhix[i] = std::max(rhix[i], x); }
        for (int y = 0; y < Oy; y++)
            for (int i = loy[y]; i < loy[y] + ny[y]; i++) { rloy[i] = std::min(rloy[i], y); rhiy[i] = std::max(rhiy[i], y); }
    }
    int slots(int ox, int oy) const { return nx[ox] * ny[oy]; }
    // index of W[(ox,oy,oz)][slot s][iz]
    int64_t at(int ox, int oy, int oz, int s, int iz) const {
        return colBase[(size_t)oy + (size_t)ox * Oy] + ((int64_t)oz * slots(ox, oy) + s) * Iz + iz;
    }
    bool covers(int ox, int oy, int ix, int iy) const {
        return ix >= lox[ox] && ix < lox[ox] + nx[ox] && iy >= loy[oy] && iy < loy[oy] + ny[oy];
    }
    int slotOf(int ox, int oy, int ix, int iy) const { return (ix - lox[ox]) * ny[oy] + (iy - loy[oy]); }
};

// SparseMatrix::multiplyOHVs (SparseMatrix.cpp:260-276): sum over the field of W[row][slot][cs[visible column]]
inline float gatherRow(const Geom &g, const FloatBuf &w, const IntBuf &cs, int ox, int oy, int oz) {
    float sum = 0.0f;
    const int n = g.slots(ox, oy);
    const float *row = w.data() + g.at(ox, oy, oz, 0, 0);
    const int x0 = g.lox[ox], y0 = g.loy[oy], nyv = g.ny[oy];
    for (int s = 0; s < n; s++) {
        int ix = x0 + s / nyv, iy = y0 + s % nyv;
        sum += row[(int64_t)s * g.Iz + cs[iy + ix * g.Iy]];
    }
    return sum;
}
// SparseMatrix::deltaOHVs (SparseMatrix.cpp:488-501)
inline void deltaRow(const Geom &g, FloatBuf &w, const IntBuf &cs, float delta, int ox, int oy, int oz) {
    const int n = g.slots(ox, oy);
    float *row = w.data() + g.at(ox, oy, oz, 0, 0);
    const int x0 = g.lox[ox], y0 = g.loy[oy], nyv = g.ny[oy];
    for (int s = 0; s < n; s++) {
        int ix = x0 + s / nyv, iy = y0 + s % nyv;
        row[(int64_t)s * g.Iz + cs[iy + ix * g.Iy]] += delta;
    }
}

// runKernel1 / runKernel2 RNG consumption (Helpers.cpp:15-72): one uniform_int<int>(0,999999) draw per tile.
struct Rng {
    std::mt19937 rng;
    int batch1 = 512, batch2x = 2, batch2y = 2; // ComputeSystem defaults, ComputeSystem.h:27-31
    void consume1(int size) {
        std::uniform_int_distribution<int> seedDist(0, 999999);
        int batches = (size + batch1 - 1) / batch1;
        for (int i = 0; i < batches; i++) (void)seedDist(rng);
    }
    void consume2(int X, int Y, std::vector<int> *seeds = nullptr) {
        std::uniform_int_distribution<int> seedDist(0, 999999);
        int bx = (X + batch2x - 1) / batch2x, by = (Y + batch2y - 1) / batch2y;
        if (seeds) seeds->resize((size_t)bx * by);
        for (int i = 0; i < bx * by; i++) {
            int s = seedDist(rng);
            if (seeds) (*seeds)[i] = s;
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
struct VLD { int x, y, z, radius; };

// SparseCoder (SparseCoder.h:15-151, SparseCoder.cpp)
struct SparseCoder {
    int Hx = 0, Hy = 0, Hz = 0;
    float alpha = 0.1f; // SparseCoder.h:86-90
    IntBuf hiddenCs, hiddenCsPrev;
    struct VL { Geom g; FloatBuf w; FloatBuf recon; };
    std::vector<VL> vls;
    std::vector<VLD> vlds;

    void allocate(int hx, int hy, int hz, const std::vector<VLD> &d) {
        Hx = hx; Hy = hy; Hz = hz; vlds = d; vls.resize(d.size());
        for (size_t i = 0; i < d.size(); i++) {
            vls[i].g.init(d[i].x, d[i].y, d[i].z, Hx, Hy, Hz, d[i].radius);
            vls[i].w.assign((size_t)vls[i].g.nnz, 0.0f);
            vls[i].recon.assign((size_t)d[i].x * d[i].y * d[i].z, 0.0f);
        }
        hiddenCs.assign((size_t)Hx * Hy, 0); hiddenCsPrev.assign((size_t)Hx * Hy, 0);
    }
    // SparseCoder::initRandom, SparseCoder.cpp:85-125: weights U(-1,0) drawn sequentially in value order
    void initRandom(Rng &cs, int hx, int hy, int hz, const std::vector<VLD> &d) {
        allocate(hx, hy, hz, d);
        std::uniform_real_distribution<float> weightDist(-1.0f, 0.0f);
        for (auto &vl : vls)
            for (auto &v : vl.w) v = weightDist(cs.rng);
    }
    // SparseCoder::forward, SparseCoder.cpp:13-43
    void forward(const std::vector<const IntBuf *> &in) {
#pragma omp parallel for collapse(2) schedule(static) num_threads(portThreads())
        for (int ox = 0; ox < Hx; ox++)
            for (int oy = 0; oy < Hy; oy++) {
                int maxIndex = 0; float maxAct = -999999.0f;
                for (int hc = 0; hc < Hz; hc++) {
                    float sum = 0.0f;
                    for (size_t v = 0; v < vls.size(); v++) sum += gatherRow(vls[v].g, vls[v].w, *in[v], ox, oy, hc);
                    if (sum > maxAct) { maxAct = sum; maxIndex = hc; }
                }
                hiddenCs[oy + ox * Hy] = maxIndex;
            }
    }
    // SparseCoder::learn, SparseCoder.cpp:45-83 (+ multiplyOHVsT :278-294, countT :215-221, deltaChangedOHVsT :572-590)
    void learn(const IntBuf &in, int vli) {
        VL &vl = vls[vli]; const Geom &g = vl.g;
#pragma omp parallel for collapse(2) schedule(static) num_threads(portThreads())
        for (int vx = 0; vx < g.Ix; vx++)
            for (int vy = 0; vy < g.Iy; vy++) {
                const int target = in[vy + vx * g.Iy];
                const int x0 = g.rlox[vx], x1 = g.rhix[vx], y0 = g.rloy[vy], y1 = g.rhiy[vy];
                const int ncover = (x1 >= x0 && y1 >= y0) ? (x1 - x0 + 1) * (y1 - y0 + 1) : 0;
                int maxIndex = -1; float maxAct = -999999.0f;
                for (int vc = 0; vc < g.Iz; vc++) {
                    float sum = 0.0f;
                    for (int ox = x0; ox <= x1; ox++)
                        for (int oy = y0; oy <= y1; oy++)
                            sum += vl.w[g.at(ox, oy, hiddenCs[oy + ox * Hy], g.slotOf(ox, oy, vx, vy), vc)];
                    sum = sum / ncover; // float / (int countT / int Zh)
                    vl.recon[vc + (vy + vx * g.Iy) * g.Iz] = sum;
                    if (sum > maxAct || maxIndex == -1) { maxAct = sum; maxIndex = vc; }
                }
                if (maxIndex != target)
                    for (int vc = 0; vc < g.Iz; vc++) {
                        float delta = alpha * ((vc == target ? 1.0f : 0.0f) - std::exp(vl.recon[vc + (vy + vx * g.Iy) * g.Iz]));
                        for (int ox = x0; ox <= x1; ox++)
                            for (int oy = y0; oy <= y1; oy++) {
                                int h = oy + ox * Hy;
                                if (hiddenCs[h] != hiddenCsPrev[h])
                                    vl.w[g.at(ox, oy, hiddenCs[h], g.slotOf(ox, oy, vx, vy), vc)] += delta;
                            }
                    }
            }
    }
    // SparseCoder::step, SparseCoder.cpp:127-147
    void step(Rng &cs, const std::vector<const IntBuf *> &in, bool learnEnabled) {
        cs.consume2(Hx, Hy);
        forward(in);
        if (learnEnabled)
            for (size_t v = 0; v < vls.size(); v++) {
                cs.consume2(vlds[v].x, vlds[v].y);
                learn(*in[v], (int)v);
            }
        cs.consume1(Hx * Hy);
        hiddenCsPrev = hiddenCs;
    }
};

// Predictor (Predictor.h:15-150, Predictor.cpp)
struct Predictor {
    int Hx = 0, Hy = 0, Hz = 0;
    float alpha = 0.5f; // Predictor.h:85-89
    FloatBuf hiddenActivations; IntBuf hiddenCs;
    struct VL { Geom g; FloatBuf w; IntBuf inputCsPrev; };
    std::vector<VL> vls;
    std::vector<VLD> vlds;

    void allocate(int hx, int hy, int hz, const std::vector<VLD> &d) {
        Hx = hx; Hy = hy; Hz = hz; vlds = d; vls.resize(d.size());
        for (size_t i = 0; i < d.size(); i++) {
            vls[i].g.init(d[i].x, d[i].y, d[i].z, Hx, Hy, Hz, d[i].radius);
            vls[i].w.assign((size_t)vls[i].g.nnz, 0.0f);
            vls[i].inputCsPrev.assign((size_t)d[i].x * d[i].y, 0);
        }
        hiddenActivations.assign((size_t)Hx * Hy * Hz, 0.0f); hiddenCs.assign((size_t)Hx * Hy, 0);
    }
    // Predictor::initRandom, Predictor.cpp:72-109: weights U(-0.01, 0.01)
    void initRandom(Rng &cs, int hx, int hy, int hz, const std::vector<VLD> &d) {
        allocate(hx, hy, hz, d);
        std::uniform_real_distribution<float> weightDist(-0.01f, 0.01f);
        for (auto &vl : vls)
            for (auto &v : vl.w) v = weightDist(cs.rng);
    }
    // Predictor::forward, Predictor.cpp:13-47; Predictor::activate :111-127
    void activate(Rng &cs, const std::vector<const IntBuf *> &in) {
        cs.consume2(Hx, Hy);
#pragma omp parallel for collapse(2) schedule(static) num_threads(portThreads())
        for (int ox = 0; ox < Hx; ox++)
            for (int oy = 0; oy < Hy; oy++) {
                int maxIndex = 0; float maxAct = -999999.0f;
                for (int hc = 0; hc < Hz; hc++) {
                    float sum = 0.0f; int count = 0;
                    for (size_t v = 0; v < vls.size(); v++) {
                        sum += gatherRow(vls[v].g, vls[v].w, *in[v], ox, oy, hc);
                        count += vls[v].g.slots(ox, oy); // count(row) / Zv
                    }
                    sum /= count;
                    hiddenActivations[hc + (oy + ox * Hy) * Hz] = sum;
                    if (sum > maxAct) { maxAct = sum; maxIndex = hc; }
                }
                hiddenCs[oy + ox * Hy] = maxIndex;
            }
        for (size_t v = 0; v < vls.size(); v++) {
            cs.consume1(vlds[v].x * vlds[v].y);
            vls[v].inputCsPrev = *in[v];
        }
    }
    // Predictor::learn, Predictor.cpp:49-70 and :129-135
    void learn(Rng &cs, const IntBuf &target) {
        cs.consume2(Hx, Hy);
#pragma omp parallel for collapse(2) schedule(static) num_threads(portThreads())
        for (int ox = 0; ox < Hx; ox++)
            for (int oy = 0; oy < Hy; oy++) {
                int t = target[oy + ox * Hy];
                for (int hc = 0; hc < Hz; hc++) {
                    float delta = alpha * ((hc == t ? 1.0f : -1.0f) - std::tanh(hiddenActivations[hc + (oy + ox * Hy) * Hz]));
                    for (auto &vl : vls) deltaRow(vl.g, vl.w, vl.inputCsPrev, delta, ox, oy, hc);
                }
            }
    }
};

// Actor (Actor.h:15-181, Actor.cpp)
struct Actor {
    int Hx = 0, Hy = 0, Hz = 0;
    float alpha = 0.02f, beta = 0.02f, gamma = 0.99f; int minSteps = 8, historyIters = 8; // Actor.h:117-125
    int historySize = 0;
    FloatBuf hiddenActivations, hiddenValues; IntBuf hiddenCs;
    struct VL { Geom gv, ga; FloatBuf wv, wa; };
    std::vector<VL> vls;
    std::vector<VLD> vlds;
    struct Sample { std::vector<IntBuf> inputCs; IntBuf hiddenTargetCsPrev; FloatBuf hiddenValuesPrev; float reward = 0.0f; };
    std::vector<Sample> samples; int start = 0; // CircleBuffer, Helpers.h:85-139
    Sample &at(int i) { return samples[(size_t)(start + i) % samples.size()]; }

    void allocate(int hx, int hy, int hz, int capacity, const std::vector<VLD> &d) {
        Hx = hx; Hy = hy; Hz = hz; vlds = d; vls.resize(d.size());
        for (size_t i = 0; i < d.size(); i++) {
            vls[i].gv.init(d[i].x, d[i].y, d[i].z, Hx, Hy, 1, d[i].radius);
            vls[i].ga.init(d[i].x, d[i].y, d[i].z, Hx, Hy, Hz, d[i].radius);
            vls[i].wv.assign((size_t)vls[i].gv.nnz, 0.0f); vls[i].wa.assign((size_t)vls[i].ga.nnz, 0.0f);
        }
        hiddenActivations.assign((size_t)Hx * Hy * Hz, 0.0f); hiddenCs.assign((size_t)Hx * Hy, 0);
        hiddenValues.assign((size_t)Hx * Hy, 0.0f);
        historySize = 0; start = 0; samples.assign(capacity, Sample());
        for (auto &s : samples) {
            s.inputCs.resize(d.size());
            for (size_t i = 0; i < d.size(); i++) s.inputCs[i].assign((size_t)d[i].x * d[i].y, 0);
            s.hiddenTargetCsPrev.assign((size_t)Hx * Hy, 0); s.hiddenValuesPrev.assign((size_t)Hx * Hy, 0.0f);
        }
    }
    // Actor::initRandom, Actor.cpp:187-246: per visible layer value weights then action weights, U(-0.01,0.01)
    void initRandom(Rng &cs, int hx, int hy, int hz, int capacity, const std::vector<VLD> &d) {
        allocate(hx, hy, hz, capacity, d);
        std::uniform_real_distribution<float> weightDist(-0.01f, 0.01f);
        for (auto &vl : vls) {
            for (auto &v : vl.wv) v = weightDist(cs.rng);
            for (auto &v : vl.wa) v = weightDist(cs.rng);
        }
    }
    // Actor::forward, Actor.cpp:13-90, for one column with its tile's sub-generator
    void forwardColumn(int ox, int oy, std::mt19937 &rng, const std::vector<const IntBuf *> &in) {
        int col = oy + ox * Hy;
        float value = 0.0f; int count = 0;
        for (size_t v = 0; v < vls.size(); v++) {
            value += gatherRow(vls[v].gv, vls[v].wv, *in[v], ox, oy, 0);
            count += vls[v].gv.slots(ox, oy);
        }
        hiddenValues[col] = value / count;
        float maxAct = -999999.0f;
        for (int hc = 0; hc < Hz; hc++) {
            float sum = 0.0f;
            for (size_t v = 0; v < vls.size(); v++) sum += gatherRow(vls[v].ga, vls[v].wa, *in[v], ox, oy, hc);
            sum /= count;
            hiddenActivations[hc + col * Hz] = sum;
            maxAct = std::max(maxAct, sum);
        }
        float total = 0.0f;
        for (int hc = 0; hc < Hz; hc++) {
            hiddenActivations[hc + col * Hz] = std::exp(hiddenActivations[hc + col * Hz] - maxAct);
            total += hiddenActivations[hc + col * Hz];
        }
        std::uniform_real_distribution<float> cuspDist(0.0f, total);
        float cusp = cuspDist(rng);
        int select = 0; float sumSoFar = 0.0f;
        for (int hc = 0; hc < Hz; hc++) {
            sumSoFar += hiddenActivations[hc + col * Hz];
            if (sumSoFar >= cusp) { select = hc; break; }
        }
        hiddenCs[col] = select;
    }
    // Actor::learn, Actor.cpp:92-185, one column
    void learnColumn(int ox, int oy, const std::vector<IntBuf> &inPrev, const IntBuf &targetPrev, const FloatBuf &valuesPrev,
                     float q, float g, bool mimic) {
        int col = oy + ox * Hy;
        float newValue = q + g * hiddenValues[col];
        float value = 0.0f; int count = 0;
        for (size_t v = 0; v < vls.size(); v++) {
            value += gatherRow(vls[v].gv, vls[v].wv, inPrev[v], ox, oy, 0);
            count += vls[v].gv.slots(ox, oy);
        }
        value /= count;
        float tdErrorValue = newValue - value;
        float deltaValue = alpha * tdErrorValue;
        for (size_t v = 0; v < vls.size(); v++) deltaRow(vls[v].gv, vls[v].wv, inPrev[v], deltaValue, ox, oy, 0);
        float tdErrorAction = newValue - valuesPrev[col];
        int target = targetPrev[col];
        float maxAct = -999999.0f;
        for (int hc = 0; hc < Hz; hc++) {
            float sum = 0.0f;
            for (size_t v = 0; v < vls.size(); v++) sum += gatherRow(vls[v].ga, vls[v].wa, inPrev[v], ox, oy, hc);
            sum /= count;
            hiddenActivations[hc + col * Hz] = sum;
            maxAct = std::max(maxAct, sum);
        }
        float total = 0.0f;
        for (int hc = 0; hc < Hz; hc++) {
            hiddenActivations[hc + col * Hz] = std::exp(hiddenActivations[hc + col * Hz] - maxAct);
            total += hiddenActivations[hc + col * Hz];
        }
        for (int hc = 0; hc < Hz; hc++) {
            float deltaAction = (mimic ? beta : (tdErrorAction > 0.0f ? beta : -beta)) *
                                ((hc == target ? 1.0f : 0.0f) - hiddenActivations[hc + col * Hz] / std::max(0.0001f, total));
            for (size_t v = 0; v < vls.size(); v++) deltaRow(vls[v].ga, vls[v].wa, inPrev[v], deltaAction, ox, oy, hc);
        }
    }
    // Actor::step, Actor.cpp:248-313
    void step(Rng &cs, const std::vector<const IntBuf *> &in, const IntBuf &hiddenTargetCsPrev, float reward, bool learnEnabled,
              bool mimic) {
        std::vector<int> seeds;
        cs.consume2(Hx, Hy, &seeds);
        const int nbx = (Hx + cs.batch2x - 1) / cs.batch2x;
#pragma omp parallel for schedule(static) num_threads(portThreads())
        for (int i = 0; i < (int)seeds.size(); i++) { // tile i: bx = i % nbx, by = i / nbx; x outer, y inner (Helpers.cpp:55-70)
            std::mt19937 sub(seeds[i]);
            int px = (i % nbx) * cs.batch2x, py = (i / nbx) * cs.batch2y;
            for (int x = px; x < std::min(Hx, px + cs.batch2x); x++)
                for (int y = py; y < std::min(Hy, py + cs.batch2y); y++) forwardColumn(x, y, sub, in);
        }
        start--; if (start < 0) start += (int)samples.size(); // pushFront
        if (historySize < (int)samples.size()) historySize++;
        {
            Sample &s = at(0);
            for (size_t v = 0; v < vls.size(); v++) { cs.consume1(vlds[v].x * vlds[v].y); s.inputCs[v] = *in[v]; }
            cs.consume1(Hx * Hy); s.hiddenTargetCsPrev = hiddenTargetCsPrev;
            cs.consume1(Hx * Hy); s.hiddenValuesPrev = hiddenValues;
            s.reward = reward;
        }
        if (learnEnabled && historySize > minSteps + 1) {
            std::uniform_int_distribution<int> historyDist(minSteps, historySize - 2);
            for (int it = 0; it < historyIters; it++) {
                int idx = historyDist(cs.rng);
                const Sample &sPrev = at(idx + 1);
                const Sample &s = at(idx);
                float q = 0.0f, g = 1.0f;
                for (int t = idx; t >= 0; t--) { q += at(t).reward * g; g *= gamma; }
                cs.consume2(Hx, Hy);
#pragma omp parallel for collapse(2) schedule(static) num_threads(portThreads())
                for (int x = 0; x < Hx; x++)
                    for (int y = 0; y < Hy; y++) learnColumn(x, y, sPrev.inputCs, s.hiddenTargetCsPrev, sPrev.hiddenValuesPrev, q, g, mimic);
            }
        }
    }
};

// Hierarchy (Hierarchy.h:39-217, Hierarchy.cpp)
struct History { std::vector<IntBuf> data; int start = 0; // CircleBuffer<IntBuffer>
    IntBuf &at(int i) { return data[(size_t)(start + i) % data.size()]; }
    const IntBuf &at(int i) const { return data[(size_t)(start + i) % data.size()]; }
    void pushFront() { start--; if (start < 0) start += (int)data.size(); } };

struct Hierarchy {
    Rng cs;
    std::vector<SparseCoder> sc;
    std::vector<std::vector<std::unique_ptr<Predictor>>> p;
    std::vector<std::unique_ptr<Actor>> a;
    std::vector<std::vector<History>> histories;
    std::vector<char> updates; std::vector<int> ticks, ticksPerUpdate;
    std::vector<int> inX, inY, inZ;

    // Hierarchy::initRandom, Hierarchy.cpp:16-147
    void initRandom(const ogn_hierarchy_desc &d) {
        const int L = d.num_layers, NI = d.num_inputs;
        sc.resize(L); p.resize(L); histories.resize(L);
        ticks.assign(L, 0); updates.assign(L, 0); ticksPerUpdate.resize(L);
        inX.resize(NI); inY.resize(NI); inZ.resize(NI);
        for (int i = 0; i < NI; i++) { inX[i] = d.input_sizes[3 * i]; inY[i] = d.input_sizes[3 * i + 1]; inZ[i] = d.input_sizes[3 * i + 2]; }
        for (int l = 0; l < L; l++) ticksPerUpdate[l] = l == 0 ? 1 : d.layers[l].ticks_per_update;
        for (int l = 0; l < L; l++) {
            const ogn_layer_desc &ld = d.layers[l];
            std::vector<VLD> scV;
            std::vector<VLD> pV(1, VLD{ld.hidden_x, ld.hidden_y, ld.hidden_z, ld.p_radius});
            if (l < L - 1) pV.push_back(pV[0]);
            if (l == 0) {
                histories[l].resize(NI);
                scV.resize((size_t)NI * ld.temporal_horizon);
                for (int i = 0; i < NI; i++) {
                    for (int t = 0; t < ld.temporal_horizon; t++) scV[t + ld.temporal_horizon * i] = VLD{inX[i], inY[i], inZ[i], ld.ff_radius};
                    histories[l][i].data.assign(ld.temporal_horizon, IntBuf((size_t)inX[i] * inY[i], 0));
                }
                p[l].resize(NI); a.resize(NI);
                std::vector<VLD> aV(1, VLD{ld.hidden_x, ld.hidden_y, ld.hidden_z, ld.a_radius});
                if (l < L - 1) aV.push_back(aV[0]);
                for (int i = 0; i < NI; i++) {
                    if (d.input_types[i] == OGN_INPUT_PREDICTION) {
                        p[l][i].reset(new Predictor());
                        p[l][i]->initRandom(cs, inX[i], inY[i], inZ[i], pV);
                    } else if (d.input_types[i] == OGN_INPUT_ACTION) {
                        a[i].reset(new Actor());
                        a[i]->initRandom(cs, inX[i], inY[i], inZ[i], ld.history_capacity, aV);
                    }
                }
            } else {
                const ogn_layer_desc &lp = d.layers[l - 1];
                histories[l].resize(1);
                scV.assign(ld.temporal_horizon, VLD{lp.hidden_x, lp.hidden_y, lp.hidden_z, ld.ff_radius});
                histories[l][0].data.assign(ld.temporal_horizon, IntBuf((size_t)lp.hidden_x * lp.hidden_y, 0));
                p[l].resize(ld.ticks_per_update);
                for (auto &pp : p[l]) {
                    pp.reset(new Predictor());
                    pp->initRandom(cs, lp.hidden_x, lp.hidden_y, lp.hidden_z, pV);
                }
            }
            sc[l].initRandom(cs, ld.hidden_x, ld.hidden_y, ld.hidden_z, scV);
        }
    }

    // Hierarchy::step, Hierarchy.cpp:192-285
    void step(const std::vector<const IntBuf *> &inputCs, bool learnEnabled, float reward, bool mimic) {
        const int L = (int)sc.size();
        ticks[0] = 0;
        for (size_t i = 0; i < inX.size(); i++) {
            histories[0][i].pushFront();
            cs.consume1((int)inputCs[i]->size());
            histories[0][i].at(0) = *inputCs[i];
        }
        updates.assign(L, 0);
        for (int l = 0; l < L; l++) {
            if (l == 0 || ticks[l] >= ticksPerUpdate[l]) {
                ticks[l] = 0; updates[l] = 1;
                const size_t T = histories[l][0].data.size();
                std::vector<const IntBuf *> in(histories[l].size() * T);
                for (size_t i = 0; i < histories[l].size(); i++)
                    for (size_t t = 0; t < T; t++) in[t + T * i] = &histories[l][i].at((int)t);
                sc[l].step(cs, in, learnEnabled);
                if (l < L - 1) {
                    histories[l + 1][0].pushFront();
                    cs.consume1((int)sc[l].hiddenCs.size());
                    histories[l + 1][0].at(0) = sc[l].hiddenCs;
                    ticks[l + 1]++;
                }
            }
        }
        for (int l = L - 1; l >= 0; l--) {
            if (!updates[l]) continue;
            std::vector<const IntBuf *> fb(l < L - 1 ? 2 : 1);
            fb[0] = &sc[l].hiddenCs;
            if (l < L - 1) fb[1] = &p[l + 1][ticksPerUpdate[l + 1] - 1 - ticks[l + 1]]->hiddenCs;
            for (size_t k = 0; k < p[l].size(); k++)
                if (p[l][k]) {
                    if (learnEnabled) p[l][k]->learn(cs, l == 0 ? *inputCs[k] : histories[l][0].at((int)k));
                    p[l][k]->activate(cs, fb);
                }
            if (l == 0)
                for (size_t k = 0; k < a.size(); k++)
                    if (a[k]) a[k]->step(cs, fb, *inputCs[k], reward, learnEnabled, mimic);
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Save / load in the reference's stream layout (Hierarchy.cpp:287-432, SparseCoder.cpp:149-203,
// Predictor.cpp:137-192, Actor.cpp:315-438, Helpers.cpp:315-343, Helpers.h:311-341). The index arrays are regenerated
// from the geometry on write and skipped on read.
struct Writer {
    std::ofstream os;
    template <typename T> void pod(const T &v) { os.write((const char *)&v, sizeof(T)); }
    void int3(int x, int y, int z) { int v[4] = {x, y, z, 0}; os.write((const char *)v, 16); } // Int3 is 16 B (pad)
    template <typename T> void buf(const std::vector<T> &v) { // writeBufferToStream
        int n = (int)v.size(); pod(n);
        if (n > 0) os.write((const char *)v.data(), (size_t)n * sizeof(T));
    }
    void vld(const VLD &d) { int3(d.x, d.y, d.z); pod(d.radius); }
    void sm(const Geom &g, const FloatBuf &w, bool withT) { // writeSMToStream
        int rows = g.Ox * g.Oy * g.Oz, cols = g.Ix * g.Iy * g.Iz;
        pod(rows); pod(cols);
        buf(w);
        std::vector<int> rowRanges((size_t)rows + 1), columnIndices((size_t)g.nnz);
        int64_t j = 0;
        for (int ox = 0; ox < g.Ox; ox++)
            for (int oy = 0; oy < g.Oy; oy++)
                for (int oz = 0; oz < g.Oz; oz++) {
                    rowRanges[(size_t)oz + ((size_t)oy + (size_t)ox * g.Oy) * g.Oz] = (int)j;
                    for (int ix = g.lox[ox]; ix < g.lox[ox] + g.nx[ox]; ix++)
                        for (int iy = g.loy[oy]; iy < g.loy[oy] + g.ny[oy]; iy++)
                            for (int iz = 0; iz < g.Iz; iz++) columnIndices[(size_t)j++] = iz + iy * g.Iz + ix * g.Iz * g.Iy;
                }
        rowRanges[rows] = (int)j;
        std::vector<int> nzvi, colRanges, rowIdx;
        if (withT) { // SparseMatrix::initT, SparseMatrix.cpp:59-106
            colRanges.assign((size_t)cols + 1, 0); rowIdx.resize((size_t)g.nnz); nzvi.resize((size_t)g.nnz);
            for (int64_t k = 0; k < g.nnz; k++) colRanges[(size_t)columnIndices[(size_t)k]]++;
            int off = 0;
            for (int c = 0; c < cols; c++) { int t = colRanges[c]; colRanges[c] = off; off += t; }
            colRanges[cols] = off;
            std::vector<int> cur(colRanges.begin(), colRanges.end() - 1);
            for (int row = 0; row < rows; row++)
                for (int k = rowRanges[row]; k < rowRanges[row + 1]; k++) {
                    int t = cur[(size_t)columnIndices[(size_t)k]]++;
                    rowIdx[(size_t)t] = row; nzvi[(size_t)t] = k;
                }
        }
        buf(nzvi); buf(rowRanges); buf(columnIndices); buf(colRanges); buf(rowIdx);
    }
};
struct Reader {
    std::ifstream is;
    template <typename T> void pod(T &v) { is.read((char *)&v, sizeof(T)); }
    void int3(int &x, int &y, int &z) { int v[4]; is.read((char *)v, 16); x = v[0]; y = v[1]; z = v[2]; }
    template <typename T> void buf(std::vector<T> &v) { int n; pod(n); v.resize((size_t)n); if (n > 0) is.read((char *)v.data(), (size_t)n * sizeof(T)); }
    void skipBuf(size_t elem) { int n; pod(n); is.seekg((std::streamoff)n * (std::streamoff)elem, std::ios::cur); }
    void vld(VLD &d) { int3(d.x, d.y, d.z); pod(d.radius); }
    void sm(FloatBuf &w) { int rows, cols; pod(rows); pod(cols); buf(w); for (int i = 0; i < 5; i++) skipBuf(4); }
};

void writeSC(Writer &w, const SparseCoder &s) {
    w.int3(s.Hx, s.Hy, s.Hz); w.pod(s.alpha); w.buf(s.hiddenCs); w.buf(s.hiddenCsPrev);
    int n = (int)s.vls.size(); w.pod(n);
    for (int v = 0; v < n; v++) { w.vld(s.vlds[v]); w.sm(s.vls[v].g, s.vls[v].w, true); }
}
void readSC(Reader &r, SparseCoder &s) {
    int hx, hy, hz; r.int3(hx, hy, hz); float alpha; r.pod(alpha);
    IntBuf cs, csPrev; r.buf(cs); r.buf(csPrev);
    int n; r.pod(n);
    std::vector<VLD> d(n); std::vector<FloatBuf> ws(n);
    for (int v = 0; v < n; v++) { r.vld(d[v]); r.sm(ws[v]); }
    s.allocate(hx, hy, hz, d); s.alpha = alpha; s.hiddenCs = cs; s.hiddenCsPrev = csPrev;
    for (int v = 0; v < n; v++) s.vls[v].w = ws[v];
}
void writeP(Writer &w, const Predictor &s) {
    w.int3(s.Hx, s.Hy, s.Hz); w.pod(s.alpha); w.buf(s.hiddenActivations); w.buf(s.hiddenCs);
    int n = (int)s.vls.size(); w.pod(n);
    for (int v = 0; v < n; v++) { w.vld(s.vlds[v]); w.sm(s.vls[v].g, s.vls[v].w, false); w.buf(s.vls[v].inputCsPrev); }
}
void readP(Reader &r, Predictor &s) {
    int hx, hy, hz; r.int3(hx, hy, hz); float alpha; r.pod(alpha);
    FloatBuf act; IntBuf cs; r.buf(act); r.buf(cs);
    int n; r.pod(n);
    std::vector<VLD> d(n); std::vector<FloatBuf> ws(n); std::vector<IntBuf> prev(n);
    for (int v = 0; v < n; v++) { r.vld(d[v]); r.sm(ws[v]); r.buf(prev[v]); }
    s.allocate(hx, hy, hz, d); s.alpha = alpha; s.hiddenActivations = act; s.hiddenCs = cs;
    for (int v = 0; v < n; v++) { s.vls[v].w = ws[v]; s.vls[v].inputCsPrev = prev[v]; }
}
void writeA(Writer &w, Actor &s) {
    w.int3(s.Hx, s.Hy, s.Hz); w.pod(s.alpha); w.pod(s.beta); w.pod(s.gamma); w.pod(s.minSteps); w.pod(s.historyIters);
    w.buf(s.hiddenCs); w.buf(s.hiddenValues);
    int n = (int)s.vls.size(); w.pod(n);
    for (int v = 0; v < n; v++) { w.vld(s.vlds[v]); w.sm(s.vls[v].gv, s.vls[v].wv, false); w.sm(s.vls[v].ga, s.vls[v].wa, false); }
    w.pod(s.historySize); int ns = (int)s.samples.size(); w.pod(ns); w.pod(s.start);
    for (int t = 0; t < ns; t++) {
        Actor::Sample &sm = s.at(t);
        for (int v = 0; v < n; v++) w.buf(sm.inputCs[v]);
        w.buf(sm.hiddenTargetCsPrev); w.buf(sm.hiddenValuesPrev); w.pod(sm.reward);
    }
}
void readA(Reader &r, Actor &s) {
    int hx, hy, hz; r.int3(hx, hy, hz);
    float alpha, beta, gamma; int minSteps, historyIters;
    r.pod(alpha); r.pod(beta); r.pod(gamma); r.pod(minSteps); r.pod(historyIters);
    IntBuf cs; FloatBuf vals; r.buf(cs); r.buf(vals);
    int n; r.pod(n);
    std::vector<VLD> d(n); std::vector<FloatBuf> wv(n), wa(n);
    for (int v = 0; v < n; v++) { r.vld(d[v]); r.sm(wv[v]); r.sm(wa[v]); }
    int historySize, ns, start; r.pod(historySize); r.pod(ns); r.pod(start);
    s.allocate(hx, hy, hz, ns, d);
    s.alpha = alpha; s.beta = beta; s.gamma = gamma; s.minSteps = minSteps; s.historyIters = historyIters;
    s.hiddenCs = cs; s.hiddenValues = vals; s.historySize = historySize; s.start = start;
    for (int v = 0; v < n; v++) { s.vls[v].wv = wv[v]; s.vls[v].wa = wa[v]; }
    for (int t = 0; t < ns; t++) {
        Actor::Sample &sm = s.at(t);
        for (int v = 0; v < n; v++) r.buf(sm.inputCs[v]);
        r.buf(sm.hiddenTargetCsPrev); r.buf(sm.hiddenValuesPrev); r.pod(sm.reward);
    }
}
int64_t saveHierarchy(Hierarchy &h, const char *path) {
    Writer w; w.os.open(path, std::ios::binary);
    if (!w.os) return -1;
    int L = (int)h.sc.size(), NI = (int)h.inX.size();
    w.pod(L); w.pod(NI);
    for (int i = 0; i < NI; i++) w.int3(h.inX[i], h.inY[i], h.inZ[i]);
    w.os.write(h.updates.data(), L);
    w.os.write((const char *)h.ticks.data(), L * 4);
    w.os.write((const char *)h.ticksPerUpdate.data(), L * 4);
    for (int l = 0; l < L; l++) {
        int nh = (int)h.histories[l].size(); w.pod(nh);
        for (auto &hist : h.histories[l]) {
            int sz = (int)hist.data.size(); w.pod(sz); w.pod(hist.start);
            for (int t = 0; t < sz; t++) w.buf(hist.at(t));
        }
        writeSC(w, h.sc[l]);
        for (auto &pp : h.p[l]) { char e = pp != nullptr; w.pod(e); if (e) writeP(w, *pp); }
    }
    for (auto &aa : h.a) { char e = aa != nullptr; w.pod(e); if (e) writeA(w, *aa); }
    return (int64_t)w.os.tellp();
}
bool loadHierarchy(Hierarchy &h, const char *path) {
    Reader r; r.is.open(path, std::ios::binary);
    if (!r.is) return false;
    int L, NI; r.pod(L); r.pod(NI);
    h.inX.resize(NI); h.inY.resize(NI); h.inZ.resize(NI);
    for (int i = 0; i < NI; i++) r.int3(h.inX[i], h.inY[i], h.inZ[i]);
    h.sc.resize(L); h.p.resize(L); h.histories.resize(L);
    h.updates.resize(L); h.ticks.resize(L); h.ticksPerUpdate.resize(L);
    r.is.read(h.updates.data(), L);
    r.is.read((char *)h.ticks.data(), L * 4);
    r.is.read((char *)h.ticksPerUpdate.data(), L * 4);
    for (int l = 0; l < L; l++) {
        int nh; r.pod(nh); h.histories[l].resize(nh);
        for (auto &hist : h.histories[l]) {
            int sz; r.pod(sz); r.pod(hist.start); hist.data.resize(sz);
            for (int t = 0; t < sz; t++) r.buf(hist.at(t));
        }
        readSC(r, h.sc[l]);
        h.p[l].resize(l == 0 ? NI : h.ticksPerUpdate[l]);
        for (auto &pp : h.p[l]) { char e; r.pod(e); if (e) { pp.reset(new Predictor()); readP(r, *pp); } else pp.reset(); }
    }
    h.a.resize(NI);
    for (auto &aa : h.a) { char e; r.pod(e); if (e) { aa.reset(new Actor()); readA(r, *aa); } else aa.reset(); }
    return (bool)r.is;
}

thread_local std::string g_err;
template <typename T> int copyOut(const std::vector<T> &v, T *out) {
    if (out) std::memcpy(out, v.data(), v.size() * sizeof(T));
    return (int)v.size();
}
int64_t copyW(const FloatBuf &v, float *out, int64_t cap) {
    if (out) std::memcpy(out, v.data(), (size_t)std::min<int64_t>(cap, (int64_t)v.size()) * sizeof(float));
    return (int64_t)v.size();
}
} // namespace

extern "C" {
OGN_DECLARE_HIERARCHY_API(oport_)

int oport_create(const ogn_hierarchy_desc *d, uint32_t seed, void **out) {
    auto *h = new Hierarchy();
    h->cs.rng.seed(seed);
    h->initRandom(*d);
    *out = h;
    return 0;
}
void oport_destroy(void *h) { delete (Hierarchy *)h; }
const char *oport_last_error(void) { return g_err.c_str(); }
int oport_set_num_threads(int n) { g_portThreads = n; return 0; }
int oport_step(void *hh, const int *const *input_cs, int learn, float reward, int mimic) {
    Hierarchy &h = *(Hierarchy *)hh;
    std::vector<IntBuf> bufs(h.inX.size());
    std::vector<const IntBuf *> ptrs(h.inX.size());
    for (size_t i = 0; i < bufs.size(); i++) { bufs[i].assign(input_cs[i], input_cs[i] + h.inX[i] * h.inY[i]); ptrs[i] = &bufs[i]; }
    h.step(ptrs, learn != 0, reward, mimic != 0);
    return 0;
}
int oport_num_layers(void *h) { return (int)((Hierarchy *)h)->sc.size(); }
int oport_get_update(void *h, int l) { return ((Hierarchy *)h)->updates[l]; }
int oport_get_ticks(void *h, int l) { return ((Hierarchy *)h)->ticks[l]; }
int oport_get_ticks_per_update(void *h, int l) { return ((Hierarchy *)h)->ticksPerUpdate[l]; }
int oport_get_prediction_cs(void *hh, int i, int *out) {
    Hierarchy &h = *(Hierarchy *)hh;
    if (h.a[i]) return copyOut(h.a[i]->hiddenCs, out);
    if (h.p[0][i]) return copyOut(h.p[0][i]->hiddenCs, out);
    return -1;
}
int oport_sc_hidden_cs(void *h, int l, int *out) { return copyOut(((Hierarchy *)h)->sc[l].hiddenCs, out); }
int oport_sc_hidden_cs_prev(void *h, int l, int *out) { return copyOut(((Hierarchy *)h)->sc[l].hiddenCsPrev, out); }
int oport_sc_num_visible(void *h, int l) { return (int)((Hierarchy *)h)->sc[l].vls.size(); }
int64_t oport_sc_weights(void *h, int l, int vli, float *out, int64_t cap) { return copyW(((Hierarchy *)h)->sc[l].vls[vli].w, out, cap); }
int oport_sc_reconstructions(void *h, int l, int vli, float *out) { return copyOut(((Hierarchy *)h)->sc[l].vls[vli].recon, out); }
int oport_pred_count(void *h, int l) { return (int)((Hierarchy *)h)->p[l].size(); }
int oport_pred_exists(void *h, int l, int p) { return ((Hierarchy *)h)->p[l][p] != nullptr; }
int oport_pred_hidden_cs(void *h, int l, int p, int *out) { return copyOut(((Hierarchy *)h)->p[l][p]->hiddenCs, out); }
int oport_pred_activations(void *h, int l, int p, float *out) { return copyOut(((Hierarchy *)h)->p[l][p]->hiddenActivations, out); }
int oport_pred_num_visible(void *h, int l, int p) { return (int)((Hierarchy *)h)->p[l][p]->vls.size(); }
int64_t oport_pred_weights(void *h, int l, int p, int vli, float *out, int64_t cap) { return copyW(((Hierarchy *)h)->p[l][p]->vls[vli].w, out, cap); }
int oport_pred_input_cs_prev(void *h, int l, int p, int vli, int *out) { return copyOut(((Hierarchy *)h)->p[l][p]->vls[vli].inputCsPrev, out); }
int oport_actor_exists(void *h, int p) { return ((Hierarchy *)h)->a[p] != nullptr; }
int oport_actor_hidden_cs(void *h, int p, int *out) { return copyOut(((Hierarchy *)h)->a[p]->hiddenCs, out); }
int oport_actor_values(void *h, int p, float *out) { return copyOut(((Hierarchy *)h)->a[p]->hiddenValues, out); }
int oport_actor_num_visible(void *h, int p) { return (int)((Hierarchy *)h)->a[p]->vls.size(); }
int64_t oport_actor_weights(void *h, int p, int vli, int which, float *out, int64_t cap) {
    Actor::VL &vl = ((Hierarchy *)h)->a[p]->vls[vli];
    return copyW(which == 0 ? vl.wv : vl.wa, out, cap);
}
int oport_set_sc_alpha(void *h, int l, float a) { ((Hierarchy *)h)->sc[l].alpha = a; return 0; }
int oport_set_pred_alpha(void *h, int l, int p, float a) { ((Hierarchy *)h)->p[l][p]->alpha = a; return 0; }
int oport_set_actor_params(void *h, int p, float alpha, float beta, float gamma, int minSteps, int historyIters) {
    Actor &a = *((Hierarchy *)h)->a[p];
    a.alpha = alpha; a.beta = beta; a.gamma = gamma; a.minSteps = minSteps; a.historyIters = historyIters;
    return 0;
}
float oport_get_sc_alpha(void *h, int l) { return ((Hierarchy *)h)->sc[l].alpha; }
float oport_get_pred_alpha(void *h, int l, int p) { return ((Hierarchy *)h)->p[l][p]->alpha; }
int oport_get_actor_params(void *h, int p, float *alpha, float *beta, float *gamma, int *minSteps, int *historyIters) {
    const Actor &a = *((Hierarchy *)h)->a[p];
    *alpha = a.alpha; *beta = a.beta; *gamma = a.gamma; *minSteps = a.minSteps; *historyIters = a.historyIters;
    return 0;
}
} // extern "C"
template <typename L> static int portShape(const L &layer, int vli, int *out3, int *out4) {
    if (out3) { out3[0] = layer.Hx; out3[1] = layer.Hy; out3[2] = layer.Hz; }
    if (out4) {
        if (vli < 0 || vli >= (int)layer.vlds.size()) { g_err = "no such visible layer"; return 1; }
        out4[0] = layer.vlds[vli].x; out4[1] = layer.vlds[vli].y; out4[2] = layer.vlds[vli].z; out4[3] = layer.vlds[vli].radius;
    }
    return 0;
}
static int portLayerShape(void *hh, int kind, int l, int vli, int *out3, int *out4) {
    Hierarchy &h = *(Hierarchy *)hh;
    if (kind == 0) return portShape(h.sc[l], vli, out3, out4);
    if (kind > 0) {
        if (kind - 1 >= (int)h.p[l].size() || !h.p[l][kind - 1]) { g_err = "no such predictor"; return 1; }
        return portShape(*h.p[l][kind - 1], vli, out3, out4);
    }
    if (-1 - kind >= (int)h.a.size() || !h.a[-1 - kind]) { g_err = "no such actor"; return 1; }
    return portShape(*h.a[-1 - kind], vli, out3, out4);
}
extern "C" {
int oport_layer_hidden_size(void *h, int kind, int l, int *out3) { return portLayerShape(h, kind, l, 0, out3, nullptr); }
int oport_layer_visible_desc(void *h, int kind, int l, int vli, int *out4) { return portLayerShape(h, kind, l, vli, nullptr, out4); }
int64_t oport_save(void *h, const char *path) { return saveHierarchy(*(Hierarchy *)h, path); }
int oport_load(const char *path, uint32_t seed, void **out) {
    auto *h = new Hierarchy();
    h->cs.rng.seed(seed);
    if (!loadHierarchy(*h, path)) { delete h; g_err = "cannot read file"; return 1; }
    *out = h;
    return 0;
}
// Hierarchy::getState / setState, Hierarchy.cpp:434-493: every CSDR, the history rings, ticks and updates (no weights)
struct PortState {
    std::vector<IntBuf> hiddenCs, hiddenCsPrev;
    std::vector<std::vector<IntBuf>> predHiddenCs;
    std::vector<std::vector<std::vector<IntBuf>>> predInputCsPrev;
    std::vector<std::vector<History>> histories;
    std::vector<char> updates;
    std::vector<int> ticks;
};
int oport_get_state(void *hh, void **state_out) {
    const Hierarchy &h = *(Hierarchy *)hh;
    auto *s = new PortState();
    const size_t L = h.sc.size();
    s->hiddenCs.resize(L); s->hiddenCsPrev.resize(L); s->predHiddenCs.resize(L); s->predInputCsPrev.resize(L);
    for (size_t l = 0; l < L; l++) {
        s->hiddenCs[l] = h.sc[l].hiddenCs; s->hiddenCsPrev[l] = h.sc[l].hiddenCsPrev;
        s->predHiddenCs[l].resize(h.p[l].size()); s->predInputCsPrev[l].resize(h.p[l].size());
        for (size_t j = 0; j < h.p[l].size(); j++) {
            if (!h.p[l][j]) continue;
            s->predHiddenCs[l][j] = h.p[l][j]->hiddenCs;
            for (auto &vl : h.p[l][j]->vls) s->predInputCsPrev[l][j].push_back(vl.inputCsPrev);
        }
    }
    s->histories = h.histories; s->updates = h.updates; s->ticks = h.ticks;
    *state_out = s;
    return 0;
}
int oport_set_state(void *hh, const void *state) {
    if (!state) { g_err = "null state"; return 1; }
    Hierarchy &h = *(Hierarchy *)hh;
    const PortState &s = *(const PortState *)state;
    for (size_t l = 0; l < h.sc.size(); l++) {
        h.sc[l].hiddenCs = s.hiddenCs[l]; h.sc[l].hiddenCsPrev = s.hiddenCsPrev[l];
        for (size_t j = 0; j < h.p[l].size(); j++) {
            if (!h.p[l][j]) continue;
            h.p[l][j]->hiddenCs = s.predHiddenCs[l][j];
            for (size_t v = 0; v < h.p[l][j]->vls.size(); v++) h.p[l][j]->vls[v].inputCsPrev = s.predInputCsPrev[l][j][v];
        }
    }
    h.histories = s.histories; h.updates = s.updates; h.ticks = s.ticks;
    return 0;
}
void oport_free_state(void *state) { delete (PortState *)state; }
uint32_t oport_rng_peek(void *h) { std::mt19937 c = ((Hierarchy *)h)->cs.rng; return (uint32_t)c(); }
} // extern "C"

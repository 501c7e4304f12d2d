// oracle/ienc_oracle.cpp — TEST INFRASTRUCTURE ONLY (see oracle/sph_oracle.cpp for the rules): CPU restatement of the
// reference's ImageEncoder (source/ogmaneo/ImageEncoder.cpp:19-225, SparseMatrix::distance2 :122-137, hebb :692-701,
// multiplyOHVsT :278-294, countT :215-221), exported through include/ogn_image_encoder_api.h with the prefix oport_.
// Written from the algorithm with closed-form receptive-field geometry instead of CSR index arrays.
// PARITY PINNING: the reference has no tests for this class either; tests/test_image_encoder.py steps this port and the
// compiled reference (oracle/_ref, prefix oref_) side by side and demands bit-equal hidden CSDRs, resources, weights,
// reconstructions and save files; tests/golden/image_encoder.npz holds a reference-generated trajectory.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <random>
#include <string>
#include <vector>

#include "../include/ogn_image_encoder_api.h"

namespace {
struct Geo { // initSMLocalRF, Helpers.cpp:236-313 (output = hidden grid, input = visible grid)
    int Ox, Oy, Oz, Ix, Iy, Iz, r;
    std::vector<int> lox, nx, loy, ny;
    std::vector<int64_t> colBase;
    int64_t nnz = 0;
    void init(int ix, int iy, int iz, int ox, int oy, int oz, int radius) {
        Ix = ix; Iy = iy; Iz = iz; Ox = ox; Oy = oy; Oz = oz; r = radius;
        const float sx = (float)Ix / (float)Ox, sy = (float)Iy / (float)Oy; // project(), Helpers.h:223-228
        lox.resize(Ox); nx.resize(Ox); loy.resize(Oy); ny.resize(Oy);
        for (int x = 0; x < Ox; x++) { int c = (int)(x * sx + 0.5f), lo = std::max(0, c - r), hi = std::min(Ix - 1, c + r); lox[x] = lo; nx[x] = std::max(0, hi - lo + 1); }
        for (int y = 0; y < Oy; y++) { int c = (int)(y * sy + 0.5f), lo = std::max(0, c - r), hi = std::min(Iy - 1, c + r); loy[y] = lo; ny[y] = std::max(0, hi - lo + 1); }
        colBase.resize((size_t)Ox * Oy);
        int64_t off = 0;
        for (int x = 0; x < Ox; x++)
            for (int y = 0; y < Oy; y++) { colBase[(size_t)y + (size_t)x * Oy] = off; off += (int64_t)nx[x] * ny[y] * Iz * Oz; }
        nnz = off;
    }
    // offset of row (ox, oy, oz): its entries follow in the order ix, iy, iz ascending
    int64_t row(int ox, int oy, int oz) const { return colBase[(size_t)oy + (size_t)ox * Oy] + (int64_t)oz * nx[ox] * ny[oy] * Iz; }
};

struct Encoder {
    int Hx, Hy, Hz;
    float alpha = 0.01f, gamma = 0.01f; // ImageEncoder.h:98-104
    std::vector<int> hiddenCs;
    std::vector<float> hiddenResources;
    struct VL { Geo g; std::vector<float> w, recon; ogn_ienc_visible_desc d; };
    std::vector<VL> vls;
    std::mt19937 rng;

    void allocate(int hx, int hy, int hz, int n, const ogn_ienc_visible_desc *d) {
        Hx = hx; Hy = hy; Hz = hz;
        vls.resize(n);
        for (int v = 0; v < n; v++) {
            vls[v].d = d[v];
            vls[v].g.init(d[v].size_x, d[v].size_y, d[v].size_z, Hx, Hy, Hz, d[v].radius);
            vls[v].w.assign((size_t)vls[v].g.nnz, 0.0f);
            vls[v].recon.assign((size_t)d[v].size_x * d[v].size_y * d[v].size_z, 0.0f);
        }
        hiddenCs.assign((size_t)Hx * Hy, 0);
        hiddenResources.assign((size_t)Hx * Hy * Hz, 1.0f);
    }
    // ImageEncoder::initRandom, ImageEncoder.cpp:96-137: weights U(0,1) drawn sequentially in value order
    void initRandom() {
        std::uniform_real_distribution<float> weightDist(0.0f, 1.0f);
        for (auto &vl : vls)
            for (auto &x : vl.w) x = weightDist(rng);
    }
    // ImageEncoder::forward, ImageEncoder.cpp:19-74 for every hidden column
    void step(const float *const *in, bool learn) {
        for (int ox = 0; ox < Hx; ox++)
            for (int oy = 0; oy < Hy; oy++) {
                int maxIndex = 0;
                float maxActivation = -999999.0f;
                std::vector<std::pair<float, int>> act((size_t)Hz);
                for (int hc = 0; hc < Hz; hc++) {
                    float sum = 0.0f;
                    for (size_t v = 0; v < vls.size(); v++) { // sum -= distance2(inputActs[v], hiddenIndex)  (SparseMatrix.cpp:122-137)
                        const Geo &g = vls[v].g;
                        const float *w = vls[v].w.data() + g.row(ox, oy, hc);
                        float d2 = 0.0f;
                        int64_t j = 0;
                        for (int ix = g.lox[ox]; ix < g.lox[ox] + g.nx[ox]; ix++)
                            for (int iy = g.loy[oy]; iy < g.loy[oy] + g.ny[oy]; iy++)
                                for (int iz = 0; iz < g.Iz; iz++, j++) {
                                    const float delta = in[v][iz + g.Iz * (iy + g.Iy * ix)] - w[j];
                                    d2 += delta * delta;
                                }
                        sum -= d2;
                    }
                    act[(size_t)hc] = std::make_pair(sum, hc);
                    if (sum > maxActivation) { maxActivation = sum; maxIndex = hc; }
                }
                hiddenCs[(size_t)oy + (size_t)ox * Hy] = maxIndex;
                if (!learn) continue;
                // the reference calls std::sort with a comparator on the value only ("largest in front"). std::sort is not
                // stable, so for equal activations the order is whatever libstdc++'s introsort leaves; this restatement makes the
                // same call (same library, same permutation). The CUDA path replays that algorithm (csrc/ogn_stdsort.h).
                std::sort(act.begin(), act.end(), [](const std::pair<float, int> &a, const std::pair<float, int> &b) { return a.first > b.first; });
                for (int i = 0; i < Hz; i++) {
                    const int hc = act[(size_t)i].second;
                    float &res = hiddenResources[(size_t)hc + (size_t)Hz * ((size_t)oy + (size_t)ox * Hy)];
                    const float strength = std::exp(-i * i * gamma / std::max(0.001f, res)) * res;
                    res -= alpha * strength;
                    for (size_t v = 0; v < vls.size(); v++) { // hebb, SparseMatrix.cpp:692-701
                        const Geo &g = vls[v].g;
                        float *w = vls[v].w.data() + g.row(ox, oy, hc);
                        int64_t j = 0;
                        for (int ix = g.lox[ox]; ix < g.lox[ox] + g.nx[ox]; ix++)
                            for (int iy = g.loy[oy]; iy < g.loy[oy] + g.ny[oy]; iy++)
                                for (int iz = 0; iz < g.Iz; iz++, j++) w[j] += strength * (in[v][iz + g.Iz * (iy + g.Iy * ix)] - w[j]);
                    }
                }
            }
    }
    // ImageEncoder::backward, ImageEncoder.cpp:76-94: recon = multiplyOHVsT(hiddenCs) / max(1, countT / Hz)
    void reconstruct(const int *hcs) {
        for (auto &vl : vls) {
            const Geo &g = vl.g;
            for (int ix = 0; ix < g.Ix; ix++)
                for (int iy = 0; iy < g.Iy; iy++)
                    for (int iz = 0; iz < g.Iz; iz++) {
                        float sum = 0.0f;
                        int count = 0;
                        for (int ox = 0; ox < Hx; ox++) { // transpose order: ascending row = ascending (ox, oy)
                            if (ix < g.lox[ox] || ix >= g.lox[ox] + g.nx[ox]) continue;
                            for (int oy = 0; oy < Hy; oy++) {
                                if (iy < g.loy[oy] || iy >= g.loy[oy] + g.ny[oy]) continue;
                                const int hc = hcs[(size_t)oy + (size_t)ox * Hy];
                                const int s = (ix - g.lox[ox]) * g.ny[oy] + (iy - g.loy[oy]);
                                sum += vl.w[(size_t)(g.row(ox, oy, hc) + (int64_t)s * g.Iz + iz)];
                                count++;
                            }
                        }
                        vl.recon[(size_t)iz + (size_t)g.Iz * ((size_t)iy + (size_t)g.Iy * ix)] = sum / std::max(1, count);
                    }
        }
    }
};

// writeSMToStream with regenerated index arrays (Helpers.cpp:315-328, SparseMatrix::initT SparseMatrix.cpp:59-106)
template <typename T> void pod(std::ofstream &os, const T &v) { os.write((const char *)&v, sizeof(T)); }
template <typename T> void buf(std::ofstream &os, const std::vector<T> &v) { int n = (int)v.size(); pod(os, n); if (n > 0) os.write((const char *)v.data(), (size_t)n * sizeof(T)); }
void writeSM(std::ofstream &os, const Geo &g, const std::vector<float> &w) {
    const int rows = g.Ox * g.Oy * g.Oz, cols = g.Ix * g.Iy * g.Iz;
    pod(os, rows); pod(os, cols);
    buf(os, w);
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
    rowRanges[(size_t)rows] = (int)j;
    std::vector<int> colRanges((size_t)cols + 1, 0), rowIdx((size_t)g.nnz), nzvi((size_t)g.nnz);
    for (int64_t k = 0; k < g.nnz; k++) colRanges[(size_t)columnIndices[(size_t)k]]++;
    int off = 0;
    for (int c = 0; c < cols; c++) { int t = colRanges[(size_t)c]; colRanges[(size_t)c] = off; off += t; }
    colRanges[(size_t)cols] = off;
    std::vector<int> cur(colRanges.begin(), colRanges.end() - 1);
    for (int row = 0; row < rows; row++)
        for (int k = rowRanges[(size_t)row]; k < rowRanges[(size_t)row + 1]; k++) {
            int t = cur[(size_t)columnIndices[(size_t)k]]++;
            rowIdx[(size_t)t] = row; nzvi[(size_t)t] = k;
        }
    buf(os, nzvi); buf(os, rowRanges); buf(os, columnIndices); buf(os, colRanges); buf(os, rowIdx);
}
thread_local std::string g_err;
} // namespace

extern "C" {
OGN_DECLARE_IMAGE_ENCODER_API(oport_)

int oport_ienc_create(int hx, int hy, int hz, int n, const ogn_ienc_visible_desc *d, uint32_t seed, void **out) {
    auto *e = new Encoder();
    e->rng.seed(seed);
    e->allocate(hx, hy, hz, n, d);
    e->initRandom();
    *out = e;
    return 0;
}
void oport_ienc_destroy(void *e) { delete (Encoder *)e; }
const char *oport_ienc_last_error(void) { return g_err.c_str(); }
int oport_ienc_step(void *e, const float *const *in, int learn) { ((Encoder *)e)->step(in, learn != 0); return 0; }
int oport_ienc_reconstruct(void *e, const int *hcs) { ((Encoder *)e)->reconstruct(hcs); return 0; }
int oport_ienc_hidden_cs(void *e, int *out) {
    auto &v = ((Encoder *)e)->hiddenCs;
    if (out) std::memcpy(out, v.data(), v.size() * sizeof(int));
    return (int)v.size();
}
int oport_ienc_hidden_resources(void *e, float *out) {
    auto &v = ((Encoder *)e)->hiddenResources;
    if (out) std::memcpy(out, v.data(), v.size() * sizeof(float));
    return (int)v.size();
}
int oport_ienc_reconstruction(void *e, int vli, float *out) {
    auto &v = ((Encoder *)e)->vls[(size_t)vli].recon;
    if (out) std::memcpy(out, v.data(), v.size() * sizeof(float));
    return (int)v.size();
}
int64_t oport_ienc_weights(void *e, int vli, float *out, int64_t cap) {
    auto &v = ((Encoder *)e)->vls[(size_t)vli].w;
    if (out) std::memcpy(out, v.data(), (size_t)std::min<int64_t>(cap, (int64_t)v.size()) * sizeof(float));
    return (int64_t)v.size();
}
int oport_ienc_set_params(void *e, float alpha, float gamma) { ((Encoder *)e)->alpha = alpha; ((Encoder *)e)->gamma = gamma; return 0; }
int64_t oport_ienc_save(void *ee, const char *path) { // ImageEncoder::writeToStream, ImageEncoder.cpp:162-195
    Encoder &e = *(Encoder *)ee;
    std::ofstream os(path, std::ios::binary);
    if (!os) return -1;
    const int hs[4] = {e.Hx, e.Hy, e.Hz, 0}; // Int3 is 16 bytes (pad)
    os.write((const char *)hs, 16);
    pod(os, e.alpha); pod(os, e.gamma);
    buf(os, e.hiddenCs); buf(os, e.hiddenResources);
    const int n = (int)e.vls.size();
    pod(os, n);
    for (auto &vl : e.vls) {
        const int d[5] = {vl.d.size_x, vl.d.size_y, vl.d.size_z, 0, vl.d.radius}; // VisibleLayerDesc = Int3 (16 B) + radius
        os.write((const char *)d, 20);
        writeSM(os, vl.g, vl.w);
        buf(os, vl.recon);
    }
    return (int64_t)os.tellp();
}
int oport_ienc_load(const char *path, void **out) { // ImageEncoder::readFromStream, ImageEncoder.cpp:197-225
    std::ifstream is(path, std::ios::binary);
    if (!is) { g_err = "cannot open file"; return 1; }
    auto rd = [&](void *p, size_t n) { is.read((char *)p, (std::streamsize)n); };
    auto rbuf = [&](auto &v) { int n = 0; rd(&n, 4); v.resize((size_t)n); if (n > 0) rd(v.data(), (size_t)n * sizeof(v[0])); };
    auto skip = [&](size_t elem) { int n = 0; rd(&n, 4); is.seekg((std::streamoff)n * (std::streamoff)elem, std::ios::cur); };
    int hs[4];
    rd(hs, 16);
    auto *e = new Encoder();
    rd(&e->alpha, 4); rd(&e->gamma, 4);
    std::vector<int> hcs; std::vector<float> res;
    rbuf(hcs); rbuf(res);
    int n = 0;
    rd(&n, 4);
    std::vector<ogn_ienc_visible_desc> d((size_t)n);
    std::vector<std::vector<float>> ws((size_t)n), recs((size_t)n);
    for (int v = 0; v < n; v++) {
        int raw[5];
        rd(raw, 20);
        d[(size_t)v] = ogn_ienc_visible_desc{raw[0], raw[1], raw[2], raw[4]};
        int rows, cols;
        rd(&rows, 4); rd(&cols, 4);
        rbuf(ws[(size_t)v]);
        for (int k = 0; k < 5; k++) skip(4);
        rbuf(recs[(size_t)v]);
    }
    e->allocate(hs[0], hs[1], hs[2], n, d.data());
    e->hiddenCs = hcs; e->hiddenResources = res;
    for (int v = 0; v < n; v++) { e->vls[(size_t)v].w = ws[(size_t)v]; e->vls[(size_t)v].recon = recs[(size_t)v]; }
    *out = e;
    return is ? 0 : 1;
}
} // extern "C"

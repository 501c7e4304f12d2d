// oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY. Not part of the product.
//
// Exposes the UNMODIFIED reference implementation (compiled from /root/reference/source/ogmaneo/*.cpp where it
// lies, namespace renamed with -Dogmaneo=ogmaneo_ref, see oracle/Makefile) through the flat hierarchy C ABI of
// include/ogn_hierarchy_api.h with the prefix oref_. It is used to (1) pin oracle/sph_oracle.cpp, (2) generate
// the fixtures under tests/golden/, (3) act as the parity checker in tests/ and (4) be the CPU baseline
// (`cpu_baseline.kind == "reference"`) in bench.py. Nothing under ogmaneo2_b200/ may link or load this.
//
// No reference source is copied: this file only includes the reference's public headers at build time.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <istream>
#include <memory>
#include <ostream>
#include <random>
#include <string>
#include <vector>
#include <omp.h>

// The reference keeps Predictor::hiddenActivations and Actor::hiddenValues private with no getter; the parity
// tests compare them, so open the classes up for this one translation unit (layout is unaffected).
#define private public
#include <ogmaneo/Hierarchy.h>
#include <ogmaneo/ImageEncoder.h>
#undef private

#include "../include/ogn_hierarchy_api.h"
#include "../include/ogn_image_encoder_api.h"

using namespace ogmaneo; // == ogmaneo_ref after macro renaming

namespace {
struct Handle {
    ComputeSystem cs;
    Hierarchy h;
};
thread_local std::string g_err;
int fail(const char *m) { g_err = m; return 1; }
template <typename T> int copyOut(const std::vector<T> &v, T *out) {
    if (out) std::memcpy(out, v.data(), v.size() * sizeof(T));
    return (int)v.size();
}
int64_t copyW(const std::vector<float> &v, float *out, int64_t cap) {
    if (out) std::memcpy(out, v.data(), (size_t)std::min<int64_t>(cap, (int64_t)v.size()) * sizeof(float));
    return (int64_t)v.size();
}
} // namespace

extern "C" {
OGN_DECLARE_HIERARCHY_API(oref_)

int oref_create(const ogn_hierarchy_desc *d, uint32_t seed, void **out) {
    try {
        auto *H = new Handle();
        H->cs.rng.seed(seed);
        std::vector<Int3> sizes(d->num_inputs);
        std::vector<InputType> types(d->num_inputs);
        for (int i = 0; i < d->num_inputs; i++) {
            sizes[i] = Int3(d->input_sizes[3 * i], d->input_sizes[3 * i + 1], d->input_sizes[3 * i + 2]);
            sizes[i].pad = 0;
            types[i] = (InputType)d->input_types[i];
        }
        std::vector<Hierarchy::LayerDesc> lds(d->num_layers);
        for (int l = 0; l < d->num_layers; l++) {
            const ogn_layer_desc &s = d->layers[l];
            lds[l].hiddenSize = Int3(s.hidden_x, s.hidden_y, s.hidden_z);
            lds[l].hiddenSize.pad = 0;
            lds[l].ffRadius = s.ff_radius;
            lds[l].pRadius = s.p_radius;
            lds[l].ticksPerUpdate = s.ticks_per_update;
            lds[l].temporalHorizon = s.temporal_horizon;
            lds[l].aRadius = s.a_radius;
            lds[l].historyCapacity = s.history_capacity;
        }
        H->h.initRandom(H->cs, sizes, types, lds);
        *out = H;
        return 0;
    } catch (const std::exception &e) {
        return fail(e.what());
    }
}
void oref_destroy(void *h) { delete (Handle *)h; }
const char *oref_last_error(void) { return g_err.c_str(); }
int oref_set_num_threads(int n) { ComputeSystem::setNumThreads(n); return 0; }

int oref_step(void *hh, const int *const *input_cs, int learn, float reward, int mimic) {
    Handle *H = (Handle *)hh;
    const std::vector<Int3> &sz = H->h.getInputSizes();
    std::vector<IntBuffer> bufs(sz.size());
    std::vector<const IntBuffer *> ptrs(sz.size());
    for (size_t i = 0; i < sz.size(); i++) {
        bufs[i].assign(input_cs[i], input_cs[i] + sz[i].x * sz[i].y);
        ptrs[i] = &bufs[i];
    }
    H->h.step(H->cs, ptrs, learn != 0, reward, mimic != 0);
    return 0;
}
int oref_num_layers(void *h) { return ((Handle *)h)->h.getNumLayers(); }
int oref_get_update(void *h, int l) { return ((Handle *)h)->h.getUpdate(l); }
int oref_get_ticks(void *h, int l) { return ((Handle *)h)->h.getTicks(l); }
int oref_get_ticks_per_update(void *h, int l) { return ((Handle *)h)->h.getTicksPerUpdate(l); }
int oref_get_prediction_cs(void *hh, int i, int *out) {
    Hierarchy &h = ((Handle *)hh)->h;
    if (h.getALayers()[i] == nullptr && h.getPLayers(0)[i] == nullptr) return -1;
    return copyOut(h.getPredictionCs(i), out);
}
int oref_sc_hidden_cs(void *h, int l, int *out) { return copyOut(((Handle *)h)->h.getSCLayer(l).getHiddenCs(), out); }
int oref_sc_hidden_cs_prev(void *h, int l, int *out) {
    return copyOut(((Handle *)h)->h.getSCLayer(l).getHiddenCsPrev(), out);
}
int oref_sc_num_visible(void *h, int l) { return ((Handle *)h)->h.getSCLayer(l).getNumVisibleLayers(); }
int64_t oref_sc_weights(void *h, int l, int vli, float *out, int64_t cap) {
    return copyW(((Handle *)h)->h.getSCLayer(l).getVisibleLayer(vli).weights.nonZeroValues, out, cap);
}
int oref_sc_reconstructions(void *h, int l, int vli, float *out) {
    return copyOut(((Handle *)h)->h.getSCLayer(l).getVisibleLayer(vli).reconstructions, out);
}
int oref_pred_count(void *h, int l) { return (int)((Handle *)h)->h.getPLayers(l).size(); }
int oref_pred_exists(void *h, int l, int p) { return ((Handle *)h)->h.getPLayers(l)[p] != nullptr; }
int oref_pred_hidden_cs(void *h, int l, int p, int *out) {
    return copyOut(((Handle *)h)->h.getPLayers(l)[p]->getHiddenCs(), out);
}
int oref_pred_activations(void *h, int l, int p, float *out) {
    return copyOut(((Handle *)h)->h.getPLayers(l)[p]->hiddenActivations, out);
}
int oref_pred_num_visible(void *h, int l, int p) { return ((Handle *)h)->h.getPLayers(l)[p]->getNumVisibleLayers(); }
int64_t oref_pred_weights(void *h, int l, int p, int vli, float *out, int64_t cap) {
    return copyW(((Handle *)h)->h.getPLayers(l)[p]->getVisibleLayer(vli).weights.nonZeroValues, out, cap);
}
int oref_pred_input_cs_prev(void *h, int l, int p, int vli, int *out) {
    return copyOut(((Handle *)h)->h.getPLayers(l)[p]->getVisibleLayer(vli).inputCsPrev, out);
}
int oref_actor_exists(void *h, int p) { return ((Handle *)h)->h.getALayers()[p] != nullptr; }
int oref_actor_hidden_cs(void *h, int p, int *out) { return copyOut(((Handle *)h)->h.getALayers()[p]->getHiddenCs(), out); }
int oref_actor_values(void *h, int p, float *out) { return copyOut(((Handle *)h)->h.getALayers()[p]->hiddenValues, out); }
int oref_actor_num_visible(void *h, int p) { return ((Handle *)h)->h.getALayers()[p]->getNumVisibleLayers(); }
int64_t oref_actor_weights(void *h, int p, int vli, int which, float *out, int64_t cap) {
    const Actor::VisibleLayer &vl = ((Handle *)h)->h.getALayers()[p]->getVisibleLayer(vli);
    return copyW(which == 0 ? vl.valueWeights.nonZeroValues : vl.actionWeights.nonZeroValues, out, cap);
}
int oref_set_sc_alpha(void *h, int l, float a) { ((Handle *)h)->h.getSCLayer(l).alpha = a; return 0; }
int oref_set_pred_alpha(void *h, int l, int p, float a) { ((Handle *)h)->h.getPLayers(l)[p]->alpha = a; return 0; }
int oref_set_actor_params(void *h, int p, float alpha, float beta, float gamma, int minSteps, int historyIters) {
    Actor &a = *((Handle *)h)->h.getALayers()[p];
    a.alpha = alpha; a.beta = beta; a.gamma = gamma; a.minSteps = minSteps; a.historyIters = historyIters;
    return 0;
}
float oref_get_sc_alpha(void *h, int l) { return ((Handle *)h)->h.getSCLayer(l).alpha; }
float oref_get_pred_alpha(void *h, int l, int p) { return ((Handle *)h)->h.getPLayers(l)[p]->alpha; }
int oref_get_actor_params(void *h, int p, float *alpha, float *beta, float *gamma, int *minSteps, int *historyIters) {
    const Actor &a = *((Handle *)h)->h.getALayers()[p];
    *alpha = a.alpha; *beta = a.beta; *gamma = a.gamma; *minSteps = a.minSteps; *historyIters = a.historyIters;
    return 0;
}
} // extern "C"
namespace {
template <typename L> void shapeOut(const L &layer, int vli, int *out3, int *out4) {
    if (out3) { const Int3 &s = layer.getHiddenSize(); out3[0] = s.x; out3[1] = s.y; out3[2] = s.z; }
    if (out4) {
        const auto &d = layer.getVisibleLayerDesc(vli);
        out4[0] = d.size.x; out4[1] = d.size.y; out4[2] = d.size.z; out4[3] = d.radius;
    }
}
int layerShape(void *hh, int kind, int l, int vli, int *out3, int *out4) {
    Hierarchy &h = ((Handle *)hh)->h;
    if (kind == 0) shapeOut(h.getSCLayer(l), vli, out3, out4);
    else if (kind > 0) {
        if (kind - 1 >= (int)h.getPLayers(l).size() || !h.getPLayers(l)[kind - 1]) return fail("no such predictor");
        shapeOut(*h.getPLayers(l)[kind - 1], vli, out3, out4);
    } else {
        if (-1 - kind >= (int)h.getALayers().size() || !h.getALayers()[-1 - kind]) return fail("no such actor");
        shapeOut(*h.getALayers()[-1 - kind], vli, out3, out4);
    }
    return 0;
}
} // namespace
extern "C" {
int oref_layer_hidden_size(void *h, int kind, int l, int *out3) { return layerShape(h, kind, l, 0, out3, nullptr); }
int oref_layer_visible_desc(void *h, int kind, int l, int vli, int *out4) { return layerShape(h, kind, l, vli, nullptr, out4); }
int64_t oref_save(void *h, const char *path) {
    std::ofstream os(path, std::ios::binary);
    if (!os) return -1;
    ((Handle *)h)->h.writeToStream(os);
    return (int64_t)os.tellp();
}
int oref_load(const char *path, uint32_t seed, void **out) {
    std::ifstream is(path, std::ios::binary);
    if (!is) return fail("cannot open file");
    auto *H = new Handle();
    H->cs.rng.seed(seed);
    H->h.readFromStream(is);
    *out = H;
    return 0;
}
int oref_get_state(void *h, void **state_out) {
    auto *s = new State();
    ((Handle *)h)->h.getState(*s);
    *state_out = s;
    return 0;
}
int oref_set_state(void *h, const void *state) {
    if (!state) return fail("null state");
    ((Handle *)h)->h.setState(*(const State *)state);
    return 0;
}
void oref_free_state(void *state) { delete (State *)state; }
uint32_t oref_rng_peek(void *h) {
    std::mt19937 c = ((Handle *)h)->cs.rng;
    return (uint32_t)c();
}
} // extern "C"

// ---- ImageEncoder (ImageEncoder.h / ImageEncoder.cpp) through include/ogn_image_encoder_api.h, prefix oref_ --------------
namespace {
struct IEnc {
    ComputeSystem cs;
    ImageEncoder e;
};
} // namespace
extern "C" {
OGN_DECLARE_IMAGE_ENCODER_API(oref_)
int oref_ienc_create(int hx, int hy, int hz, int n, const ogn_ienc_visible_desc *d, uint32_t seed, void **out) {
    try {
        auto *E = new IEnc();
        E->cs.rng.seed(seed);
        std::vector<ImageEncoder::VisibleLayerDesc> vlds((size_t)n);
        for (int v = 0; v < n; v++) {
            vlds[(size_t)v].size = Int3(d[v].size_x, d[v].size_y, d[v].size_z);
            vlds[(size_t)v].size.pad = 0;
            vlds[(size_t)v].radius = d[v].radius;
        }
        Int3 hs(hx, hy, hz);
        hs.pad = 0;
        E->e.initRandom(E->cs, hs, vlds);
        *out = E;
        return 0;
    } catch (const std::exception &ex) {
        return fail(ex.what());
    }
}
void oref_ienc_destroy(void *e) { delete (IEnc *)e; }
const char *oref_ienc_last_error(void) { return g_err.c_str(); }
int oref_ienc_step(void *ee, const float *const *in, int learn) {
    IEnc *E = (IEnc *)ee;
    const int n = E->e.getNumVisibleLayers();
    std::vector<FloatBuffer> bufs((size_t)n);
    std::vector<const FloatBuffer *> ptrs((size_t)n);
    for (int v = 0; v < n; v++) {
        const Int3 &s = E->e.getVisibleLayerDesc(v).size;
        bufs[(size_t)v].assign(in[v], in[v] + (size_t)s.x * s.y * s.z);
        ptrs[(size_t)v] = &bufs[(size_t)v];
    }
    E->e.step(E->cs, ptrs, learn != 0);
    return 0;
}
int oref_ienc_reconstruct(void *ee, const int *hcs) {
    IEnc *E = (IEnc *)ee;
    const Int3 &s = E->e.getHiddenSize();
    IntBuffer b(hcs, hcs + (size_t)s.x * s.y);
    E->e.reconstruct(E->cs, &b);
    return 0;
}
int oref_ienc_hidden_cs(void *e, int *out) { return copyOut(((IEnc *)e)->e.getHiddenCs(), out); }
int oref_ienc_hidden_resources(void *e, float *out) { return copyOut(((IEnc *)e)->e.hiddenResources, out); }
int oref_ienc_reconstruction(void *e, int vli, float *out) { return copyOut(((IEnc *)e)->e.getVisibleLayer(vli).reconstructions, out); }
int64_t oref_ienc_weights(void *e, int vli, float *out, int64_t cap) {
    return copyW(((IEnc *)e)->e.getVisibleLayer(vli).weights.nonZeroValues, out, cap);
}
int oref_ienc_set_params(void *e, float alpha, float gamma) { ((IEnc *)e)->e.alpha = alpha; ((IEnc *)e)->e.gamma = gamma; return 0; }
int64_t oref_ienc_save(void *e, const char *path) {
    std::ofstream os(path, std::ios::binary);
    if (!os) return -1;
    // the reference leaves Int3::pad uninitialised: zero the pads of the descriptors so that files compare
    IEnc *E = (IEnc *)e;
    E->e.hiddenSize.pad = 0;
    for (auto &d : E->e.visibleLayerDescs) d.size.pad = 0;
    E->e.writeToStream(os);
    return (int64_t)os.tellp();
}
int oref_ienc_load(const char *path, void **out) {
    std::ifstream is(path, std::ios::binary);
    if (!is) return fail("cannot open file");
    auto *E = new IEnc();
    E->e.readFromStream(is);
    *out = E;
    return 0;
}
} // extern "C"


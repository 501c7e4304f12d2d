// csrc/host/FlatApi.cpp — the flat C ABI (include/ogn_hierarchy_api.h, prefix ogn_) on top of the C++ drop-in classes,
// plus product-only entry points (include/ogn_b200.h): device-resident stepping, ensembles, sharding, hash init.
#include <algorithm>
#include <ogmaneo/Hierarchy.h>

#include <cstring>
#include <fstream>
#include <string>

#include <ogmaneo/ImageEncoder.h>

#include "../../../include/ogn_b200.h"
#include "../../../include/ogn_image_encoder_api.h"
#include "Device.h"

using namespace ogmaneo;

namespace {
struct Handle {
    ComputeSystem cs;
    Hierarchy h;
    std::vector<InputType> types;
    std::vector<Hierarchy::LayerDesc> lds; // empty when the handle came from ogn_load (the reference format has no descs)
};
thread_local std::string g_err;

template <typename F> int guard(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    } catch (...) {
        g_err = "unknown error";
        return 1;
    }
}
template <typename T> int copyOut(const std::vector<T> &v, T *out) {
    if (out && !v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(T));
    return (int)v.size();
}

void unpack(const ogn_hierarchy_desc *d, std::vector<Int3> &sizes, std::vector<InputType> &types, std::vector<Hierarchy::LayerDesc> &lds) {
    sizes.resize(d->num_inputs); types.resize(d->num_inputs); lds.resize(d->num_layers);
    for (int i = 0; i < d->num_inputs; i++) {
        sizes[i] = Int3(d->input_sizes[3 * i], d->input_sizes[3 * i + 1], d->input_sizes[3 * i + 2]);
        types[i] = (InputType)d->input_types[i];
    }
    for (int l = 0; l < d->num_layers; l++) {
        const ogn_layer_desc &s = d->layers[l];
        lds[l].hiddenSize = Int3(s.hidden_x, s.hidden_y, s.hidden_z);
        lds[l].ffRadius = s.ff_radius; lds[l].pRadius = s.p_radius; lds[l].ticksPerUpdate = s.ticks_per_update;
        lds[l].temporalHorizon = s.temporal_horizon; lds[l].aRadius = s.a_radius; lds[l].historyCapacity = s.history_capacity;
    }
}
int64_t weightsOut(const std::vector<float> &w, float *out, int64_t cap) {
    if (out && cap > 0) std::memcpy(out, w.data(), (size_t)std::min<int64_t>(cap, (int64_t)w.size()) * sizeof(float));
    return (int64_t)w.size();
}
} // namespace

extern "C" {

const char *ogn_last_error(void) { return g_err.c_str(); }
int ogn_set_num_threads(int n) { ComputeSystem::setNumThreads(n); return 0; }

int ogn_create_ex(const ogn_hierarchy_desc *d, uint32_t seed, const ogn_create_options *opt, void **out) {
    *out = nullptr;
    return guard([&] {
        std::unique_ptr<Handle> H(new Handle());
        H->cs.rng.seed(seed);
        std::vector<Int3> sizes; std::vector<Hierarchy::LayerDesc> lds;
        unpack(d, sizes, H->types, lds);
        H->lds = lds;
        int ensemble = 1;
        if (opt) {
            H->cs.device = opt->device;
            H->cs.deviceInit = opt->device_init != 0;
            H->cs.deviceInitSeed = seed;
            if (opt->replay_rng) H->cs.replayRng = true;
            H->cs.shardRank = opt->shard_rank; H->cs.shardWorld = opt->shard_world > 0 ? opt->shard_world : 1;
            H->cs.allGather = opt->all_gather; H->cs.allGatherUser = opt->all_gather_user;
            H->cs.peerExchange = opt->peer_exchange != 0;
            ensemble = opt->ensemble > 0 ? opt->ensemble : 1;
        }
        if (ensemble > 1 || (opt && opt->ensemble_seeds)) {
            std::vector<unsigned> seeds(ensemble);
            for (int k = 0; k < ensemble; k++) seeds[k] = opt->ensemble_seeds ? opt->ensemble_seeds[k] : seed + (unsigned)k;
            H->h.initRandomEnsemble(H->cs, seeds, sizes, H->types, lds);
        } else
            H->h.initRandom(H->cs, sizes, H->types, lds);
        *out = H.release();
    });
}
int ogn_create(const ogn_hierarchy_desc *d, uint32_t seed, void **out) {
    ogn_create_options opt;
    std::memset(&opt, 0, sizeof(opt));
    opt.device = -1; opt.replay_rng = 1; opt.shard_world = 1; opt.ensemble = 1;
    return ogn_create_ex(d, seed, &opt, out);
}
void ogn_destroy(void *h) {
    try { delete (Handle *)h; } catch (...) {}
}

int ogn_step(void *hh, const int *const *input_cs, int learn, float reward, int mimic) {
    return guard([&] {
        Handle *H = (Handle *)hh;
        H->h.stepHost(H->cs, input_cs, learn != 0, reward, mimic != 0);
    });
}
int ogn_step_device(void *hh, const int *const *input_cs_dev, int learn, float reward, int mimic) {
    return guard([&] {
        Handle *H = (Handle *)hh;
        std::vector<const int *> ptrs(input_cs_dev, input_cs_dev + H->h.getInputSizes().size());
        H->h.stepDevice(H->cs, ptrs, learn != 0, reward, mimic != 0);
    });
}
int ogn_synchronize(void *hh) { return guard([&] { ((Handle *)hh)->cs.synchronize(); }); }
void *ogn_stream(void *hh) {
    void *s = nullptr;
    guard([&] { s = ((Handle *)hh)->cs.stream(); });
    return s;
}
const int *ogn_prediction_cs_device(void *hh, int i) {
    const int *p = nullptr;
    guard([&] { p = ((Handle *)hh)->h.getPredictionCsDevice(i); });
    return p;
}
int ogn_prediction_readback_queue(void *hh, int i, int x0, int x1) {
    int t = -1;
    guard([&] { t = ((Handle *)hh)->h.queuePredictionReadback(i, x0, x1); });
    return t;
}
int ogn_prediction_readback_finish(void *hh, int ticket, int *out) {
    int n = -1;
    guard([&] { n = ((Handle *)hh)->h.finishPredictionReadback(ticket, out); });
    return n;
}
int ogn_ensemble_size(void *hh) { return ((Handle *)hh)->h.getEnsembleSize(); }
uint64_t ogn_take_sc_update_count(void *hh, int l) {
    uint64_t v = 0;
    guard([&] { v = ((Handle *)hh)->h.getSCLayer(l).takeUpdateCount(); });
    return v;
}

int ogn_num_layers(void *h) { return ((Handle *)h)->h.getNumLayers(); }
int ogn_get_update(void *h, int l) { return ((Handle *)h)->h.getUpdate(l); }
int ogn_get_ticks(void *h, int l) { return ((Handle *)h)->h.getTicks(l); }
int ogn_get_ticks_per_update(void *h, int l) { return ((Handle *)h)->h.getTicksPerUpdate(l); }

int ogn_get_prediction_cs(void *hh, int i, int *out) {
    int n = -1;
    guard([&] {
        Hierarchy &h = ((Handle *)hh)->h;
        if (h.getALayers()[i] == nullptr && h.getPLayers(0)[i] == nullptr) return;
        n = h.copyPredictionCs(i, out);
    });
    return n;
}
int ogn_sc_hidden_cs(void *h, int l, int *out) {
    int n = -1;
    guard([&] { n = copyOut(((Handle *)h)->h.getSCLayer(l).getHiddenCs(), out); });
    return n;
}
int ogn_sc_hidden_cs_prev(void *h, int l, int *out) {
    int n = -1;
    guard([&] { n = copyOut(((Handle *)h)->h.getSCLayer(l).getHiddenCsPrev(), out); });
    return n;
}
int ogn_sc_num_visible(void *h, int l) { return ((Handle *)h)->h.getSCLayer(l).getNumVisibleLayers(); }
int64_t ogn_sc_weights(void *h, int l, int vli, float *out, int64_t cap) {
    int64_t n = -1;
    guard([&] {
        std::vector<float> w;
        ((Handle *)h)->h.getSCLayer(l).getWeights(vli, w);
        n = weightsOut(w, out, cap);
    });
    return n;
}
int ogn_sc_reconstructions(void *h, int l, int vli, float *out) {
    int n = -1;
    guard([&] {
        FloatBuffer r;
        ((Handle *)h)->h.getSCLayer(l).getReconstructions(vli, r);
        n = copyOut(r, out);
    });
    return n;
}
int ogn_pred_count(void *h, int l) { return (int)((Handle *)h)->h.getPLayers(l).size(); }
int ogn_pred_exists(void *h, int l, int p) { return ((Handle *)h)->h.getPLayers(l)[p] != nullptr; }
int ogn_pred_hidden_cs(void *h, int l, int p, int *out) {
    int n = -1;
    guard([&] { n = copyOut(((Handle *)h)->h.getPLayers(l)[p]->getHiddenCs(), out); });
    return n;
}
int ogn_pred_activations(void *h, int l, int p, float *out) {
    int n = -1;
    guard([&] { n = copyOut(((Handle *)h)->h.getPLayers(l)[p]->getHiddenActivations(), out); });
    return n;
}
int ogn_pred_num_visible(void *h, int l, int p) { return ((Handle *)h)->h.getPLayers(l)[p]->getNumVisibleLayers(); }
int64_t ogn_pred_weights(void *h, int l, int p, int vli, float *out, int64_t cap) {
    int64_t n = -1;
    guard([&] {
        std::vector<float> w;
        ((Handle *)h)->h.getPLayers(l)[p]->getWeights(vli, w);
        n = weightsOut(w, out, cap);
    });
    return n;
}
int ogn_pred_input_cs_prev(void *h, int l, int p, int vli, int *out) {
    int n = -1;
    guard([&] { n = copyOut(((Handle *)h)->h.getPLayers(l)[p]->getInputCsPrev(vli), out); });
    return n;
}
int ogn_actor_exists(void *h, int p) { return ((Handle *)h)->h.getALayers()[p] != nullptr; }
int ogn_actor_hidden_cs(void *h, int p, int *out) {
    int n = -1;
    guard([&] { n = copyOut(((Handle *)h)->h.getALayers()[p]->getHiddenCs(), out); });
    return n;
}
int ogn_actor_values(void *h, int p, float *out) {
    int n = -1;
    guard([&] { n = copyOut(((Handle *)h)->h.getALayers()[p]->getHiddenValues(), out); });
    return n;
}
int ogn_actor_num_visible(void *h, int p) { return ((Handle *)h)->h.getALayers()[p]->getNumVisibleLayers(); }
int64_t ogn_actor_weights(void *h, int p, int vli, int which, float *out, int64_t cap) {
    int64_t n = -1;
    guard([&] {
        std::vector<float> w;
        ((Handle *)h)->h.getALayers()[p]->getWeights(vli, which, w);
        n = weightsOut(w, out, cap);
    });
    return n;
}
int ogn_set_sc_alpha(void *h, int l, float a) { ((Handle *)h)->h.getSCLayer(l).alpha = a; return 0; }
int ogn_set_pred_alpha(void *h, int l, int p, float a) { ((Handle *)h)->h.getPLayers(l)[p]->alpha = a; return 0; }
int ogn_set_actor_params(void *h, int p, float alpha, float beta, float gamma, int minSteps, int historyIters) {
    Actor &a = *((Handle *)h)->h.getALayers()[p];
    a.alpha = alpha; a.beta = beta; a.gamma = gamma; a.minSteps = minSteps; a.historyIters = historyIters;
    return 0;
}
float ogn_get_sc_alpha(void *h, int l) { return ((Handle *)h)->h.getSCLayer(l).alpha; }
float ogn_get_pred_alpha(void *h, int l, int p) { return ((Handle *)h)->h.getPLayers(l)[p]->alpha; }
int ogn_get_actor_params(void *h, int p, float *alpha, float *beta, float *gamma, int *minSteps, int *historyIters) {
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
    return guard([&] {
        Hierarchy &h = ((Handle *)hh)->h;
        if (kind == 0) shapeOut(h.getSCLayer(l), vli, out3, out4);
        else if (kind > 0) {
            if (kind - 1 >= (int)h.getPLayers(l).size() || !h.getPLayers(l)[kind - 1]) throw std::out_of_range("no such predictor");
            shapeOut(*h.getPLayers(l)[kind - 1], vli, out3, out4);
        } else {
            if (-1 - kind >= (int)h.getALayers().size() || !h.getALayers()[-1 - kind]) throw std::out_of_range("no such actor");
            shapeOut(*h.getALayers()[-1 - kind], vli, out3, out4);
        }
    });
}
} // namespace
extern "C" {
int ogn_layer_hidden_size(void *h, int kind, int l, int *out3) { return layerShape(h, kind, l, 0, out3, nullptr); }
int ogn_layer_visible_desc(void *h, int kind, int l, int vli, int *out4) { return layerShape(h, kind, l, vli, nullptr, out4); }
int64_t ogn_save(void *h, const char *path) {
    int64_t n = -1;
    guard([&] {
        std::ofstream os(path, std::ios::binary);
        if (!os) throw std::runtime_error("cannot open file for writing");
        ((Handle *)h)->h.writeToStream(os);
        n = (int64_t)os.tellp();
    });
    return n;
}
int ogn_load(const char *path, uint32_t seed, void **out) {
    *out = nullptr;
    return guard([&] {
        std::ifstream is(path, std::ios::binary);
        if (!is) throw std::runtime_error("cannot open file");
        std::unique_ptr<Handle> H(new Handle());
        H->cs.rng.seed(seed);
        H->cs.replayRng = true;
        H->h.readFromStream(is, H->cs);
        *out = H.release();
    });
}
int64_t ogn_checkpoint_save(void *h, const char *path) {
    int64_t n = -1;
    guard([&] {
        Handle *H = (Handle *)h;
        if (H->lds.empty()) throw std::runtime_error("checkpointing needs a handle made by ogn_create / ogn_create_ex / ogn_checkpoint_load");
        std::ofstream os(path, std::ios::binary);
        if (!os) throw std::runtime_error("cannot open file for writing");
        H->h.writeCheckpoint(os, H->types, H->lds);
        // trailer: the ComputeSystem's generator (Actor sampling and history draws continue from it), libstdc++ text form
        os.write("OGNRNG1\n", 8);
        os << H->cs.rng << "\n";
        n = (int64_t)os.tellp();
    });
    return n;
}
int ogn_checkpoint_load(const char *path, const ogn_create_options *opt, void **out) {
    *out = nullptr;
    return guard([&] {
        std::ifstream is(path, std::ios::binary);
        if (!is) throw std::runtime_error("cannot open file");
        std::unique_ptr<Handle> H(new Handle());
        if (opt) {
            H->cs.device = opt->device;
            if (opt->replay_rng) H->cs.replayRng = true;
            H->cs.shardRank = opt->shard_rank; H->cs.shardWorld = opt->shard_world > 0 ? opt->shard_world : 1;
            H->cs.allGather = opt->all_gather; H->cs.allGatherUser = opt->all_gather_user;
            H->cs.peerExchange = opt->peer_exchange != 0;
        }
        H->h.readCheckpoint(is, H->cs, &H->types);
        {   // optional trailer with the generator state (files written before it existed simply end here)
            char tag[8] = {0};
            is.read(tag, 8);
            if (is && std::memcmp(tag, "OGNRNG1\n", 8) == 0) is >> H->cs.rng;
            is.clear();
        }
        // the layer descriptors travel in the file; keep them so that the handle can be checkpointed again
        const int L = H->h.getNumLayers();
        H->lds.resize(L);
        is.clear();
        is.seekg(8 + 5 * 4 + (std::streamoff)H->types.size() * 16);
        for (int l = 0; l < L; l++) {
            int32_t v[9];
            is.read(reinterpret_cast<char *>(v), sizeof(v));
            H->lds[l].hiddenSize = Int3(v[0], v[1], v[2]); H->lds[l].ffRadius = v[3]; H->lds[l].pRadius = v[4];
            H->lds[l].ticksPerUpdate = v[5]; H->lds[l].temporalHorizon = v[6]; H->lds[l].aRadius = v[7]; H->lds[l].historyCapacity = v[8];
        }
        *out = H.release();
    });
}
int ogn_get_state(void *h, void **state_out) {
    *state_out = nullptr;
    return guard([&] {
        std::unique_ptr<State> s(new State());
        ((Handle *)h)->h.getState(*s);
        *state_out = s.release();
    });
}
int ogn_set_state(void *h, const void *state) {
    return guard([&] {
        if (!state) throw std::invalid_argument("ogn_set_state: null state");
        ((Handle *)h)->h.setState(*(const State *)state);
    });
}
void ogn_free_state(void *state) { delete (State *)state; }
uint32_t ogn_rng_peek(void *h) {
    std::mt19937 c = ((Handle *)h)->cs.rng;
    return (uint32_t)c();
}
int ogn_shard_plan(int Hx, int Vx, int r, int rank, int world, int out[6]) {
    return guard([&] {
        int a = 0, b = 0;
        detail::slabOf(Hx, rank, world, a, b);
        // project(): (int)(pos * (float)in / (float)out + 0.5f), field [c - r, c + r] clamped (Helpers.cpp:245-262)
        const float scale = (float)Vx / (float)Hx;
        auto lo = [&](int o) { return std::max(0, (int)(o * scale + 0.5f) - r); };
        auto hi = [&](int o) { return std::min(Vx - 1, (int)(o * scale + 0.5f) + r); };
        int vx0 = Vx, vx1 = 0;
        for (int o = a; o < b; o++)
            if (hi(o) >= lo(o)) { vx0 = std::min(vx0, lo(o)); vx1 = std::max(vx1, hi(o) + 1); }
        if (vx1 < vx0) vx0 = vx1 = 0;
        int rx0 = Hx, rx1 = 0;
        for (int o = 0; o < Hx; o++)
            if (hi(o) >= lo(o) && lo(o) < vx1 && hi(o) >= vx0) { rx0 = std::min(rx0, o); rx1 = std::max(rx1, o + 1); }
        if (rx1 < rx0) rx0 = rx1 = 0;
        if (world <= 1) { vx0 = 0; vx1 = Vx; rx0 = 0; rx1 = Hx; }
        out[0] = a; out[1] = b; out[2] = vx0; out[3] = vx1; out[4] = rx0; out[5] = rx1;
    });
}
int ogn_peer_export(void *hh, void *handle_out) {
    return guard([&] { ((Handle *)hh)->cs.peerExport(handle_out); });
}
int ogn_peer_attach(void *hh, const void *handles, int world) {
    return guard([&] { ((Handle *)hh)->cs.peerAttach(handles, world); });
}
uint64_t ogn_peer_take_wait_ns(void *hh) {
    uint64_t v = 0;
    guard([&] { v = ((Handle *)hh)->cs.context().peerTakeWaitNs(); });
    return v;
}
int ogn_profile_enable(void *hh, int on) {
    return guard([&] {
        detail::DeviceContext &c = ((Handle *)hh)->cs.context();
        c.profileReset();
        c.profiling = on != 0;
        c.profilePeriod = on > 1 ? on : 1;
        c.sampleThis = c.profiling;
    });
}
int ogn_profile_read(void *hh, double *ms, int64_t *launches, int64_t *timed_launches) {
    return guard([&] {
        detail::DeviceContext &c = ((Handle *)hh)->cs.context();
        c.profileCollect();
        for (int k = 0; k < detail::kKernelKinds; k++) {
            if (ms) ms[k] = c.kernelMs[k];
            if (launches) launches[k] = c.launches[k];
            if (timed_launches) timed_launches[k] = c.timedLaunches[k];
        }
    });
}
// --- debug / test helpers (declared in ogn_b200.h) ---------------------------------------------------------------
int64_t ogn_debug_sc_fr_mismatch(void *hh, int l, int vli) {
    int64_t n = -1;
    guard([&] { n = (int64_t)((Handle *)hh)->h.getSCLayer(l).debugFRMismatch(vli); });
    return n;
}
int64_t ogn_debug_column_weights(void *hh, int kind, int l, int vli, int ox, int oy, float *out, int64_t cap) {
    int64_t n = -1;
    guard([&] {
        std::vector<float> w;
        Hierarchy &h = ((Handle *)hh)->h;
        if (kind == 0) h.getSCLayer(l).getColumnWeights(vli, ox, oy, w);
        else if (kind >= 1 && kind - 1 < (int)h.getPLayers(l).size() && h.getPLayers(l)[kind - 1])
            h.getPLayers(l)[kind - 1]->getColumnWeights(vli, ox, oy, w);
        else throw std::runtime_error("ogn_debug_column_weights: kind must be 0 (SparseCoder) or 1 + predictor index");
        n = weightsOut(w, out, cap);
    });
    return n;
}
int ogn_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

} // extern "C"

// ---- ImageEncoder (include/ogn_image_encoder_api.h) -------------------------------------------------------------------
namespace {
struct IEncHandle {
    ComputeSystem cs;
    ImageEncoder e;
};
} // namespace
extern "C" {
const char *ogn_ienc_last_error(void) { return g_err.c_str(); }
int ogn_ienc_create(int hx, int hy, int hz, int n, const ogn_ienc_visible_desc *d, uint32_t seed, void **out) {
    *out = nullptr;
    return guard([&] {
        std::unique_ptr<IEncHandle> H(new IEncHandle());
        H->cs.rng.seed(seed);
        std::vector<ImageEncoder::VisibleLayerDesc> vlds((size_t)n);
        for (int v = 0; v < n; v++) { vlds[(size_t)v].size = Int3(d[v].size_x, d[v].size_y, d[v].size_z); vlds[(size_t)v].radius = d[v].radius; }
        H->e.initRandom(H->cs, Int3(hx, hy, hz), vlds);
        *out = H.release();
    });
}
int ogn_ienc_create_on(void *hierarchy, int hx, int hy, int hz, int n, const ogn_ienc_visible_desc *d, uint32_t seed, void **out) {
    *out = nullptr;
    return guard([&] {
        if (!hierarchy) throw std::invalid_argument("null hierarchy handle");
        std::unique_ptr<IEncHandle> H(new IEncHandle());
        H->cs.shareContextWith(((Handle *)hierarchy)->cs);
        H->cs.rng.seed(seed);
        std::vector<ImageEncoder::VisibleLayerDesc> vlds((size_t)n);
        for (int v = 0; v < n; v++) { vlds[(size_t)v].size = Int3(d[v].size_x, d[v].size_y, d[v].size_z); vlds[(size_t)v].radius = d[v].radius; }
        H->e.initRandom(H->cs, Int3(hx, hy, hz), vlds);
        *out = H.release();
    });
}
int ogn_ienc_step_device(void *e, const float *const *in, int learn) {
    return guard([&] { IEncHandle *H = (IEncHandle *)e; H->e.stepDevice(H->cs, in, learn != 0); });
}
const int *ogn_ienc_hidden_cs_device(void *e) { return ((IEncHandle *)e)->e.getHiddenCsDevice(); }
int ogn_ienc_synchronize(void *e) {
    return guard([&] { ((IEncHandle *)e)->cs.synchronize(); });
}
void ogn_ienc_destroy(void *e) {
    try { delete (IEncHandle *)e; } catch (...) {}
}
int ogn_ienc_step(void *e, const float *const *in, int learn) {
    return guard([&] { IEncHandle *H = (IEncHandle *)e; H->e.stepHost(H->cs, in, learn != 0); });
}
int ogn_ienc_reconstruct(void *e, const int *hcs) {
    return guard([&] { IEncHandle *H = (IEncHandle *)e; H->e.reconstructHost(H->cs, hcs); });
}
int ogn_ienc_hidden_cs(void *e, int *out) {
    int n = -1;
    guard([&] { n = copyOut(((IEncHandle *)e)->e.getHiddenCs(), out); });
    return n;
}
int ogn_ienc_hidden_resources(void *e, float *out) {
    int n = -1;
    guard([&] { n = copyOut(((IEncHandle *)e)->e.getHiddenResources(), out); });
    return n;
}
int ogn_ienc_reconstruction(void *e, int vli, float *out) {
    int n = -1;
    guard([&] {
        FloatBuffer r;
        ((IEncHandle *)e)->e.getReconstruction(vli, r);
        n = copyOut(r, out);
    });
    return n;
}
int64_t ogn_ienc_weights(void *e, int vli, float *out, int64_t cap) {
    int64_t n = -1;
    guard([&] {
        std::vector<float> w;
        ((IEncHandle *)e)->e.getWeights(vli, w);
        n = weightsOut(w, out, cap);
    });
    return n;
}
int ogn_ienc_set_params(void *e, float alpha, float gamma) { ((IEncHandle *)e)->e.alpha = alpha; ((IEncHandle *)e)->e.gamma = gamma; return 0; }
int64_t ogn_ienc_save(void *e, const char *path) {
    int64_t n = -1;
    guard([&] {
        std::ofstream os(path, std::ios::binary);
        if (!os) throw std::runtime_error("cannot open file for writing");
        ((IEncHandle *)e)->e.writeToStream(os);
        n = (int64_t)os.tellp();
    });
    return n;
}
int ogn_ienc_load(const char *path, void **out) {
    *out = nullptr;
    return guard([&] {
        std::ifstream is(path, std::ios::binary);
        if (!is) throw std::runtime_error("cannot open file");
        std::unique_ptr<IEncHandle> H(new IEncHandle());
        H->e.readFromStream(is, H->cs);
        *out = H.release();
    });
}
} // extern "C"


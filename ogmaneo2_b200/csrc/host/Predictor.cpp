// csrc/host/Predictor.cpp — ogmaneo::Predictor on the device (public interface: include/ogmaneo/Predictor.h).
#include "LayerImpl.h"

#include <cstdlib>

namespace ogmaneo {
using namespace detail;

Predictor::Predictor() : alpha(0.5f), impl(nullptr) {}
Predictor::~Predictor() { delete impl; }
Predictor::Predictor(const Predictor &o) : alpha(0.5f), impl(nullptr) { *this = o; }

Predictor &Predictor::operator=(const Predictor &o) {
    if (this == &o) return *this;
    alpha = o.alpha; hiddenSize = o.hiddenSize; visibleLayerDescs = o.visibleLayerDescs;
    hiddenActivations = o.hiddenActivations; hiddenCs = o.hiddenCs; visibleLayers = o.visibleLayers;
    delete impl; impl = nullptr;
    if (o.impl) {
        impl = new Impl();
        const Impl &s = *o.impl;
        DeviceContext &ctx = *s.ctx;
        cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
        ctx.joinSide();
        impl->ctx = s.ctx; impl->batch = s.batch; impl->x0 = s.x0; impl->x1 = s.x1; impl->world = s.world;
        impl->prevCur = s.prevCur; impl->csDirty = s.csDirty; impl->actDirty = s.actDirty; impl->prevDirty = s.prevDirty;
        impl->wDirty = s.wDirty;
        impl->ws.cloneFrom(ctx, s.ws);
        impl->acts.cloneFrom(s.acts, ctx.stream); impl->hiddenCs.cloneFrom(s.hiddenCs, ctx.stream);
        for (int p = 0; p < 2; p++) {
            impl->prev_[p].resize(s.prev_[p].size());
            for (size_t v = 0; v < s.prev_[p].size(); v++) impl->prev_[p][v].cloneFrom(s.prev_[p][v], ctx.stream);
        }
        impl->inStage.resize(s.inStage.size());
        ctx.sync();
    }
    return *this;
}

void Predictor::Impl::allocate(ComputeSystem &cs, const Int3 &hidden, const std::vector<VisibleLayerDesc> &vlds, int batch_) {
    ctx = cs.sharedContext();
    batch = batch_; world = cs.shardWorld;
    slabOf(hidden.x, cs.shardRank, cs.shardWorld, x0, x1);
    ws.sets.clear(); ws.sets.resize(vlds.size());
    for (size_t v = 0; v < vlds.size(); v++)
        ws.sets[v].allocate(*ctx, Geometry::get(hidden, vlds[v].size, vlds[v].radius), batch, false, false, x0, x1);
    const size_t ncols = (size_t)hidden.x * hidden.y * batch;
    acts.allocZero(ncols * hidden.z, ctx->stream);
    if (ctx->arena) hiddenCs.borrow(ctx->arenaAlloc(ncols), ncols); // exchanged CSDR: see SparseCoder::Impl::allocate
    else hiddenCs.allocZero(ncols, ctx->stream);
    for (int p = 0; p < 2; p++) {
        prev_[p].clear(); prev_[p].resize(vlds.size());
        for (size_t v = 0; v < vlds.size(); v++) prev_[p][v].allocZero((size_t)vlds[v].size.x * vlds[v].size.y * batch, ctx->stream);
    }
    inStage.clear(); inStage.resize(vlds.size());
    prevCur = 0;
    csDirty = actDirty = prevDirty = true; wDirty.assign(vlds.size(), 1);
}

// Predictor.cpp:72-109
void Predictor::initRandom(ComputeSystem &cs, const Int3 &hiddenSize_, const std::vector<VisibleLayerDesc> &vlds) {
    initRandomMember(cs, hiddenSize_, vlds, 1, 0);
}

void Predictor::initRandomMember(ComputeSystem &cs, const Int3 &hiddenSize_, const std::vector<VisibleLayerDesc> &vlds, int batch,
                                 int member) {
    if (member == 0) {
        hiddenSize = hiddenSize_; visibleLayerDescs = vlds;
        visibleLayers.clear(); visibleLayers.resize(vlds.size());
        delete impl; impl = new Impl();
        impl->allocate(cs, hiddenSize, vlds, batch);
    }
    for (size_t v = 0; v < vlds.size(); v++)
        impl->ws.sets[v].initialise(*impl->ctx, member, WeightInit{&cs, -0.01f, 0.01f, cs.deviceInitMatrix++});
}

// Predictor::learn (Predictor.cpp:129-135) followed by Predictor::activate (:111-127) as one launch (K6+K7+K8)
void Predictor::stepDevice(ComputeSystem &cs, bool learn, const int *targetDev, bool activate, const int *const *inputDev,
                           bool deferExchange, bool onSideStream) {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    const int nvl = (int)visibleLayerDescs.size();
    if (nvl > 8) throw std::runtime_error("ogmaneo-b200: more than 8 visible layers on one Predictor");
    ogn_pred_args a;
    std::memset(&a, 0, sizeof(a));
    int ngeo = 0;
    a.nvl = nvl;
    for (int v = 0; v < nvl; v++) {
        const WeightSet &w = m.ws.sets[v];
        a.vl[v].wf = w.fBase(); a.vl[v].mstride = w.fCount(); a.vl[v].geo = geoIndex(a.geo, ngeo, *w.geo);
        a.vl[v].cs = activate ? inputDev[v] : nullptr;
        a.vl[v].cs_prev = m.prev_[m.prevCur][v].ptr();
        a.vl[v].cs_prev_out = activate ? m.prev_[1 - m.prevCur][v].ptr() : nullptr;
    }
    if (learn) consumeKernel2(cs, hiddenSize.x, hiddenSize.y);
    if (activate) {
        consumeKernel2(cs, hiddenSize.x, hiddenSize.y);
        for (int v = 0; v < nvl; v++) consumeKernel1(cs, visibleLayerDescs[v].size.x * visibleLayerDescs[v].size.y);
    }
    // onSideStream: a learn-only pass the caller hoisted to the side stream (see Hierarchy::step)
    cudaStream_t st = (onSideStream && ctx.useSide && !activate) ? ctx.side : ctx.stream;
    if (ctx.collect) { // single-launch step: record the phase instead of launching it
        ogn_small_phase P;
        std::memset(&P, 0, sizeof(P));
        const int pf = Geometry::packFactor(hiddenSize.z);
        P.kind = OGN_PHASE_PRED; P.z = hiddenSize.z; P.Hx = hiddenSize.x; P.Hy = hiddenSize.y;
        P.ppr = (hiddenSize.y + pf - 1) / pf; P.npacks = hiddenSize.x * P.ppr;
        P.learn = learn ? 1 : 0; P.activate = activate ? 1 : 0; P.target = targetDev; P.alpha = alpha;
        P.acts = m.acts.ptr(); P.out0 = m.hiddenCs.ptr(); P.u.pred = a;
        bool ok = ogn_vec_row_ok(hiddenSize.z) != 0 && m.world == 1;
        for (int v = 0; v < nvl && ok; v++) ok = ogn_vec_geo_ok(&a.geo[a.vl[v].geo]) != 0;
        ctx.collect->ok = ctx.collect->ok && ok;
        ctx.collect->add(P);
    } else {
    // OGN_PRED_SPLIT=1 (experiment, profiles/README.md): the learn pass (a pure read-modify-write of 128-byte lines) and the
    // activate pass (a pure gather) as two launches instead of one kernel that mixes reads and writes 2:1
    static const bool split = [] { const char *e = std::getenv("OGN_PRED_SPLIT"); return e && std::atoi(e) != 0; }();
    if (split && learn && activate) {
        { LaunchScope ls(ctx, kPredStep, st);
        if (!ctx.dry) kernelCheck(ogn_pred_step(&a, hiddenSize.x, hiddenSize.y, hiddenSize.z, m.x0, m.x1, m.batch, 1, 0, targetDev, alpha,
                                  m.acts.ptr(), m.hiddenCs.ptr(), st), "pred_step(learn)"); }
        { LaunchScope ls(ctx, kPredStep, st);
        if (!ctx.dry) kernelCheck(ogn_pred_step(&a, hiddenSize.x, hiddenSize.y, hiddenSize.z, m.x0, m.x1, m.batch, 0, 1, targetDev, alpha,
                                  m.acts.ptr(), m.hiddenCs.ptr(), st), "pred_step(activate)"); }
    } else {
    LaunchScope ls(ctx, kPredStep, st);
    if (!ctx.dry) kernelCheck(ogn_pred_step(&a, hiddenSize.x, hiddenSize.y, hiddenSize.z, m.x0, m.x1, m.batch, learn ? 1 : 0, activate ? 1 : 0,
                              targetDev, alpha, m.acts.ptr(), m.hiddenCs.ptr(), st), "pred_step"); } }
    if (learn) std::fill(m.wDirty.begin(), m.wDirty.end(), 1);
    if (activate) {
        m.prevCur = 1 - m.prevCur;
        m.csDirty = m.actDirty = m.prevDirty = true;
        if (m.world > 1) {
            if (ctx.arena && ctx.inArena(m.hiddenCs.ptr())) ctx.peerAllGather(m.hiddenCs.ptr(), (m.x1 - m.x0) * hiddenSize.y, deferExchange);
            else if (cs.allGather) cs.allGather(cs.allGatherUser, m.hiddenCs.ptr(), (m.x1 - m.x0) * hiddenSize.y, (void *)ctx.stream);
            else throw std::runtime_error("ogmaneo-b200: shardWorld > 1 needs ComputeSystem::allGather or peerExchange");
        }
    }
}

void Predictor::activate(ComputeSystem &cs, const std::vector<const IntBuffer *> &inputCs) {
    Impl &m = *impl;
    cudaCheck(cudaSetDevice(m.ctx->device), "cudaSetDevice");
    std::vector<const int *> ptrs(inputCs.size());
    for (size_t v = 0; v < inputCs.size(); v++) {
        if (m.inStage[v].size() != inputCs[v]->size()) m.inStage[v].alloc(inputCs[v]->size());
        m.inStage[v].upload(inputCs[v]->data(), inputCs[v]->size(), m.ctx->stream);
        ptrs[v] = m.inStage[v].ptr();
    }
    stepDevice(cs, false, nullptr, true, ptrs.data());
    m.ctx->sync();
}

void Predictor::learn(ComputeSystem &cs, const IntBuffer *hiddenTargetCs) {
    Impl &m = *impl;
    cudaCheck(cudaSetDevice(m.ctx->device), "cudaSetDevice");
    if (m.targetStage.size() != hiddenTargetCs->size()) m.targetStage.alloc(hiddenTargetCs->size());
    m.targetStage.upload(hiddenTargetCs->data(), hiddenTargetCs->size(), m.ctx->stream);
    stepDevice(cs, true, m.targetStage.ptr(), false, nullptr);
    m.ctx->sync();
}

// --- host mirrors ------------------------------------------------------------------------------------------------
namespace {
// device hiddenCs -> pinned staging, waits; returns the staging pointer
const int *stageHiddenCs(Predictor::Impl &m) {
    cudaCheck(cudaSetDevice(m.ctx->device), "cudaSetDevice");
    m.ctx->joinExchange(); // a deferred exchange of this CSDR may still be running on the side stream
    if (m.csPinned.size() != m.hiddenCs.size()) m.csPinned.alloc(m.hiddenCs.size());
    m.hiddenCs.download(m.csPinned.ptr(), m.hiddenCs.size(), m.ctx->stream);
    cudaCheck(cudaStreamSynchronize(m.ctx->stream), "cudaStreamSynchronize");
    m.ctx->peerCheck();
    return m.csPinned.ptr();
}
} // namespace

const IntBuffer &Predictor::getHiddenCs() const {
    Impl &m = *impl;
    if (m.csDirty) {
        const int *p = stageHiddenCs(m);
        hiddenCs.assign(p, p + m.hiddenCs.size());
        m.csDirty = false;
    }
    return hiddenCs;
}

void Predictor::hiddenRows(int &lo, int &hi) const { lo = impl->x0; hi = impl->x1; }
void Predictor::visibleRows(int vli, int &lo, int &hi) const { lo = impl->ws.sets[vli].vx0; hi = impl->ws.sets[vli].vx1; }

int Predictor::copyHiddenCs(int *out) const {
    Impl &m = *impl;
    const size_t n = m.hiddenCs.size();
    if (!out) return (int)n; // size query
    if (!m.csDirty) std::memcpy(out, hiddenCs.data(), n * sizeof(int));
    else std::memcpy(out, stageHiddenCs(m), n * sizeof(int));
    return (int)n;
}

const FloatBuffer &Predictor::getHiddenActivations() const {
    Impl &m = *impl;
    if (m.actDirty) {
        cudaCheck(cudaSetDevice(m.ctx->device), "cudaSetDevice");
        hiddenActivations.resize(m.acts.size());
        m.acts.download(hiddenActivations.data(), hiddenActivations.size(), m.ctx->stream);
        m.ctx->sync();
        m.actDirty = false;
    }
    return hiddenActivations;
}

const IntBuffer &Predictor::getInputCsPrev(int i) const {
    Impl &m = *impl;
    if (m.prevDirty) {
        cudaCheck(cudaSetDevice(m.ctx->device), "cudaSetDevice");
        for (size_t v = 0; v < visibleLayers.size(); v++) {
            IntBuffer &b = visibleLayers[v].inputCsPrev;
            b.resize(m.prev_[m.prevCur][v].size());
            m.prev_[m.prevCur][v].download(b.data(), b.size(), m.ctx->stream);
        }
        m.ctx->sync();
        m.prevDirty = false;
    }
    return visibleLayers[i].inputCsPrev;
}

void Predictor::getWeights(int i, std::vector<float> &out) const {
    cudaCheck(cudaSetDevice(impl->ctx->device), "cudaSetDevice");
    const WeightSet &w = impl->ws.sets[i];
    out.assign((size_t)w.geo->nnz, 0.0f);
    w.downloadCSR(*impl->ctx, 0, out.data());
}

void Predictor::getColumnWeights(int i, int ox, int oy, std::vector<float> &out) const {
    DeviceContext &ctx = *impl->ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    const WeightSet &w = impl->ws.sets[i];
    const Geometry &g = *w.geo;
    if (ox < w.fx0 || ox >= w.fx1) throw std::out_of_range("ogmaneo-b200: column not stored on this device");
    const int64_t n = (int64_t)g.slots(ox, oy) * g.Vz * g.Hz, base = g.colBlock(ox, oy);
    const ogn_wset d = w.desc();
    DevBuf<float> stage;
    stage.alloc((size_t)n);
    const int col = oy + ox * g.Hy;
    ctx.joinSide();
    kernelCheck(ogn_weights_to_csr(&d, &g.host, 0, 0, col, col + 1, stage.ptr(), base, ctx.stream), "weights_to_csr");
    out.resize((size_t)n);
    stage.download(out.data(), (size_t)n, ctx.stream);
    ctx.sync();
}

const Predictor::VisibleLayer &Predictor::getVisibleLayer(int i) const {
    VisibleLayer &vl = visibleLayers[i];
    if (impl->wDirty[i]) {
        if (vl.weights.rowRanges.empty()) initSMLocalRF(visibleLayerDescs[i].size, hiddenSize, visibleLayerDescs[i].radius, vl.weights);
        getWeights(i, vl.weights.nonZeroValues);
        impl->wDirty[i] = 0;
    }
    getInputCsPrev(i);
    return vl;
}

// --- streams: Predictor.cpp:137-192 ----------------------------------------------------------------------------------
void Predictor::writeToStream(std::ostream &os) const {
    if (impl->world > 1 || impl->batch > 1) throw std::runtime_error("ogmaneo-b200: writeToStream needs an unsharded single hierarchy");
    os.write(reinterpret_cast<const char *>(&hiddenSize), sizeof(Int3));
    os.write(reinterpret_cast<const char *>(&alpha), sizeof(float));
    writeBufferToStream(os, &getHiddenActivations());
    writeBufferToStream(os, &getHiddenCs());
    int n = (int)visibleLayerDescs.size();
    os.write(reinterpret_cast<const char *>(&n), sizeof(int));
    for (int v = 0; v < n; v++) {
        const VisibleLayer &vl = getVisibleLayer(v);
        os.write(reinterpret_cast<const char *>(&visibleLayerDescs[v]), sizeof(VisibleLayerDesc));
        writeSMToStream(os, vl.weights);
        writeBufferToStream(os, &vl.inputCsPrev);
    }
}

void Predictor::readFromStream(std::istream &is, ComputeSystem &cs) {
    is.read(reinterpret_cast<char *>(&hiddenSize), sizeof(Int3));
    hiddenSize.pad = 0;
    is.read(reinterpret_cast<char *>(&alpha), sizeof(float));
    FloatBuffer act; IntBuffer hc;
    readBufferFromStream(is, &act);
    readBufferFromStream(is, &hc);
    int n = 0;
    is.read(reinterpret_cast<char *>(&n), sizeof(int));
    visibleLayerDescs.resize(n);
    std::vector<SparseMatrix> mats(n);
    std::vector<IntBuffer> prev(n);
    for (int v = 0; v < n; v++) {
        is.read(reinterpret_cast<char *>(&visibleLayerDescs[v]), sizeof(VisibleLayerDesc));
        visibleLayerDescs[v].size.pad = 0;
        readSMFromStream(is, mats[v]);
        readBufferFromStream(is, &prev[v]);
    }
    visibleLayers.clear(); visibleLayers.resize(n);
    delete impl; impl = new Impl();
    impl->allocate(cs, hiddenSize, visibleLayerDescs, 1);
    Impl &m = *impl;
    for (int v = 0; v < n; v++) {
        if ((int64_t)mats[v].nonZeroValues.size() != m.ws.sets[v].geo->nnz)
            throw std::runtime_error("ogmaneo-b200: Predictor weight count in stream does not match its geometry");
        m.ws.sets[v].uploadCSR(*m.ctx, 0, mats[v].nonZeroValues.data());
    }
    if (act.size() == m.acts.size()) m.acts.upload(act.data(), act.size(), m.ctx->stream);
    m.ctx->sync();
    setState(hc, prev);
}

void Predictor::readFromStream(std::istream &is) {
    ComputeSystem cs;
    if (impl) cs.device = impl->ctx->device;
    readFromStream(is, cs);
}

void Predictor::setState(const IntBuffer &hc, const std::vector<IntBuffer> &prev) {
    Impl &m = *impl;
    cudaCheck(cudaSetDevice(m.ctx->device), "cudaSetDevice");
    if (hc.size() != m.hiddenCs.size()) throw std::invalid_argument("ogmaneo-b200: Predictor::setState: hiddenCs size mismatch");
    if (prev.size() != m.prev_[m.prevCur].size()) throw std::invalid_argument("ogmaneo-b200: Predictor::setState: visible layer count mismatch");
    for (size_t v = 0; v < prev.size(); v++)
        if (prev[v].size() != m.prev_[m.prevCur][v].size())
            throw std::invalid_argument("ogmaneo-b200: Predictor::setState: inputCsPrev size mismatch");
    m.ctx->joinSide(); // a deferred exchange may still be writing hiddenCs
    m.hiddenCs.upload(hc.data(), hc.size(), m.ctx->stream);
    for (size_t v = 0; v < prev.size(); v++) m.prev_[m.prevCur][v].upload(prev[v].data(), prev[v].size(), m.ctx->stream);
    m.ctx->sync();
    m.csDirty = m.prevDirty = true;
}

const int *Predictor::hiddenCsDevice() const { return impl->hiddenCs.ptr(); }

void Predictor::writeDeviceState(std::ostream &os) const {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    const int hdr[3] = {m.prevCur, m.x0, m.x1};
    os.write(reinterpret_cast<const char *>(hdr), sizeof(hdr));
    os.write(reinterpret_cast<const char *>(&alpha), sizeof(alpha));
    writeDev(os, ctx, m.acts); writeDev(os, ctx, m.hiddenCs);
    for (int p = 0; p < 2; p++)
        for (const auto &b : m.prev_[p]) writeDev(os, ctx, b);
    for (const WeightSet &w : m.ws.sets) writeDev(os, ctx, w.wf);
}

void Predictor::readDeviceState(std::istream &is) {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    int hdr[3] = {0, 0, 0};
    is.read(reinterpret_cast<char *>(hdr), sizeof(hdr));
    if (hdr[1] != m.x0 || hdr[2] != m.x1) throw std::runtime_error("ogmaneo-b200: checkpoint was written by a different shard");
    m.prevCur = hdr[0];
    is.read(reinterpret_cast<char *>(&alpha), sizeof(alpha));
    readDev(is, ctx, m.acts); readDev(is, ctx, m.hiddenCs);
    for (int p = 0; p < 2; p++)
        for (auto &b : m.prev_[p]) readDev(is, ctx, b);
    for (WeightSet &w : m.ws.sets) readDev(is, ctx, w.wf);
    m.csDirty = m.actDirty = m.prevDirty = true;
    std::fill(m.wDirty.begin(), m.wDirty.end(), 1);
}

int Predictor::smallUnits() const {
    const Impl &m = *impl;
    if (m.world > 1 || !ogn_vec_row_ok(hiddenSize.z)) return -1;
    for (const WeightSet &w : m.ws.sets)
        if (!ogn_vec_geo_ok(&w.geo->host)) return -1;
    const int pf = Geometry::packFactor(hiddenSize.z);
    return hiddenSize.x * ((hiddenSize.y + pf - 1) / pf);
}

void Predictor::stateSignature(std::vector<int> &key) const {
    key.push_back(impl->prevCur);
    int a;
    std::memcpy(&a, &alpha, sizeof(a));
    key.push_back(a);
}

} // namespace ogmaneo

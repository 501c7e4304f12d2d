// csrc/host/Actor.cpp — ogmaneo::Actor on the device (public interface: include/ogmaneo/Actor.h).
#include "LayerImpl.h"

namespace ogmaneo {
using namespace detail;

Actor::Actor() : alpha(0.02f), beta(0.02f), gamma(0.99f), minSteps(8), historyIters(8), impl(nullptr) {}
Actor::~Actor() { delete impl; }
Actor::Actor(const Actor &o) : impl(nullptr) { *this = o; }

Actor &Actor::operator=(const Actor &o) {
    if (this == &o) return *this;
    alpha = o.alpha; beta = o.beta; gamma = o.gamma; minSteps = o.minSteps; historyIters = o.historyIters;
    hiddenSize = o.hiddenSize; visibleLayerDescs = o.visibleLayerDescs;
    hiddenCs = o.hiddenCs; hiddenValues = o.hiddenValues; visibleLayers = o.visibleLayers;
    delete impl; impl = nullptr;
    if (o.impl) {
        impl = new Impl();
        const Impl &s = *o.impl;
        DeviceContext &ctx = *s.ctx;
        cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
        impl->ctx = s.ctx; impl->capacity = s.capacity; impl->start = s.start; impl->historySize = s.historySize;
        impl->rewards = s.rewards; impl->csDirty = s.csDirty; impl->valDirty = s.valDirty; impl->wDirty = s.wDirty;
        impl->vw.cloneFrom(ctx, s.vw); impl->aw.cloneFrom(ctx, s.aw);
        impl->acts.cloneFrom(s.acts, ctx.stream); impl->values.cloneFrom(s.values, ctx.stream);
        impl->u01.alloc(s.u01.size()); impl->u01Host.alloc(s.u01Host.size());
        impl->hiddenCs.cloneFrom(s.hiddenCs, ctx.stream);
        impl->histInput.resize(s.histInput.size());
        for (size_t v = 0; v < s.histInput.size(); v++) impl->histInput[v].cloneFrom(s.histInput[v], ctx.stream);
        impl->histTarget.cloneFrom(s.histTarget, ctx.stream); impl->histValues.cloneFrom(s.histValues, ctx.stream);
        impl->inStage.resize(s.inStage.size());
        for (auto &e : impl->u01Free) {
            cudaCheck(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
            cudaCheck(cudaEventRecord(e, ctx.stream), "cudaEventRecord");
        }
        ctx.sync();
    }
    return *this;
}

void Actor::Impl::allocate(ComputeSystem &cs, const Int3 &hidden, int historyCapacity, const std::vector<VisibleLayerDesc> &vlds) {
    if (cs.shardWorld > 1) throw std::runtime_error("ogmaneo-b200: Actor layers cannot be sharded");
    ctx = cs.sharedContext();
    const int ncols = hidden.x * hidden.y;
    vw.sets.clear(); vw.sets.resize(vlds.size()); aw.sets.clear(); aw.sets.resize(vlds.size());
    for (size_t v = 0; v < vlds.size(); v++) {
        vw.sets[v].allocate(*ctx, Geometry::get(Int3(hidden.x, hidden.y, 1), vlds[v].size, vlds[v].radius), 1, false, false, 0, hidden.x);
        aw.sets[v].allocate(*ctx, Geometry::get(hidden, vlds[v].size, vlds[v].radius), 1, false, false, 0, hidden.x);
    }
    acts.allocZero((size_t)ncols * hidden.z, ctx->stream);
    values.allocZero(ncols, ctx->stream);
    hiddenCs.allocZero(ncols, ctx->stream);
    u01.allocZero(ncols, ctx->stream); u01Host.alloc((size_t)ncols * kU01Slots);
    capacity = historyCapacity; start = 0; historySize = 0;
    histInput.clear(); histInput.resize(vlds.size());
    for (size_t v = 0; v < vlds.size(); v++) histInput[v].allocZero((size_t)capacity * vlds[v].size.x * vlds[v].size.y, ctx->stream);
    histTarget.allocZero((size_t)capacity * ncols, ctx->stream);
    histValues.allocZero((size_t)capacity * ncols, ctx->stream);
    rewards.assign(capacity, 0.0f);
    inStage.clear(); inStage.resize(vlds.size());
    for (auto &e : u01Free) {
        if (!e) cudaCheck(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
        cudaCheck(cudaEventRecord(e, ctx->stream), "cudaEventRecord");
    }
    csDirty = valDirty = true; wDirty.assign(vlds.size(), 1);
}

// Actor.cpp:187-246 — per visible layer: value weights, then action weights, U(-0.01,0.01)
void Actor::initRandom(ComputeSystem &cs, const Int3 &hiddenSize_, int historyCapacity, const std::vector<VisibleLayerDesc> &vlds) {
    hiddenSize = hiddenSize_; visibleLayerDescs = vlds;
    visibleLayers.clear(); visibleLayers.resize(vlds.size());
    delete impl; impl = new Impl();
    impl->allocate(cs, hiddenSize, historyCapacity, vlds);
    for (size_t v = 0; v < vlds.size(); v++) {
        impl->vw.sets[v].initialise(*impl->ctx, 0, WeightInit{&cs, -0.01f, 0.01f, cs.deviceInitMatrix++});
        impl->aw.sets[v].initialise(*impl->ctx, 0, WeightInit{&cs, -0.01f, 0.01f, cs.deviceInitMatrix++});
    }
}

// Actor::step, Actor.cpp:248-313
void Actor::stepDevice(ComputeSystem &cs, const int *const *inputDev, const int *targetPrevDev, float reward, bool learnEnabled,
                       bool mimic) {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    const int nvl = (int)visibleLayerDescs.size();
    if (nvl > 8) throw std::runtime_error("ogmaneo-b200: more than 8 visible layers on one Actor");
    const int Hx = hiddenSize.x, Hy = hiddenSize.y, Hz = hiddenSize.z, ncols = Hx * Hy;

    // --- forward (K9). Replay the per-tile generators: tile i covers x in [bx*2, ..), y in [by*2, ..), bx = i % nbx,
    // by = i / nbx, columns visited x-outer / y-inner, one canonical float each (Helpers.cpp:46-70, Actor.cpp:70-72).
    std::vector<int> seeds;
    consumeKernel2(cs, Hx, Hy, &seeds);
    const int us = m.u01Slot;
    m.u01Slot = (m.u01Slot + 1) % Impl::kU01Slots;
    float *u01h = m.u01Host.ptr() + (size_t)us * ncols;
    cudaCheck(cudaEventSynchronize(m.u01Free[us]), "cudaEventSynchronize");
    const int nbx = (Hx + cs.batchSize2.x - 1) / cs.batchSize2.x;
    for (size_t i = 0; i < seeds.size(); i++) {
        std::mt19937 sub(seeds[i]);
        const int px = (int)(i % nbx) * cs.batchSize2.x, py = (int)(i / nbx) * cs.batchSize2.y;
        for (int x = px; x < std::min(Hx, px + cs.batchSize2.x); x++)
            for (int y = py; y < std::min(Hy, py + cs.batchSize2.y); y++)
                u01h[y + x * Hy] = std::generate_canonical<float, std::numeric_limits<float>::digits>(sub);
    }
    cudaCheck(cudaMemcpyAsync(m.u01.ptr(), u01h, ncols * sizeof(float), cudaMemcpyHostToDevice, ctx.stream), "u01 H2D");
    cudaCheck(cudaEventRecord(m.u01Free[us], ctx.stream), "cudaEventRecord");

    ogn_actor_args a;
    std::memset(&a, 0, sizeof(a));
    int ngeo = 0;
    a.nvl = nvl;
    for (int v = 0; v < nvl; v++) {
        a.vl[v].wv = m.vw.sets[v].fBase(); a.vl[v].wa = m.aw.sets[v].fBase(); a.vl[v].cs = inputDev[v];
        a.vl[v].geo_v = geoIndex(a.geo, ngeo, *m.vw.sets[v].geo);
        a.vl[v].geo_a = geoIndex(a.geo, ngeo, *m.aw.sets[v].geo);
    }
    { LaunchScope ls(ctx, kActorForward);
    kernelCheck(ogn_actor_forward(&a, Hx, Hy, Hz, m.u01.ptr(), m.values.ptr(), m.acts.ptr(), m.hiddenCs.ptr(), ctx.stream),
                "actor_forward"); }
    m.csDirty = m.valDirty = true;

    // --- push the history sample (K10)
    m.start = (m.start - 1 + m.capacity) % m.capacity;
    if (m.historySize < m.capacity) m.historySize++;
    const int slot = m.start;
    for (int v = 0; v < nvl; v++) {
        const size_t n = (size_t)visibleLayerDescs[v].size.x * visibleLayerDescs[v].size.y;
        consumeKernel1(cs, (int)n);
        cudaCheck(cudaMemcpyAsync(m.histInput[v].ptr() + slot * n, inputDev[v], n * sizeof(int), cudaMemcpyDeviceToDevice, ctx.stream), "history D2D");
    }
    consumeKernel1(cs, ncols);
    cudaCheck(cudaMemcpyAsync(m.histTarget.ptr() + (size_t)slot * ncols, targetPrevDev, ncols * sizeof(int), cudaMemcpyDeviceToDevice, ctx.stream), "history D2D");
    consumeKernel1(cs, ncols);
    cudaCheck(cudaMemcpyAsync(m.histValues.ptr() + (size_t)slot * ncols, m.values.ptr(), ncols * sizeof(float), cudaMemcpyDeviceToDevice, ctx.stream), "history D2D");
    m.rewards[slot] = reward;

    // --- learn on historyIters random past samples (K11)
    if (learnEnabled && m.historySize > minSteps + 1) {
        std::uniform_int_distribution<int> historyDist(minSteps, m.historySize - 2);
        ogn_actor_batch batch;
        std::memset(&batch, 0, sizeof(batch));
        auto flush = [&]() {
            if (batch.n == 0) return;
            LaunchScope ls(ctx, kActorLearn);
            kernelCheck(ogn_actor_learn(&a, &batch, Hx, Hy, Hz, m.values.ptr(), alpha, beta, mimic ? 1 : 0, m.acts.ptr(), ctx.stream),
                        "actor_learn");
            batch.n = 0;
        };
        for (int it = 0; it < historyIters; it++) {
            const int idx = historyDist(cs.rng);
            const int sPrev = m.slot(idx + 1), s = m.slot(idx);
            float q = 0.0f, g = 1.0f;
            for (int t = idx; t >= 0; t--) {
                q += m.rewards[m.slot(t)] * g;
                g *= gamma;
            }
            consumeKernel2(cs, Hx, Hy);
            ogn_actor_sample &S = batch.it[batch.n++];
            for (int v = 0; v < nvl; v++)
                S.cs[v] = m.histInput[v].ptr() + (size_t)sPrev * visibleLayerDescs[v].size.x * visibleLayerDescs[v].size.y;
            S.target_prev = m.histTarget.ptr() + (size_t)s * ncols;
            S.values_prev = m.histValues.ptr() + (size_t)sPrev * ncols;
            S.q = q; S.g = g;
            if (batch.n == OGN_MAX_ACTOR_SAMPLES) flush();
        }
        flush();
        std::fill(m.wDirty.begin(), m.wDirty.end(), 1);
    }
}

void Actor::step(ComputeSystem &cs, const std::vector<const IntBuffer *> &inputCs, const IntBuffer *hiddenTargetCsPrev, float reward,
                 bool learnEnabled, bool mimic) {
    Impl &m = *impl;
    cudaCheck(cudaSetDevice(m.ctx->device), "cudaSetDevice");
    std::vector<const int *> ptrs(inputCs.size());
    for (size_t v = 0; v < inputCs.size(); v++) {
        if (m.inStage[v].size() != inputCs[v]->size()) m.inStage[v].alloc(inputCs[v]->size());
        m.inStage[v].upload(inputCs[v]->data(), inputCs[v]->size(), m.ctx->stream);
        ptrs[v] = m.inStage[v].ptr();
    }
    if (m.targetStage.size() != hiddenTargetCsPrev->size()) m.targetStage.alloc(hiddenTargetCsPrev->size());
    m.targetStage.upload(hiddenTargetCsPrev->data(), hiddenTargetCsPrev->size(), m.ctx->stream);
    stepDevice(cs, ptrs.data(), m.targetStage.ptr(), reward, learnEnabled, mimic);
    m.ctx->sync();
}

// --- host mirrors ------------------------------------------------------------------------------------------------
const IntBuffer &Actor::getHiddenCs() const {
    Impl &m = *impl;
    if (m.csDirty) {
        cudaCheck(cudaSetDevice(m.ctx->device), "cudaSetDevice");
        hiddenCs.resize(m.hiddenCs.size());
        m.hiddenCs.download(hiddenCs.data(), hiddenCs.size(), m.ctx->stream);
        m.ctx->sync();
        m.csDirty = false;
    }
    return hiddenCs;
}

const FloatBuffer &Actor::getHiddenValues() const {
    Impl &m = *impl;
    if (m.valDirty) {
        cudaCheck(cudaSetDevice(m.ctx->device), "cudaSetDevice");
        hiddenValues.resize(m.values.size());
        m.values.download(hiddenValues.data(), hiddenValues.size(), m.ctx->stream);
        m.ctx->sync();
        m.valDirty = false;
    }
    return hiddenValues;
}

void Actor::getWeights(int i, int which, std::vector<float> &out) const {
    cudaCheck(cudaSetDevice(impl->ctx->device), "cudaSetDevice");
    const WeightSet &w = which == 0 ? impl->vw.sets[i] : impl->aw.sets[i];
    out.assign((size_t)w.geo->nnz, 0.0f);
    w.downloadCSR(*impl->ctx, 0, out.data());
}

const Actor::VisibleLayer &Actor::getVisibleLayer(int i) const {
    VisibleLayer &vl = visibleLayers[i];
    if (impl->wDirty[i]) {
        if (vl.valueWeights.rowRanges.empty()) {
            initSMLocalRF(visibleLayerDescs[i].size, Int3(hiddenSize.x, hiddenSize.y, 1), visibleLayerDescs[i].radius, vl.valueWeights);
            initSMLocalRF(visibleLayerDescs[i].size, hiddenSize, visibleLayerDescs[i].radius, vl.actionWeights);
        }
        getWeights(i, 0, vl.valueWeights.nonZeroValues);
        getWeights(i, 1, vl.actionWeights.nonZeroValues);
        impl->wDirty[i] = 0;
    }
    return vl;
}

// --- streams: Actor.cpp:315-438 ----------------------------------------------------------------------------------------
void Actor::writeToStream(std::ostream &os) const {
    Impl &m = *impl;
    os.write(reinterpret_cast<const char *>(&hiddenSize), sizeof(Int3));
    os.write(reinterpret_cast<const char *>(&alpha), sizeof(float));
    os.write(reinterpret_cast<const char *>(&beta), sizeof(float));
    os.write(reinterpret_cast<const char *>(&gamma), sizeof(float));
    os.write(reinterpret_cast<const char *>(&minSteps), sizeof(int));
    os.write(reinterpret_cast<const char *>(&historyIters), sizeof(int));
    writeBufferToStream(os, &getHiddenCs());
    writeBufferToStream(os, &getHiddenValues());
    int n = (int)visibleLayerDescs.size();
    os.write(reinterpret_cast<const char *>(&n), sizeof(int));
    for (int v = 0; v < n; v++) {
        const VisibleLayer &vl = getVisibleLayer(v);
        os.write(reinterpret_cast<const char *>(&visibleLayerDescs[v]), sizeof(VisibleLayerDesc));
        writeSMToStream(os, vl.valueWeights);
        writeSMToStream(os, vl.actionWeights);
    }
    os.write(reinterpret_cast<const char *>(&m.historySize), sizeof(int));
    os.write(reinterpret_cast<const char *>(&m.capacity), sizeof(int));
    os.write(reinterpret_cast<const char *>(&m.start), sizeof(int));
    cudaCheck(cudaSetDevice(m.ctx->device), "cudaSetDevice");
    const size_t ncols = (size_t)hiddenSize.x * hiddenSize.y;
    std::vector<IntBuffer> hin(n);
    for (int v = 0; v < n; v++) { hin[v].resize(m.histInput[v].size()); m.histInput[v].download(hin[v].data(), hin[v].size(), m.ctx->stream); }
    IntBuffer htar(m.histTarget.size()); FloatBuffer hval(m.histValues.size());
    m.histTarget.download(htar.data(), htar.size(), m.ctx->stream);
    m.histValues.download(hval.data(), hval.size(), m.ctx->stream);
    m.ctx->sync();
    for (int t = 0; t < m.capacity; t++) { // logical order, like CircleBuffer::operator[]
        const size_t s = (size_t)m.slot(t);
        for (int v = 0; v < n; v++) {
            const size_t nv = (size_t)visibleLayerDescs[v].size.x * visibleLayerDescs[v].size.y;
            IntBuffer b(hin[v].begin() + s * nv, hin[v].begin() + (s + 1) * nv);
            writeBufferToStream(os, &b);
        }
        IntBuffer bt(htar.begin() + s * ncols, htar.begin() + (s + 1) * ncols);
        FloatBuffer bv(hval.begin() + s * ncols, hval.begin() + (s + 1) * ncols);
        writeBufferToStream(os, &bt);
        writeBufferToStream(os, &bv);
        os.write(reinterpret_cast<const char *>(&m.rewards[s]), sizeof(float));
    }
}

void Actor::readFromStream(std::istream &is, ComputeSystem &cs) {
    is.read(reinterpret_cast<char *>(&hiddenSize), sizeof(Int3));
    hiddenSize.pad = 0;
    is.read(reinterpret_cast<char *>(&alpha), sizeof(float));
    is.read(reinterpret_cast<char *>(&beta), sizeof(float));
    is.read(reinterpret_cast<char *>(&gamma), sizeof(float));
    is.read(reinterpret_cast<char *>(&minSteps), sizeof(int));
    is.read(reinterpret_cast<char *>(&historyIters), sizeof(int));
    IntBuffer hc; FloatBuffer hv;
    readBufferFromStream(is, &hc);
    readBufferFromStream(is, &hv);
    int n = 0;
    is.read(reinterpret_cast<char *>(&n), sizeof(int));
    visibleLayerDescs.resize(n);
    std::vector<SparseMatrix> mv(n), ma(n);
    for (int v = 0; v < n; v++) {
        is.read(reinterpret_cast<char *>(&visibleLayerDescs[v]), sizeof(VisibleLayerDesc));
        visibleLayerDescs[v].size.pad = 0;
        readSMFromStream(is, mv[v]);
        readSMFromStream(is, ma[v]);
    }
    int historySize = 0, capacity = 0, start = 0;
    is.read(reinterpret_cast<char *>(&historySize), sizeof(int));
    is.read(reinterpret_cast<char *>(&capacity), sizeof(int));
    is.read(reinterpret_cast<char *>(&start), sizeof(int));
    visibleLayers.clear(); visibleLayers.resize(n);
    delete impl; impl = new Impl();
    impl->allocate(cs, hiddenSize, capacity, visibleLayerDescs);
    Impl &m = *impl;
    m.historySize = historySize; m.start = start;
    for (int v = 0; v < n; v++) {
        m.vw.sets[v].uploadCSR(*m.ctx, 0, mv[v].nonZeroValues.data());
        m.aw.sets[v].uploadCSR(*m.ctx, 0, ma[v].nonZeroValues.data());
    }
    const size_t ncols = (size_t)hiddenSize.x * hiddenSize.y;
    for (int t = 0; t < capacity; t++) {
        const size_t s = (size_t)m.slot(t);
        IntBuffer b; FloatBuffer f;
        for (int v = 0; v < n; v++) {
            readBufferFromStream(is, &b);
            m.histInput[v].upload(b.data(), b.size(), m.ctx->stream, s * b.size());
            m.ctx->sync();
        }
        readBufferFromStream(is, &b);
        m.histTarget.upload(b.data(), b.size(), m.ctx->stream, s * ncols);
        readBufferFromStream(is, &f);
        m.histValues.upload(f.data(), f.size(), m.ctx->stream, s * ncols);
        m.ctx->sync();
        is.read(reinterpret_cast<char *>(&m.rewards[s]), sizeof(float));
    }
    m.hiddenCs.upload(hc.data(), hc.size(), m.ctx->stream);
    m.values.upload(hv.data(), hv.size(), m.ctx->stream);
    m.ctx->sync();
}

void Actor::readFromStream(std::istream &is) {
    ComputeSystem cs;
    if (impl) cs.device = impl->ctx->device;
    readFromStream(is, cs);
}

const int *Actor::hiddenCsDevice() const { return impl->hiddenCs.ptr(); }

void Actor::writeDeviceState(std::ostream &os) const {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    const float hp[3] = {alpha, beta, gamma};
    const int hi[5] = {minSteps, historyIters, m.capacity, m.start, m.historySize};
    os.write(reinterpret_cast<const char *>(hp), sizeof(hp));
    os.write(reinterpret_cast<const char *>(hi), sizeof(hi));
    os.write(reinterpret_cast<const char *>(m.rewards.data()), (std::streamsize)(m.rewards.size() * sizeof(float)));
    writeDev(os, ctx, m.acts); writeDev(os, ctx, m.values); writeDev(os, ctx, m.hiddenCs);
    for (const auto &b : m.histInput) writeDev(os, ctx, b);
    writeDev(os, ctx, m.histTarget); writeDev(os, ctx, m.histValues);
    for (const WeightSet &w : m.vw.sets) writeDev(os, ctx, w.wf);
    for (const WeightSet &w : m.aw.sets) writeDev(os, ctx, w.wf);
}

void Actor::readDeviceState(std::istream &is) {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    float hp[3];
    int hi[5];
    is.read(reinterpret_cast<char *>(hp), sizeof(hp));
    is.read(reinterpret_cast<char *>(hi), sizeof(hi));
    if (!is || hi[2] != m.capacity) throw std::runtime_error("ogmaneo-b200: checkpoint Actor history capacity mismatch");
    alpha = hp[0]; beta = hp[1]; gamma = hp[2];
    minSteps = hi[0]; historyIters = hi[1]; m.start = hi[3]; m.historySize = hi[4];
    is.read(reinterpret_cast<char *>(m.rewards.data()), (std::streamsize)(m.rewards.size() * sizeof(float)));
    readDev(is, ctx, m.acts); readDev(is, ctx, m.values); readDev(is, ctx, m.hiddenCs);
    for (auto &b : m.histInput) readDev(is, ctx, b);
    readDev(is, ctx, m.histTarget); readDev(is, ctx, m.histValues);
    for (WeightSet &w : m.vw.sets) readDev(is, ctx, w.wf);
    for (WeightSet &w : m.aw.sets) readDev(is, ctx, w.wf);
    m.csDirty = m.valDirty = true;
    std::fill(m.wDirty.begin(), m.wDirty.end(), 1);
}

} // namespace ogmaneo

// csrc/host/Hierarchy.cpp — ogmaneo::Hierarchy: device-resident histories + the tick schedule that sequences the
// layer kernels on one stream (public interface: include/ogmaneo/Hierarchy.h).
#include "LayerImpl.h"

#include <algorithm>
#include <cstdlib>
#include <cassert>
#include <map>

namespace ogmaneo {
using namespace detail;

namespace {
// CircleBuffer<IntBuffer> (Helpers.h:85-139) in HBM: T slots of n ints, logical index 0 = most recent.
struct Ring {
    DevBuf<int> buf;
    int T = 0, start = 0;
    size_t n = 0;
    void alloc(int T_, size_t n_, cudaStream_t s) { T = T_; n = n_; start = 0; buf.allocZero((size_t)T * n, s); }
    void pushFront() { if (--start < 0) start += T; }
    int *slot(int logical) const { return buf.ptr() + (size_t)((start + logical) % T) * n; }
    void cloneFrom(const Ring &o, cudaStream_t s) { T = o.T; n = o.n; start = o.start; buf.cloneFrom(o.buf, s); }
};
} // namespace

// One captured step: the kernels Hierarchy::step queues are a pure function of the tick pattern, the ring positions and
// the ping-pong parities, so each distinct combination is captured once (cudaStreamBeginCapture) and replayed with a
// single cudaGraphLaunch afterwards. Small hierarchies are otherwise bound by the host's launch calls (8-16 per step).
struct StepGraph {
    int seen = 0;
    cudaGraphExec_t exec = nullptr;
};

// The phases of one step shape, resident in device memory for the single-launch kernel (same key as the graphs).
struct SmallStep {
    DevBuf<ogn_small_phase> phases;
    int nphases = 0;
    int blocks = 1; // blocks per member of this step's launch
};

struct Hierarchy::Impl {
    std::shared_ptr<DeviceContext> ctx;
    int batch = 1;
    std::map<std::vector<int>, StepGraph> graphs;
    // single-launch small-hierarchy step: eligibility (-1 unknown, 0 no, 1 yes), launch shape, cached phase arrays
    int smallOk = -1, smallBlocks = 1, smallThreads = 256, smallWide = 1;
    std::map<std::vector<int>, SmallStep> smallSteps;
    DevBuf<unsigned> smallCounters;
    void dropGraphs() {
        for (auto &g : graphs)
            if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
        graphs.clear();
    }
    // pipelined prediction readback (queuePredictionReadback): two pinned slots on a copy stream of their own
    struct Readback { PinnedBuf<int> buf; cudaEvent_t done = nullptr; size_t n = 0; bool pending = false; };
    Readback rb[2];
    int rbNext = 0;
    cudaStream_t copyStream = nullptr;
    cudaEvent_t rbSrcReady = nullptr;
    std::vector<std::vector<Ring>> histories; // [layer][input]
    std::vector<std::unique_ptr<PinnedBuf<int>>> inPinned;
    cudaEvent_t inFree = nullptr; // pinned input staging may be overwritten after this event
    ~Impl() {
        dropGraphs();
        if (inFree) cudaEventDestroy(inFree);
        for (auto &r : rb)
            if (r.done) cudaEventDestroy(r.done);
        if (rbSrcReady) cudaEventDestroy(rbSrcReady);
        if (copyStream) cudaStreamDestroy(copyStream);
    }
};

Hierarchy::Hierarchy() : impl(nullptr) {}
Hierarchy::~Hierarchy() { delete impl; }
Hierarchy::Hierarchy(const Hierarchy &o) : impl(nullptr) { *this = o; }

// Hierarchy.cpp:149-190 (deep copy)
const Hierarchy &Hierarchy::operator=(const Hierarchy &o) {
    if (this == &o) return *this;
    scLayers = o.scLayers;
    updates = o.updates; ticks = o.ticks; ticksPerUpdate = o.ticksPerUpdate; inputSizes = o.inputSizes;
    pLayers.clear(); pLayers.resize(o.pLayers.size());
    for (size_t l = 0; l < o.pLayers.size(); l++) {
        pLayers[l].resize(o.pLayers[l].size());
        for (size_t v = 0; v < o.pLayers[l].size(); v++)
            if (o.pLayers[l][v]) pLayers[l][v].reset(new Predictor(*o.pLayers[l][v]));
    }
    aLayers.clear(); aLayers.resize(o.aLayers.size());
    for (size_t v = 0; v < o.aLayers.size(); v++)
        if (o.aLayers[v]) aLayers[v].reset(new Actor(*o.aLayers[v]));
    delete impl; impl = nullptr;
    if (o.impl) {
        impl = new Impl();
        impl->ctx = o.impl->ctx; impl->batch = o.impl->batch;
        cudaCheck(cudaSetDevice(impl->ctx->device), "cudaSetDevice");
        impl->histories.resize(o.impl->histories.size());
        for (size_t l = 0; l < impl->histories.size(); l++) {
            impl->histories[l].resize(o.impl->histories[l].size());
            for (size_t i = 0; i < impl->histories[l].size(); i++) impl->histories[l][i].cloneFrom(o.impl->histories[l][i], impl->ctx->stream);
        }
        for (size_t i = 0; i < o.impl->inPinned.size(); i++) {
            impl->inPinned.emplace_back(new PinnedBuf<int>());
            impl->inPinned.back()->alloc(o.impl->inPinned[i]->size());
        }
        cudaCheck(cudaEventCreateWithFlags(&impl->inFree, cudaEventDisableTiming), "cudaEventCreate");
        cudaCheck(cudaEventRecord(impl->inFree, impl->ctx->stream), "cudaEventRecord");
        impl->ctx->sync();
    }
    return *this;
}

// Hierarchy.cpp:16-147. The order of the initRandom calls (layer-0 predictors / actors by input index, SparseCoder 0,
// layer-1 predictors, SparseCoder 1, ...) is the order in which cs.rng is consumed and must not change.
void Hierarchy::initImpl(ComputeSystem &cs, const std::vector<Int3> &inSizes, const std::vector<InputType> &inTypes,
                         const std::vector<LayerDesc> &lds, int batch, int member) {
    const int L = (int)lds.size(), NI = (int)inSizes.size();
    const bool first = member == 0;
    if (first) {
        for (int i = 0; i < NI; i++)
            if (inTypes[i] == InputType::action) {
                if (batch > 1) throw std::runtime_error("ogmaneo-b200: ensembles do not support action inputs");
                cs.replayRng = true; // Actor sampling depends on every earlier draw from cs.rng
            }
        delete impl; impl = new Impl();
        impl->ctx = cs.sharedContext(); impl->batch = batch;
        scLayers.clear(); scLayers.resize(L);
        pLayers.clear(); pLayers.resize(L);
        aLayers.clear();
        ticks.assign(L, 0); updates.assign(L, 0); ticksPerUpdate.resize(L);
        inputSizes = inSizes;
        for (auto &s : inputSizes) s.pad = 0;
        for (int l = 0; l < L; l++) ticksPerUpdate[l] = l == 0 ? 1 : lds[l].ticksPerUpdate;
        impl->histories.clear(); impl->histories.resize(L);
        impl->inPinned.clear();
        cudaCheck(cudaEventCreateWithFlags(&impl->inFree, cudaEventDisableTiming), "cudaEventCreate");
        cudaCheck(cudaEventRecord(impl->inFree, impl->ctx->stream), "cudaEventRecord");
    }
    cudaStream_t st = impl->ctx->stream;
    for (int l = 0; l < L; l++) {
        std::vector<SparseCoder::VisibleLayerDesc> scV;
        std::vector<Predictor::VisibleLayerDesc> pV(1);
        pV[0].size = lds[l].hiddenSize; pV[0].radius = lds[l].pRadius;
        if (l < L - 1) pV.push_back(pV[0]);
        if (l == 0) {
            const int T = lds[l].temporalHorizon;
            scV.resize((size_t)NI * T);
            for (int i = 0; i < NI; i++)
                for (int t = 0; t < T; t++) { scV[t + T * i].size = inSizes[i]; scV[t + T * i].radius = lds[l].ffRadius; }
            std::vector<Actor::VisibleLayerDesc> aV(1);
            aV[0].size = lds[l].hiddenSize; aV[0].radius = lds[l].aRadius;
            if (l < L - 1) aV.push_back(aV[0]);
            if (first) {
                impl->histories[l].resize(NI);
                for (int i = 0; i < NI; i++) {
                    const size_t n = (size_t)inSizes[i].x * inSizes[i].y * batch;
                    impl->histories[l][i].alloc(T, n, st);
                    impl->inPinned.emplace_back(new PinnedBuf<int>());
                    impl->inPinned.back()->alloc(n);
                }
                pLayers[l].resize(NI); aLayers.resize(NI);
            }
            for (int p = 0; p < NI; p++) {
                if (inTypes[p] == InputType::prediction) {
                    if (first) pLayers[l][p].reset(new Predictor());
                    pLayers[l][p]->initRandomMember(cs, inSizes[p], pV, batch, member);
                } else if (inTypes[p] == InputType::action) {
                    aLayers[p].reset(new Actor());
                    aLayers[p]->initRandom(cs, inSizes[p], lds[l].historyCapacity, aV);
                }
            }
        } else {
            const int T = lds[l].temporalHorizon;
            scV.resize(T);
            for (int t = 0; t < T; t++) { scV[t].size = lds[l - 1].hiddenSize; scV[t].radius = lds[l].ffRadius; }
            if (first) {
                impl->histories[l].resize(1);
                impl->histories[l][0].alloc(T, (size_t)lds[l - 1].hiddenSize.x * lds[l - 1].hiddenSize.y * batch, st);
                pLayers[l].resize(lds[l].ticksPerUpdate);
            }
            for (size_t p = 0; p < pLayers[l].size(); p++) {
                if (first) pLayers[l][p].reset(new Predictor());
                pLayers[l][p]->initRandomMember(cs, lds[l - 1].hiddenSize, pV, batch, member);
            }
        }
        scLayers[l].initRandomMember(cs, lds[l].hiddenSize, scV, batch, member);
    }
    impl->ctx->sync();
}

void Hierarchy::initRandom(ComputeSystem &cs, const std::vector<Int3> &inSizes, const std::vector<InputType> &inTypes,
                           const std::vector<LayerDesc> &lds) {
    initImpl(cs, inSizes, inTypes, lds, 1, 0);
}

void Hierarchy::initRandomEnsemble(ComputeSystem &cs, const std::vector<unsigned> &seeds, const std::vector<Int3> &inSizes,
                                   const std::vector<InputType> &inTypes, const std::vector<LayerDesc> &lds) {
    if (cs.shardWorld > 1) throw std::runtime_error("ogmaneo-b200: shard an ensemble by giving each device its own members");
    for (size_t k = 0; k < seeds.size(); k++) {
        cs.rng.seed(seeds[k]);
        cs.deviceInitSeed = seeds[k];
        cs.deviceInitMatrix = 0;
        initImpl(cs, inSizes, inTypes, lds, (int)seeds.size(), (int)k);
    }
}

int Hierarchy::getEnsembleSize() const { return impl ? impl->batch : 0; }

// Sharded top layer: the kernels queued behind the exchange of its hidden CSDR (its own learn pass, the Predictors of
// the layer) read only rows near this device's slab, so the exchange waits for the ranks owning those rows instead of for
// everybody. Requires a later exchange of the same step that does wait for everybody (the deferred exchange of the
// layer-0 predictions), which is what makes the other slabs complete for getters. 0 = wait for everybody.
unsigned Hierarchy::haloWaitMask(ComputeSystem &cs, int l) {
    if (cs.shardWorld <= 1 || l != (int)scLayers.size() - 1) return 0;
    bool finalExchange = false;
    for (auto &p : pLayers[0]) finalExchange = finalExchange || p != nullptr;
    if (!finalExchange) return 0;
    int lo = 0, hi = 0;
    scLayers[l].learnRows(lo, hi);
    for (auto &p : pLayers[l]) {
        if (!p) continue;
        int a = 0, b = 0;
        p->visibleRows(0, a, b); // visible layer 0 of a Predictor is this layer's hidden CSDR
        lo = std::min(lo, a); hi = std::max(hi, b);
    }
    const int Hx = scLayers[l].getHiddenSize().x, rows = Hx / cs.shardWorld;
    if (rows <= 0 || hi <= lo) return 0;
    unsigned mask = 0;
    for (int r = 0; r < cs.shardWorld && r < 32; r++)
        if (r != cs.shardRank && r * rows < hi && (r + 1) * rows > lo) mask |= 1u << r;
    return mask ? mask : 0;
}

// Whether this hierarchy steps through the single-launch kernel, and with which launch shape (decided once).
bool Hierarchy::smallEligible(ComputeSystem &cs) {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    if (m.smallOk >= 0) return m.smallOk == 1;
    m.smallOk = 0;
    // Stream choice, once per hierarchy (OGN_SIDE_STREAM overrides): sharded layers always use the side stream (§5); on one
    // device it pays when the layers are small, where a step is a chain of short latency-bound kernels and SparseCoder
    // learn can run next to the upper layers' forward passes and the Predictors (measured: C1 +18 %, C2 +23 %, C3 +38 %,
    // C5 +6 %; captured into the step graphs with its fork/join edges). Large layers keep one stream: their kernels are
    // bandwidth-bound either way and the hoisted Predictor learn would cost a second pass over the launch overheads.
    if (!ctx.sideForced && cs.shardWorld == 1 && !ctx.small) {
        int cols = 0;
        for (auto &sc : scLayers) cols = std::max(cols, sc.getHiddenSize().x * sc.getHiddenSize().y);
        for (auto &sz : inputSizes) cols = std::max(cols, sz.x * sz.y);
        ctx.useSide = cols <= 4096;
    }
    if (!ctx.small || ctx.useSide || cs.shardWorld != 1 || (int)inputSizes.size() > OGN_SMALL_MAX_INPUTS) return false;
    for (auto &a : aLayers)
        if (a) return false; // the Actor takes fresh host-drawn numbers every step
    int units = 0, lists = 0;
    for (size_t l = 0; l < scLayers.size(); l++) {
        const int u = scLayers[l].smallUnits();
        if (u < 0) return false;
        units = std::max(units, u);
        lists += scLayers[l].getNumVisibleLayers(); // upper bound on the layer's runs
        for (auto &p : pLayers[l]) {
            if (!p) continue;
            const int pu = p->smallUnits();
            if (pu < 0) return false;
            units = std::max(units, pu);
        }
    }
    if (units <= 0 || lists > OGN_SMALL_MAX_LISTS) return false;
    if (m.batch > 1) { // ensemble: one block per member, one octet per pack; members never wait for each other
        if (units > 2048) return false;
        m.smallWide = 0; m.smallBlocks = 1;
        m.smallThreads = std::min(512, std::max(64, ((units * 8 + 31) / 32) * 32));
    } else {           // one hierarchy: one warp per pack over as many co-resident blocks as the largest phase can use
        if (units > ctx.smallMaxUnits) return false;
        m.smallWide = 1;
        if (units <= 64) { m.smallBlocks = 1; m.smallThreads = std::min(512, 32 * units); } // one block: __syncthreads is the only barrier
        else {
            m.smallThreads = 256;
            const int cap = ogn_hier_small_max_blocks(m.smallThreads);
            if (cap < 1) return false;
            m.smallBlocks = std::max(1, std::min((units + 7) / 8, cap));
        }
    }
    m.smallCounters.allocZero((size_t)m.batch * OGN_SMALL_COUNTERS, ctx.stream);
    m.smallOk = 1;
    return true;
}

// Hierarchy.cpp:192-285
void Hierarchy::stepImpl(ComputeSystem &cs, const int *const *hostIn, const std::vector<const int *> *devIn,
                         bool learnEnabled, float reward, bool mimic) {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    const int L = (int)scLayers.size(), NI = (int)inputSizes.size();
    if (devIn && devIn->size() != (size_t)NI) throw std::invalid_argument("ogmaneo-b200: wrong number of input CSDRs");

    ctx.beginStep();
    ticks[0] = 0;
    // a queued prediction readback still reads the buffer this step's Predictor will overwrite
    for (auto &r : m.rb)
        if (r.pending) cudaCheck(cudaStreamWaitEvent(ctx.stream, r.done, 0), "cudaStreamWaitEvent");
    const bool small = smallEligible(cs);
    ogn_small_inputs smallIn;
    std::memset(&smallIn, 0, sizeof(smallIn));
    // K1: inputs -> front of the layer-0 history rings
    if (hostIn) cudaCheck(cudaEventSynchronize(m.inFree), "cudaEventSynchronize");
    for (int i = 0; i < NI; i++) {
        Ring &r = m.histories[0][i];
        r.pushFront();
        consumeKernel1(cs, (int)(r.n / m.batch));
        if (small && !hostIn) { // device-resident inputs: the single-launch kernel copies them itself
            smallIn.src[i] = (*devIn)[i]; smallIn.dst[i] = r.slot(0); smallIn.count[i] = (int)(r.n / m.batch); smallIn.n = i + 1;
            continue;
        }
        if (hostIn) {
            // a shard only reads the input rows its slab's fields and its Predictor rows touch: copy just those
            size_t a = 0, b = r.n;
            if (cs.shardWorld > 1 && m.batch == 1) {
                int lo = 0, hi = 0;
                scLayers[0].visibleRows(i * (int)r.T, lo, hi);
                if ((size_t)i < pLayers[0].size() && pLayers[0][i]) {
                    int pl = 0, ph = 0;
                    pLayers[0][i]->hiddenRows(pl, ph);
                    lo = std::min(lo, pl); hi = std::max(hi, ph);
                }
                a = (size_t)lo * inputSizes[i].y; b = (size_t)hi * inputSizes[i].y;
            }
            std::memcpy(m.inPinned[i]->ptr() + a, hostIn[i] + a, (b - a) * sizeof(int));
            cudaCheck(cudaMemcpyAsync(r.slot(0) + a, m.inPinned[i]->ptr() + a, (b - a) * sizeof(int), cudaMemcpyHostToDevice, ctx.stream),
                      "input H2D");
        } else
            cudaCheck(cudaMemcpyAsync(r.slot(0), (*devIn)[i], r.n * sizeof(int), cudaMemcpyDeviceToDevice, ctx.stream), "input D2D");
    }
    if (hostIn) cudaCheck(cudaEventRecord(m.inFree, ctx.stream), "cudaEventRecord");

    auto body = [&]() {
    // Sharded layers: the learn pass of the layer-0 Predictors needs nothing this step produces (its target is the
    // input just copied, its gather rows are last step's inputs), so it is hoisted to the side stream and runs while the
    // main stream does the SparseCoder forward pass and its slab exchange; the activate pass follows on the main stream.
    static const bool hoistAllowed = [] { const char *e = std::getenv("OGN_PRED_HOIST"); return !e || std::atoi(e) != 0; }();
    const bool hoistPredLearn = hoistAllowed && ctx.useSide && learnEnabled && !cs.replayRng;
    bool predLearnPending = false;
    if (hoistPredLearn) {
        bool any = false;
        for (size_t p = 0; p < pLayers[0].size(); p++) {
            if (!pLayers[0][p]) continue;
            if (!any) { ctx.forkSide(); any = true; }
            pLayers[0][p]->stepDevice(cs, true, m.histories[0][p].slot(0), false, nullptr, false, true);
        }
        if (any && !ctx.dry) cudaCheck(cudaEventRecord(ctx.evPredLearn, ctx.side), "cudaEventRecord");
        predLearnPending = any;
    }

    updates.assign(L, 0);
    std::vector<const int *> ptrs;
    // forward (bottom-up) pass of the SparseCoders
    for (int l = 0; l < L; l++) {
        if (!(l == 0 || ticks[l] >= ticksPerUpdate[l])) continue;
        ticks[l] = 0; updates[l] = 1;
        const size_t NH = m.histories[l].size(), T = (size_t)m.histories[l][0].T;
        ptrs.resize(NH * T);
        for (size_t i = 0; i < NH; i++)
            for (size_t t = 0; t < T; t++) ptrs[t + T * i] = m.histories[l][i].slot((int)t);
        int *copyOut = nullptr;
        if (l < L - 1) { // K5: the new hidden state also becomes the next layer's most recent input
            m.histories[l + 1][0].pushFront();
            copyOut = m.histories[l + 1][0].slot(0);
        }
        if (ctx.collect) ctx.collect->curLayer = l;
        scLayers[l].stepDevice(cs, ptrs.data(), learnEnabled, copyOut, haloWaitMask(cs, l));
        if (l < L - 1) {
            consumeKernel1(cs, scLayers[l].getHiddenSize().x * scLayers[l].getHiddenSize().y);
            ticks[l + 1]++;
        }
    }
    // backward (top-down) pass of the Predictors / Actors
    for (int l = L - 1; l >= 0; l--) {
        if (!updates[l]) continue;
        if (ctx.collect) ctx.collect->curLayer = l;
        const int *fb[2] = {scLayers[l].hiddenCsDevice(), nullptr};
        if (l < L - 1) fb[1] = pLayers[l + 1][ticksPerUpdate[l + 1] - 1 - ticks[l + 1]]->hiddenCsDevice();
        for (size_t p = 0; p < pLayers[l].size(); p++) {
            if (!pLayers[l][p]) continue;
            const int *target = l == 0 ? m.histories[0][p].slot(0) : m.histories[l][0].slot((int)p);
            // layer 0's predictions are the step's output: nothing queued later in this step reads them, so their
            // exchange between shards is deferred to the side stream
            const bool hoisted = l == 0 && hoistPredLearn;
            if (hoisted && predLearnPending) {
                if (!ctx.dry) cudaCheck(cudaStreamWaitEvent(ctx.stream, ctx.evPredLearn, 0), "cudaStreamWaitEvent");
                predLearnPending = false;
            }
            pLayers[l][p]->stepDevice(cs, learnEnabled && !hoisted, target, true, fb, l == 0);
        }
        if (l == 0)
            for (size_t p = 0; p < aLayers.size(); p++)
                if (aLayers[p]) aLayers[p]->stepDevice(cs, fb, m.histories[0][p].slot(0), reward, learnEnabled, mimic);
    }
    // SparseCoder learn ran on the side stream, concurrently with the Predictors above: everything queued on the main
    // stream from here on (the next step, getters, copies) is ordered behind it
    ctx.joinLearn();
    };

    // Small hierarchies: the whole step is ONE launch (ogn_hier_step_small). The phases are recorded once per step shape
    // (same key as the graphs below), ordered into barrier levels and kept in device memory; afterwards a step is the
    // host-side state transitions plus one launch.
    if (small) {
        std::vector<int> key;
        key.push_back(learnEnabled ? 1 : 0);
        key.insert(key.end(), ticks.begin(), ticks.end());
        for (int l = 0; l < L; l++) {
            scLayers[l].stateSignature(key);
            for (auto &r : m.histories[l]) key.push_back(r.start);
            for (auto &pl : pLayers[l])
                if (pl) pl->stateSignature(key);
        }
        if (m.smallSteps.size() >= 4096) m.smallSteps.clear();
        SmallStep &ss = m.smallSteps[key];
        if (ss.nphases == 0) {
            SmallPlan plan;
            ctx.collect = &plan;
            try { body(); } catch (...) { ctx.collect = nullptr; throw; }
            ctx.collect = nullptr;
            if (!plan.ok) throw std::logic_error("ogmaneo-b200: single-launch plan holds a phase without a 128-bit kernel");
            // barrier levels: with layers 0..n-1 ticking, forward(l) is level l, its reconstruction l + 1, its
            // update l + 2 and the Predictors of layer l level n + (n - 1 - l); phases of one level are independent
            int n = 0;
            for (int l = 0; l < L; l++) n += updates[l] ? 1 : 0;
            std::vector<std::pair<int, int>> order; // (level * 4 + priority inside the level, node index)
            for (size_t i = 0; i < plan.nodes.size(); i++) {
                const SmallPlan::Node &nd = plan.nodes[i];
                int level = 0, prio = 0;
                switch (nd.ph.kind) {
                case OGN_PHASE_SC_FORWARD: level = nd.layer; prio = 0; break;
                case OGN_PHASE_SC_RECON: level = nd.layer + 1; prio = 2; break;
                case OGN_PHASE_SC_UPDATE: level = nd.layer + 2; prio = 3; break;
                default: level = n + (n - 1 - nd.layer); prio = 1; break;
                }
                order.emplace_back(level * 4 + prio, (int)i);
            }
            std::stable_sort(order.begin(), order.end());
            std::vector<ogn_small_phase> ph(order.size());
            // unit slots: inside a level every phase starts where the previous one's units end (update phases, whose unit
            // count is data-dependent, come last in their level), and the launch gets enough blocks for its widest level
            static const int spread = std::getenv("OGN_SMALL_SPREAD") ? std::atoi(std::getenv("OGN_SMALL_SPREAD")) : 1;
            int shift = 0, widest = 0;
            for (size_t i = 0; i < order.size(); i++) {
                ph[i] = plan.nodes[order[i].second].ph;
                ph[i].barrierAfter = (i + 1 < order.size() && order[i + 1].first / 4 != order[i].first / 4) ? 1 : 0;
                ph[i].unitShift = (spread && m.smallWide && m.smallBlocks > 1) ? shift : 0;
                const int units = ph[i].kind == OGN_PHASE_SC_RECON ? ph[i].npacks * ph[i].u.learn.nvl
                                  : ph[i].kind == OGN_PHASE_SC_UPDATE ? ph[i].npacks : ph[i].npacks;
                shift += units;
                widest = std::max(widest, shift);
                if (ph[i].barrierAfter) shift = 0;
            }
            ss.blocks = m.smallBlocks;
            if (spread && m.smallWide && m.smallBlocks > 1) {
                const int cap = ogn_hier_small_max_blocks(m.smallThreads);
                ss.blocks = std::max(m.smallBlocks, std::min((widest + 7) / 8, cap));
            }
            ss.phases.alloc(ph.size());
            cudaCheck(cudaMemcpyAsync(ss.phases.ptr(), ph.data(), ph.size() * sizeof(ogn_small_phase), cudaMemcpyHostToDevice, ctx.stream),
                      "phase upload");
            cudaCheck(cudaStreamSynchronize(ctx.stream), "cudaStreamSynchronize"); // `ph` is pageable and goes out of scope
            ss.nphases = (int)ph.size();
        } else {
            ctx.dry = true; ctx.dryUncounted = true;
            try { body(); } catch (...) { ctx.dry = false; ctx.dryUncounted = false; throw; }
            ctx.dry = false; ctx.dryUncounted = false;
        }
        LaunchScope ls(ctx, kHierSmall);
        kernelCheck(ogn_hier_step_small(ss.phases.ptr(), ss.nphases, &smallIn, m.batch, ss.blocks, m.smallThreads, m.smallWide,
                                        m.smallCounters.ptr(), ctx.stream), "hier_step_small");
        return;
    }

    // Graph replay (single device, single stream, no Actor: the Actor uploads fresh host-drawn numbers every step; a
    // profiled step is queued normally so that its kernels can be bracketed by events).
    // With the side stream on one device (SparseCoder learn next to the upper layers' forward passes and the Predictors) the
    // capture records the fork/join edges, so the graph runs the independent kernels of a step side by side.
    bool graphOk = ctx.graphs && cs.shardWorld == 1 && !(ctx.profiling && ctx.sampleThis) && m.graphs.size() < 1024;
    for (size_t p = 0; p < aLayers.size() && graphOk; p++)
        if (aLayers[p]) graphOk = false;
    if (!graphOk) { body(); return; }
    std::vector<int> key;
    key.push_back(learnEnabled ? 1 : 0);
    key.insert(key.end(), ticks.begin(), ticks.end());
    for (int l = 0; l < L; l++) {
        scLayers[l].stateSignature(key);
        for (auto &r : m.histories[l]) key.push_back(r.start);
        for (auto &pl : pLayers[l])
            if (pl) pl->stateSignature(key);
    }
    StepGraph &g = m.graphs[key];
    if (g.exec) { // host-side state transitions only, then one launch
        ctx.dry = true;
        try { body(); } catch (...) { ctx.dry = false; throw; }
        ctx.dry = false;
        cudaCheck(cudaGraphLaunch(g.exec, ctx.stream), "cudaGraphLaunch");
    } else if (g.seen >= 1) { // second occurrence: every lazy initialisation behind it, capture
        cudaCheck(cudaStreamBeginCapture(ctx.stream, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
        cudaGraph_t graph = nullptr;
        try { body(); } catch (...) { cudaStreamEndCapture(ctx.stream, &graph); if (graph) cudaGraphDestroy(graph); throw; }
        cudaCheck(cudaStreamEndCapture(ctx.stream, &graph), "cudaStreamEndCapture");
        cudaError_t e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        cudaCheck(e, "cudaGraphInstantiate");
        cudaCheck(cudaGraphLaunch(g.exec, ctx.stream), "cudaGraphLaunch");
    } else {
        g.seen++;
        body();
    }
}

void Hierarchy::step(ComputeSystem &cs, const std::vector<const IntBuffer *> &inputCs, bool learnEnabled, float reward, bool mimic) {
    if (inputCs.size() != inputSizes.size()) throw std::invalid_argument("ogmaneo-b200: wrong number of input CSDRs");
    std::vector<const int *> raw(inputCs.size());
    for (size_t i = 0; i < inputCs.size(); i++) {
        if (inputCs[i]->size() != impl->histories[0][i].n) throw std::invalid_argument("ogmaneo-b200: input CSDR has the wrong number of columns");
        raw[i] = inputCs[i]->data();
    }
    stepImpl(cs, raw.data(), nullptr, learnEnabled, reward, mimic);
}

void Hierarchy::stepHost(ComputeSystem &cs, const int *const *inputCs, bool learnEnabled, float reward, bool mimic) {
    stepImpl(cs, inputCs, nullptr, learnEnabled, reward, mimic);
}

int Hierarchy::copyPredictionCs(int i, int *out) const {
    if (aLayers[i] != nullptr) {
        const IntBuffer &b = aLayers[i]->getHiddenCs();
        if (out) std::memcpy(out, b.data(), b.size() * sizeof(int));
        return (int)b.size();
    }
    return pLayers.front()[i]->copyHiddenCs(out);
}

void Hierarchy::stepDevice(ComputeSystem &cs, const std::vector<const int *> &inputCsDevice, bool learnEnabled, float reward, bool mimic) {
    stepImpl(cs, nullptr, &inputCsDevice, learnEnabled, reward, mimic);
}

const int *Hierarchy::getPredictionCsDevice(int i) const {
    // sharded: the exchange of the step's output predictions was deferred to the side stream; a caller that orders its
    // work against the main stream (ogn_stream) must find the peers' slabs in place, so join it here
    if (impl) impl->ctx->joinExchange();
    if (aLayers[i] != nullptr) return aLayers[i]->hiddenCsDevice();
    return pLayers.front()[i]->hiddenCsDevice();
}

int Hierarchy::queuePredictionReadback(int i, int x0, int x1) {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    if (aLayers[i] != nullptr || !pLayers.front()[i]) throw std::invalid_argument("ogmaneo-b200: readback needs a prediction input");
    Predictor &p = *pLayers.front()[i];
    const int X = p.getHiddenSize().x, Y = p.getHiddenSize().y;
    size_t off = 0, n = (size_t)X * Y * m.batch;
    bool own = false;
    if (x1 > x0) {
        if (m.batch > 1 || x0 < 0 || x1 > X) throw std::invalid_argument("ogmaneo-b200: readback row range");
        off = (size_t)x0 * Y; n = (size_t)(x1 - x0) * Y;
        int a = 0, b = 0;
        p.hiddenRows(a, b);
        own = x0 >= a && x1 <= b;
    }
    if (!own) ctx.joinExchange(); // rows of other ranks arrive with the deferred exchange
    if (!m.copyStream) {
        cudaCheck(cudaStreamCreateWithFlags(&m.copyStream, cudaStreamNonBlocking), "cudaStreamCreate");
        cudaCheck(cudaEventCreateWithFlags(&m.rbSrcReady, cudaEventDisableTiming), "cudaEventCreate");
    }
    const int t = m.rbNext;
    Impl::Readback &r = m.rb[t];
    if (r.pending) throw std::runtime_error("ogmaneo-b200: two prediction readbacks are already outstanding");
    if (!r.done) cudaCheck(cudaEventCreateWithFlags(&r.done, cudaEventDisableTiming), "cudaEventCreate");
    if (r.buf.size() < n) r.buf.alloc(n);
    cudaCheck(cudaEventRecord(m.rbSrcReady, ctx.stream), "cudaEventRecord");
    cudaCheck(cudaStreamWaitEvent(m.copyStream, m.rbSrcReady, 0), "cudaStreamWaitEvent");
    cudaCheck(cudaMemcpyAsync(r.buf.ptr(), p.hiddenCsDevice() + off, n * sizeof(int), cudaMemcpyDeviceToHost, m.copyStream), "prediction D2H");
    cudaCheck(cudaEventRecord(r.done, m.copyStream), "cudaEventRecord");
    r.n = n; r.pending = true;
    m.rbNext = 1 - t;
    return t;
}

int Hierarchy::finishPredictionReadback(int ticket, int *out) {
    Impl &m = *impl;
    if (ticket < 0 || ticket > 1 || !m.rb[ticket].pending) throw std::invalid_argument("ogmaneo-b200: no such readback ticket");
    Impl::Readback &r = m.rb[ticket];
    cudaCheck(cudaEventSynchronize(r.done), "cudaEventSynchronize");
    m.ctx->peerCheck();
    if (out) std::memcpy(out, r.buf.ptr(), r.n * sizeof(int));
    r.pending = false;
    return (int)r.n;
}

// --- state ---------------------------------------------------------------------------------------------------------
namespace {
void ringToHost(const Ring &r, DeviceContext &ctx, CircleBuffer<IntBuffer> &out) {
    out.resize(r.T);
    out.start = r.start;
    for (int t = 0; t < r.T; t++) { // physical order, so that `start` means the same thing
        out.data[t].resize(r.n);
        cudaCheck(cudaMemcpyAsync(out.data[t].data(), r.buf.ptr() + (size_t)t * r.n, r.n * sizeof(int), cudaMemcpyDeviceToHost, ctx.stream), "history D2H");
    }
    ctx.sync();
}
void ringFromHost(Ring &r, DeviceContext &ctx, const CircleBuffer<IntBuffer> &in) {
    if (in.size() != r.T) throw std::invalid_argument("ogmaneo-b200: history length mismatch");
    r.start = in.start;
    for (int t = 0; t < r.T; t++) {
        if (in.data[t].size() != r.n) throw std::invalid_argument("ogmaneo-b200: history CSDR size mismatch");
        cudaCheck(cudaMemcpyAsync(r.buf.ptr() + (size_t)t * r.n, in.data[t].data(), r.n * sizeof(int), cudaMemcpyHostToDevice, ctx.stream), "history H2D");
    }
    ctx.sync();
}
} // namespace

// Hierarchy.cpp:434-466
void Hierarchy::getState(State &state) const {
    const int L = (int)scLayers.size();
    cudaCheck(cudaSetDevice(impl->ctx->device), "cudaSetDevice");
    state.hiddenCs.resize(L); state.hiddenCsPrev.resize(L); state.predHiddenCs.resize(L); state.predInputCsPrev.resize(L);
    state.histories.resize(L);
    for (int l = 0; l < L; l++) {
        state.hiddenCs[l] = scLayers[l].getHiddenCs();
        state.hiddenCsPrev[l] = scLayers[l].getHiddenCsPrev();
        state.predHiddenCs[l].resize(pLayers[l].size());
        state.predInputCsPrev[l].resize(pLayers[l].size());
        for (size_t j = 0; j < pLayers[l].size(); j++) {
            if (!pLayers[l][j]) continue;
            state.predHiddenCs[l][j] = pLayers[l][j]->getHiddenCs();
            state.predInputCsPrev[l][j].resize(pLayers[l][j]->getNumVisibleLayers());
            for (int v = 0; v < pLayers[l][j]->getNumVisibleLayers(); v++) state.predInputCsPrev[l][j][v] = pLayers[l][j]->getInputCsPrev(v);
        }
        state.histories[l].resize(impl->histories[l].size());
        for (size_t i = 0; i < impl->histories[l].size(); i++) ringToHost(impl->histories[l][i], *impl->ctx, state.histories[l][i]);
    }
    state.ticks = ticks;
    state.updates = updates;
}

// Hierarchy.cpp:468-493
void Hierarchy::setState(const State &state) {
    const int L = (int)scLayers.size();
    cudaCheck(cudaSetDevice(impl->ctx->device), "cudaSetDevice");
    for (int l = 0; l < L; l++) {
        scLayers[l].setHiddenState(state.hiddenCs[l], state.hiddenCsPrev[l]);
        for (size_t i = 0; i < impl->histories[l].size(); i++) ringFromHost(impl->histories[l][i], *impl->ctx, state.histories[l][i]);
        for (size_t j = 0; j < pLayers[l].size(); j++)
            if (pLayers[l][j]) pLayers[l][j]->setState(state.predHiddenCs[l][j], state.predInputCsPrev[l][j]);
    }
    ticks = state.ticks;
    updates = state.updates;
}

// --- streams: Hierarchy.cpp:287-432, layout in SURVEY.md Appendix B --------------------------------------------------
void Hierarchy::writeToStream(std::ostream &os) const {
    if (impl->batch > 1) throw std::runtime_error("ogmaneo-b200: writeToStream needs a single hierarchy (not an ensemble)");
    cudaCheck(cudaSetDevice(impl->ctx->device), "cudaSetDevice");
    const int L = (int)scLayers.size(), NI = (int)inputSizes.size();
    os.write(reinterpret_cast<const char *>(&L), sizeof(int));
    os.write(reinterpret_cast<const char *>(&NI), sizeof(int));
    os.write(reinterpret_cast<const char *>(inputSizes.data()), NI * sizeof(Int3));
    os.write(reinterpret_cast<const char *>(updates.data()), L * sizeof(char));
    os.write(reinterpret_cast<const char *>(ticks.data()), L * sizeof(int));
    os.write(reinterpret_cast<const char *>(ticksPerUpdate.data()), L * sizeof(int));
    for (int l = 0; l < L; l++) {
        int nh = (int)impl->histories[l].size();
        os.write(reinterpret_cast<const char *>(&nh), sizeof(int));
        for (int i = 0; i < nh; i++) {
            CircleBuffer<IntBuffer> h;
            ringToHost(impl->histories[l][i], *impl->ctx, h);
            int size = h.size(), start = h.start;
            os.write(reinterpret_cast<const char *>(&size), sizeof(int));
            os.write(reinterpret_cast<const char *>(&start), sizeof(int));
            for (int t = 0; t < size; t++) writeBufferToStream(os, &h[t]); // logical order
        }
        scLayers[l].writeToStream(os);
        for (size_t v = 0; v < pLayers[l].size(); v++) {
            char exists = pLayers[l][v] != nullptr;
            os.write(&exists, 1);
            if (exists) pLayers[l][v]->writeToStream(os);
        }
    }
    for (size_t v = 0; v < aLayers.size(); v++) {
        char exists = aLayers[v] != nullptr;
        os.write(&exists, 1);
        if (exists) aLayers[v]->writeToStream(os);
    }
}

void Hierarchy::readFromStream(std::istream &is, ComputeSystem &cs) {
    int L = 0, NI = 0;
    is.read(reinterpret_cast<char *>(&L), sizeof(int));
    is.read(reinterpret_cast<char *>(&NI), sizeof(int));
    inputSizes.resize(NI);
    is.read(reinterpret_cast<char *>(inputSizes.data()), NI * sizeof(Int3));
    for (auto &s : inputSizes) s.pad = 0;
    scLayers.clear(); scLayers.resize(L);
    pLayers.clear(); pLayers.resize(L);
    ticks.resize(L); ticksPerUpdate.resize(L); updates.resize(L);
    is.read(reinterpret_cast<char *>(updates.data()), L * sizeof(char));
    is.read(reinterpret_cast<char *>(ticks.data()), L * sizeof(int));
    is.read(reinterpret_cast<char *>(ticksPerUpdate.data()), L * sizeof(int));
    delete impl; impl = new Impl();
    impl->ctx = cs.sharedContext(); impl->batch = 1;
    cudaCheck(cudaEventCreateWithFlags(&impl->inFree, cudaEventDisableTiming), "cudaEventCreate");
    cudaCheck(cudaEventRecord(impl->inFree, impl->ctx->stream), "cudaEventRecord");
    impl->histories.resize(L);
    for (int l = 0; l < L; l++) {
        int nh = 0;
        is.read(reinterpret_cast<char *>(&nh), sizeof(int));
        impl->histories[l].resize(nh);
        for (int i = 0; i < nh; i++) {
            int size = 0, start = 0;
            is.read(reinterpret_cast<char *>(&size), sizeof(int));
            is.read(reinterpret_cast<char *>(&start), sizeof(int));
            CircleBuffer<IntBuffer> h;
            h.resize(size);
            h.start = start;
            for (int t = 0; t < size; t++) readBufferFromStream(is, &h[t]);
            impl->histories[l][i].alloc(size, size > 0 ? h.data[0].size() : 0, impl->ctx->stream);
            ringFromHost(impl->histories[l][i], *impl->ctx, h);
        }
        scLayers[l].readFromStream(is, cs);
        pLayers[l].resize(l == 0 ? NI : ticksPerUpdate[l]);
        for (size_t v = 0; v < pLayers[l].size(); v++) {
            char exists = 0;
            is.read(&exists, 1);
            if (exists) {
                pLayers[l][v].reset(new Predictor());
                pLayers[l][v]->readFromStream(is, cs);
            }
        }
    }
    aLayers.clear(); aLayers.resize(NI);
    for (int v = 0; v < NI; v++) {
        char exists = 0;
        is.read(&exists, 1);
        if (exists) {
            aLayers[v].reset(new Actor());
            aLayers[v]->readFromStream(is, cs);
            cs.replayRng = true;
        }
    }
    impl->inPinned.clear();
    for (int i = 0; i < NI; i++) {
        impl->inPinned.emplace_back(new PinnedBuf<int>());
        impl->inPinned.back()->alloc((size_t)inputSizes[i].x * inputSizes[i].y);
    }
}

// --- checkpoints (include/ogmaneo/Hierarchy.h) --------------------------------------------------------------------------
namespace {
const char kCkptMagic[8] = {'O', 'G', 'N', 'C', 'K', 'P', 'T', '1'};
template <typename T> void put(std::ostream &os, const T &v) { os.write(reinterpret_cast<const char *>(&v), sizeof(T)); }
template <typename T> T get(std::istream &is) {
    T v{};
    is.read(reinterpret_cast<char *>(&v), sizeof(T));
    if (!is) throw std::runtime_error("ogmaneo-b200: checkpoint is truncated");
    return v;
}
} // namespace

void Hierarchy::writeCheckpoint(std::ostream &os, const std::vector<InputType> &inputTypes, const std::vector<LayerDesc> &lds) const {
    const Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    const int L = (int)scLayers.size(), NI = (int)inputSizes.size();
    if ((int)inputTypes.size() != NI || (int)lds.size() != L) throw std::invalid_argument("ogmaneo-b200: checkpoint descriptors do not match");
    os.write(kCkptMagic, sizeof(kCkptMagic));
    put<int32_t>(os, NI); put<int32_t>(os, L); put<int32_t>(os, m.batch);
    put<int32_t>(os, scLayers[0].shardRank()); put<int32_t>(os, scLayers[0].shardWorld());
    for (int i = 0; i < NI; i++) { put<int32_t>(os, inputSizes[i].x); put<int32_t>(os, inputSizes[i].y); put<int32_t>(os, inputSizes[i].z); put<int32_t>(os, (int)inputTypes[i]); }
    for (int l = 0; l < L; l++) {
        const LayerDesc &d = lds[l];
        const int32_t v[9] = {d.hiddenSize.x, d.hiddenSize.y, d.hiddenSize.z, d.ffRadius, d.pRadius, d.ticksPerUpdate, d.temporalHorizon, d.aRadius, d.historyCapacity};
        os.write(reinterpret_cast<const char *>(v), sizeof(v));
    }
    for (int l = 0; l < L; l++) { put<int32_t>(os, ticks[l]); put<int32_t>(os, (int)updates[l]); }
    for (int l = 0; l < L; l++)
        for (const Ring &r : m.histories[l]) { put<int32_t>(os, r.T); put<int32_t>(os, r.start); writeDev(os, ctx, r.buf); }
    for (int l = 0; l < L; l++) {
        scLayers[l].writeDeviceState(os);
        for (auto &p : pLayers[l])
            if (p) p->writeDeviceState(os);
    }
    for (auto &a : aLayers)
        if (a) a->writeDeviceState(os);
    if (!os) throw std::runtime_error("ogmaneo-b200: checkpoint write failed");
}

void Hierarchy::readCheckpoint(std::istream &is, ComputeSystem &cs, std::vector<InputType> *typesOut) {
    char magic[8];
    is.read(magic, sizeof(magic));
    if (!is || std::memcmp(magic, kCkptMagic, sizeof(magic)) != 0) throw std::runtime_error("ogmaneo-b200: not a checkpoint file");
    const int NI = get<int32_t>(is), L = get<int32_t>(is), batch = get<int32_t>(is), rank = get<int32_t>(is), world = get<int32_t>(is);
    if (rank != cs.shardRank || world != cs.shardWorld)
        throw std::runtime_error("ogmaneo-b200: checkpoint of shard " + std::to_string(rank) + "/" + std::to_string(world) +
                                 " loaded into shard " + std::to_string(cs.shardRank) + "/" + std::to_string(cs.shardWorld));
    std::vector<Int3> sizes(NI);
    std::vector<InputType> types(NI);
    for (int i = 0; i < NI; i++) {
        sizes[i].x = get<int32_t>(is); sizes[i].y = get<int32_t>(is); sizes[i].z = get<int32_t>(is);
        types[i] = (InputType)get<int32_t>(is);
    }
    std::vector<LayerDesc> lds(L);
    for (int l = 0; l < L; l++) {
        int32_t v[9];
        is.read(reinterpret_cast<char *>(v), sizeof(v));
        lds[l].hiddenSize = Int3(v[0], v[1], v[2]); lds[l].ffRadius = v[3]; lds[l].pRadius = v[4]; lds[l].ticksPerUpdate = v[5];
        lds[l].temporalHorizon = v[6]; lds[l].aRadius = v[7]; lds[l].historyCapacity = v[8];
    }
    // same shapes and allocations as initRandom; the weights are overwritten below, so the cheap device-side init will do
    const bool savedInit = cs.deviceInit;
    cs.deviceInit = true;
    if (batch > 1) initRandomEnsemble(cs, std::vector<unsigned>(batch, 0u), sizes, types, lds);
    else initRandom(cs, sizes, types, lds);
    cs.deviceInit = savedInit;
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    for (int l = 0; l < L; l++) { ticks[l] = get<int32_t>(is); updates[l] = (char)get<int32_t>(is); }
    for (int l = 0; l < L; l++)
        for (Ring &r : m.histories[l]) {
            const int T = get<int32_t>(is), start = get<int32_t>(is);
            if (T != r.T) throw std::runtime_error("ogmaneo-b200: checkpoint history length mismatch");
            r.start = start;
            readDev(is, ctx, r.buf);
        }
    for (int l = 0; l < L; l++) {
        scLayers[l].readDeviceState(is);
        for (auto &p : pLayers[l])
            if (p) p->readDeviceState(is);
    }
    for (auto &a : aLayers)
        if (a) a->readDeviceState(is);
    if (typesOut) *typesOut = types;
}

void Hierarchy::readFromStream(std::istream &is) {
    ComputeSystem cs;
    if (impl) cs.device = impl->ctx->device;
    readFromStream(is, cs);
}

} // namespace ogmaneo

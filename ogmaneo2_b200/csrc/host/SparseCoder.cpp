// csrc/host/SparseCoder.cpp — ogmaneo::SparseCoder on the device (public interface: include/ogmaneo/SparseCoder.h).
#include "LayerImpl.h"

#include <algorithm>

namespace ogmaneo {
using namespace detail;

SparseCoder::SparseCoder() : alpha(0.1f), impl(nullptr) {}
SparseCoder::~SparseCoder() { delete impl; }
SparseCoder::SparseCoder(const SparseCoder &o) : alpha(0.1f), impl(nullptr) { *this = o; }

SparseCoder &SparseCoder::operator=(const SparseCoder &o) {
    if (this == &o) return *this;
    alpha = o.alpha; hiddenSize = o.hiddenSize; visibleLayerDescs = o.visibleLayerDescs;
    hiddenCs = o.hiddenCs; hiddenCsPrev = o.hiddenCsPrev; visibleLayers = o.visibleLayers;
    delete impl; impl = nullptr;
    if (o.impl) {
        impl = new Impl();
        DeviceContext &ctx = *o.impl->ctx;
        cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
        impl->ctx = o.impl->ctx; impl->batch = o.impl->batch; impl->x0 = o.impl->x0; impl->x1 = o.impl->x1;
        impl->world = o.impl->world; impl->rank = o.impl->rank; impl->cur = o.impl->cur; impl->prevAliasesCur = o.impl->prevAliasesCur;
        impl->csDirty = o.impl->csDirty; impl->wDirty = o.impl->wDirty; impl->reconDirty = o.impl->reconDirty;
        impl->ws.cloneFrom(ctx, o.impl->ws);
        impl->cs_[0].cloneFrom(o.impl->cs_[0], ctx.stream); impl->cs_[1].cloneFrom(o.impl->cs_[1], ctx.stream);
        impl->upd.allocZero(1, ctx.stream);
        impl->work.alloc(o.impl->work.size());
        impl->runOffset = o.impl->runOffset;
        impl->workCounters.allocZero(2, ctx.stream);
        impl->inStage.resize(o.impl->inStage.size());
        ctx.sync();
    }
    return *this;
}

void SparseCoder::Impl::allocate(ComputeSystem &cs, const Int3 &hidden, const std::vector<VisibleLayerDesc> &vlds, int batch_) {
    ctx = cs.sharedContext();
    batch = batch_; world = cs.shardWorld; rank = cs.shardRank;
    slabOf(hidden.x, cs.shardRank, cs.shardWorld, x0, x1);
    ws.sets.clear(); ws.sets.resize(vlds.size());
    for (size_t v = 0; v < vlds.size(); v++)
        ws.sets[v].allocate(*ctx, Geometry::get(hidden, vlds[v].size, vlds[v].radius), batch, true, true, x0, x1);
    const size_t ncols = (size_t)hidden.x * hidden.y * batch;
    if (ctx->arena) { // sharded with peer exchange: the CSDRs live in the exchange arena (same offset on every rank)
        cs_[0].borrow(ctx->arenaAlloc(ncols), ncols); cs_[1].borrow(ctx->arenaAlloc(ncols), ncols);
    } else {
        cs_[0].allocZero(ncols, ctx->stream); cs_[1].allocZero(ncols, ctx->stream);
    }
    upd.allocZero(1, ctx->stream);
    {   // one learn launch covers a run of equally-shaped visible layers: size the list for the largest run
        size_t need = 0, v = 0;
        runOffset.assign(1, 0);
        while (v < vlds.size()) {
            size_t e = v + 1;
            while (e < vlds.size() && ws.sets[e].geo == ws.sets[v].geo) e++;
            need += (e - v) * (size_t)(ws.sets[v].vx1 - ws.sets[v].vx0) * vlds[v].size.y;
            runOffset.push_back(need);
            v = e;
        }
        work.alloc(std::max<size_t>(need * batch, 1));
        workCounters.allocZero(2, ctx->stream);
    }
    inStage.clear(); inStage.resize(vlds.size());
    cur = 0; prevAliasesCur = true;
    csDirty = true; reconDirty = true; wDirty.assign(vlds.size(), 1);
}

// SparseCoder.cpp:85-125
void SparseCoder::initRandom(ComputeSystem &cs, const Int3 &hiddenSize_, const std::vector<VisibleLayerDesc> &vlds) {
    initRandomMember(cs, hiddenSize_, vlds, 1, 0);
}

void SparseCoder::initRandomMember(ComputeSystem &cs, const Int3 &hiddenSize_, const std::vector<VisibleLayerDesc> &vlds, int batch,
                                   int member) {
    if (member == 0) {
        hiddenSize = hiddenSize_; visibleLayerDescs = vlds;
        visibleLayers.clear(); visibleLayers.resize(vlds.size());
        delete impl; impl = new Impl();
        impl->allocate(cs, hiddenSize, vlds, batch);
    }
    for (size_t v = 0; v < vlds.size(); v++)
        impl->ws.sets[v].initialise(*impl->ctx, member, WeightInit{&cs, -1.0f, 0.0f, cs.deviceInitMatrix++});
}

// SparseCoder.cpp:127-147 with device-resident inputs. copyOut (optional) receives hiddenCs too (next layer's history).
void SparseCoder::stepDevice(ComputeSystem &cs, const int *const *inputCsDev, bool learnEnabled, int *copyOut, unsigned exchangeWaitMask) {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    const int nvl = (int)visibleLayerDescs.size();
    if (nvl > OGN_MAX_VL) throw std::runtime_error("ogmaneo-b200: more than 32 visible layers on one SparseCoder");
    ogn_fwd_args fa;
    std::memset(&fa, 0, sizeof(fa));
    int ngeo = 0;
    fa.nvl = nvl;
    for (int v = 0; v < nvl; v++) {
        const WeightSet &w = m.ws.sets[v];
        fa.vl[v].wf = w.fBase(); fa.vl[v].cs = inputCsDev[v]; fa.vl[v].mstride = w.fCount();
        fa.vl[v].geo = geoIndex(fa.geo, ngeo, *w.geo);
    }

    const int outIdx = m.prevAliasesCur ? 1 - m.cur : m.cur;
    int *out = m.cs_[outIdx].ptr();
    const int *prev = m.cs_[m.prevAliasesCur ? m.cur : 1 - m.cur].ptr();
    const bool sharded = m.world > 1;

    consumeKernel2(cs, hiddenSize.x, hiddenSize.y);
    if (ctx.collect) { // single-launch step: record the phase instead of launching it
        ogn_small_phase P;
        std::memset(&P, 0, sizeof(P));
        P.kind = OGN_PHASE_SC_FORWARD; P.z = hiddenSize.z; P.Hx = hiddenSize.x; P.Hy = hiddenSize.y;
        P.ppr = (hiddenSize.y + Geometry::packFactor(hiddenSize.z) - 1) / Geometry::packFactor(hiddenSize.z);
        P.npacks = hiddenSize.x * P.ppr;
        P.out0 = out; P.out1 = copyOut; P.u.fwd = fa;
        bool ok = ogn_vec_row_ok(hiddenSize.z) != 0 && !sharded;
        for (int v = 0; v < nvl && ok; v++) ok = ogn_vec_geo_ok(&fa.geo[fa.vl[v].geo]) != 0;
        ctx.collect->ok = ctx.collect->ok && ok;
        ctx.collect->add(P);
    } else {
    LaunchScope ls(ctx, kScForward);
    if (!ctx.dry) kernelCheck(ogn_sc_forward(&fa, hiddenSize.x, hiddenSize.y, hiddenSize.z, m.x0, m.x1, m.batch, out, sharded ? nullptr : copyOut,
                               ctx.stream), "sc_forward"); }
    if (sharded) {
        if (ctx.arena && ctx.inArena(out)) ctx.peerAllGather(out, (m.x1 - m.x0) * hiddenSize.y, false, copyOut ? 0u : exchangeWaitMask);
        else if (cs.allGather) cs.allGather(cs.allGatherUser, out, (m.x1 - m.x0) * hiddenSize.y, (void *)ctx.stream);
        else throw std::runtime_error("ogmaneo-b200: shardWorld > 1 needs ComputeSystem::allGather or peerExchange");
        if (copyOut)
            cudaCheck(cudaMemcpyAsync(copyOut, out, (size_t)hiddenSize.x * hiddenSize.y * sizeof(int), cudaMemcpyDeviceToDevice,
                                      ctx.stream), "hiddenCs D2D");
    }
    if (learnEnabled) {
        // learn touches only this layer's weights: queue it on the side stream, behind everything queued so far, so that
        // it overlaps whatever the caller queues next on the main stream (upper layers, the Predictors). Hierarchy::step
        // joins it at the end of the step; the host-facing step() synchronises both streams.
        cudaStream_t lst = ctx.stream;
        const bool side = ctx.useSide && !ctx.collect;
        if (side) { ctx.forkSide(); lst = ctx.side; }
        int v = 0, run = 0;
        while (v < nvl) { // one launch per run of visible layers with the same shape and radius
            int e = v + 1;
            while (e < nvl && m.ws.sets[e].geo == m.ws.sets[v].geo) e++;
            const WeightSet &w0 = m.ws.sets[v];
            ogn_learn_args la;
            std::memset(&la, 0, sizeof(la));
            la.geo = w0.geo->host; la.nvl = e - v; la.fx0 = w0.fx0; la.fx1 = w0.fx1;
            la.work = m.work.ptr() + (size_t)m.batch * m.runOffset[run]; la.work_counters = m.workCounters.ptr();
            for (int k = v; k < e; k++) {
                consumeKernel2(cs, visibleLayerDescs[k].size.x, visibleLayerDescs[k].size.y);
                const WeightSet &w = m.ws.sets[k];
                ogn_learn_vl &l = la.vl[k - v];
                l.wf = w.fBase(); l.wr = w.rBase(); l.recon = w.recon.ptr(); l.cs = inputCsDev[k];
                l.f_mstride = w.fCount(); l.r_mstride = w.rCount();
            }
            if (ctx.collect) { // single-launch step: reconstruction and update become two phases sharing one work list
                ogn_small_phase P;
                std::memset(&P, 0, sizeof(P));
                const int Vz = la.geo.Vz, pr = Geometry::packFactor(Vz);
                P.kind = OGN_PHASE_SC_RECON; P.z = Vz; P.Hx = la.geo.Vx; P.Hy = la.geo.Vy;
                P.ppr = (la.geo.Vy + pr - 1) / pr; P.npacks = la.geo.Vx * P.ppr;
                P.hcs = out; P.hcsPrev = prev; P.alpha = alpha; P.u.learn = la;
                P.work = la.work; P.workStride = (int64_t)(m.runOffset[run + 1] - m.runOffset[run]);
                P.updCount = m.upd.ptr();
                P.counter = ctx.collect->lists++;
                P.updVer = std::min(2, ogn_sc_update_version(&la)); P.chunks = ogn_sc_update_chunks(&la, P.updVer); P.updBatches = ogn_sc_update_batches();
                const bool ok = ogn_vec_row_ok(Vz) != 0 && ogn_vec_geo_ok(&la.geo) != 0 && !sharded && P.counter < OGN_SMALL_MAX_LISTS &&
                                P.chunks > 0;
                ctx.collect->ok = ctx.collect->ok && ok;
                ctx.collect->add(P);
                P.kind = OGN_PHASE_SC_UPDATE;
                ctx.collect->add(P);
                v = e; run++;
                continue;
            }
            // reconstruction (or the whole fused learn when the 128-bit path does not apply), then the work-list update
            { LaunchScope ls(ctx, kScLearn, lst);
            if (!ctx.dry) kernelCheck(ogn_sc_learn(&la, out, prev, w0.vx0, w0.vx1, alpha, m.batch, m.upd.ptr(), OGN_LEARN_RECONSTRUCT, lst),
                        "sc_learn"); }
            if (ogn_sc_learn_is_split(&la)) {
                LaunchScope ls(ctx, kScUpdate, lst);
                if (!ctx.dry) kernelCheck(ogn_sc_learn(&la, out, prev, w0.vx0, w0.vx1, alpha, m.batch, m.upd.ptr(), OGN_LEARN_UPDATE, lst),
                            "sc_update");
            }
            v = e; run++;
        }
        if (side) ctx.learnQueued();
        std::fill(m.wDirty.begin(), m.wDirty.end(), 1);
        m.reconDirty = true;
    }
    consumeKernel1(cs, hiddenSize.x * hiddenSize.y);
    m.cur = outIdx; m.prevAliasesCur = true; // hiddenCsPrev := hiddenCs without a copy
    m.csDirty = true;
}

void SparseCoder::step(ComputeSystem &cs, const std::vector<const IntBuffer *> &inputCs, bool learnEnabled) {
    Impl &m = *impl;
    cudaCheck(cudaSetDevice(m.ctx->device), "cudaSetDevice");
    std::vector<const int *> ptrs(inputCs.size());
    for (size_t v = 0; v < inputCs.size(); v++) {
        if (m.inStage[v].size() != inputCs[v]->size()) m.inStage[v].alloc(inputCs[v]->size());
        m.inStage[v].upload(inputCs[v]->data(), inputCs[v]->size(), m.ctx->stream);
        ptrs[v] = m.inStage[v].ptr();
    }
    stepDevice(cs, ptrs.data(), learnEnabled, nullptr);
    m.ctx->sync(); // the caller's buffers were pageable: make the H2D copies complete before returning
}

// --- host mirrors ------------------------------------------------------------------------------------------------
void SparseCoder::Impl::refreshCs(IntBuffer &hc, IntBuffer &hcPrev) {
    if (!csDirty) return;
    cudaCheck(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t n = cs_[0].size();
    hc.resize(n); hcPrev.resize(n);
    cs_[cur].download(hc.data(), n, ctx->stream);
    cs_[prevAliasesCur ? cur : 1 - cur].download(hcPrev.data(), n, ctx->stream);
    ctx->sync();
    csDirty = false;
}

const IntBuffer &SparseCoder::getHiddenCs() const { impl->refreshCs(hiddenCs, hiddenCsPrev); return hiddenCs; }
const IntBuffer &SparseCoder::getHiddenCsPrev() const { impl->refreshCs(hiddenCs, hiddenCsPrev); return hiddenCsPrev; }

void SparseCoder::getWeights(int i, std::vector<float> &out) const {
    cudaCheck(cudaSetDevice(impl->ctx->device), "cudaSetDevice");
    const WeightSet &w = impl->ws.sets[i];
    out.assign((size_t)w.geo->nnz, 0.0f);
    w.downloadCSR(*impl->ctx, 0, out.data());
}

void SparseCoder::getReconstructions(int i, FloatBuffer &out) const {
    cudaCheck(cudaSetDevice(impl->ctx->device), "cudaSetDevice");
    const WeightSet &w = impl->ws.sets[i];
    out.resize(w.recon.size());
    w.recon.download(out.data(), out.size(), impl->ctx->stream);
    impl->ctx->sync();
}

const SparseCoder::VisibleLayer &SparseCoder::getVisibleLayer(int i) const {
    VisibleLayer &vl = visibleLayers[i];
    if (impl->wDirty[i]) {
        if (vl.weights.rowRanges.empty()) { // pattern once, values every time they changed
            initSMLocalRF(visibleLayerDescs[i].size, hiddenSize, visibleLayerDescs[i].radius, vl.weights);
            vl.weights.initT();
        }
        getWeights(i, vl.weights.nonZeroValues);
        impl->wDirty[i] = 0;
    }
    if (impl->reconDirty) {
        for (size_t v = 0; v < visibleLayers.size(); v++) getReconstructions((int)v, visibleLayers[v].reconstructions);
        impl->reconDirty = false;
    }
    return vl;
}

unsigned long long SparseCoder::debugFRMismatch(int i) const {
    DeviceContext &ctx = *impl->ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    const WeightSet &w = impl->ws.sets[i];
    const ogn_wset d = w.desc();
    DevBuf<unsigned long long> cnt;
    cnt.allocZero(1, ctx.stream);
    unsigned long long total = 0;
    for (int m = 0; m < impl->batch; m++) kernelCheck(ogn_weights_fr_mismatch(&d, &w.geo->host, m, cnt.ptr(), ctx.stream), "fr_mismatch");
    cnt.download(&total, 1, ctx.stream);
    ctx.sync();
    return total;
}

void SparseCoder::getColumnWeights(int i, int ox, int oy, std::vector<float> &out) const {
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
    ctx.joinSide(); // learn may still be rewriting weights on the side stream
    kernelCheck(ogn_weights_to_csr(&d, &g.host, 0, 0, col, col + 1, stage.ptr(), base, ctx.stream), "weights_to_csr");
    out.resize((size_t)n);
    stage.download(out.data(), (size_t)n, ctx.stream);
    ctx.sync();
}

unsigned long long SparseCoder::takeUpdateCount() {
    cudaCheck(cudaSetDevice(impl->ctx->device), "cudaSetDevice");
    unsigned long long v = 0;
    impl->upd.download(&v, 1, impl->ctx->stream);
    cudaCheck(cudaMemsetAsync(impl->upd.ptr(), 0, sizeof(v), impl->ctx->stream), "memset");
    impl->ctx->sync();
    return v;
}

// --- streams: SparseCoder.cpp:149-203, layout in SURVEY.md Appendix B ------------------------------------------------
void SparseCoder::writeToStream(std::ostream &os) const {
    if (impl->world > 1 || impl->batch > 1) throw std::runtime_error("ogmaneo-b200: writeToStream needs an unsharded single hierarchy");
    os.write(reinterpret_cast<const char *>(&hiddenSize), sizeof(Int3));
    os.write(reinterpret_cast<const char *>(&alpha), sizeof(float));
    writeBufferToStream(os, &getHiddenCs());
    writeBufferToStream(os, &getHiddenCsPrev());
    int n = (int)visibleLayerDescs.size();
    os.write(reinterpret_cast<const char *>(&n), sizeof(int));
    for (int v = 0; v < n; v++) {
        os.write(reinterpret_cast<const char *>(&visibleLayerDescs[v]), sizeof(VisibleLayerDesc));
        writeSMToStream(os, getVisibleLayer(v).weights);
    }
}

void SparseCoder::readFromStream(std::istream &is, ComputeSystem &cs) {
    is.read(reinterpret_cast<char *>(&hiddenSize), sizeof(Int3));
    hiddenSize.pad = 0;
    is.read(reinterpret_cast<char *>(&alpha), sizeof(float));
    IntBuffer hc, hcPrev;
    readBufferFromStream(is, &hc);
    readBufferFromStream(is, &hcPrev);
    int n = 0;
    is.read(reinterpret_cast<char *>(&n), sizeof(int));
    visibleLayerDescs.resize(n);
    std::vector<SparseMatrix> mats(n);
    for (int v = 0; v < n; v++) {
        is.read(reinterpret_cast<char *>(&visibleLayerDescs[v]), sizeof(VisibleLayerDesc));
        visibleLayerDescs[v].size.pad = 0;
        readSMFromStream(is, mats[v]);
    }
    visibleLayers.clear(); visibleLayers.resize(n);
    delete impl; impl = new Impl();
    impl->allocate(cs, hiddenSize, visibleLayerDescs, 1);
    for (int v = 0; v < n; v++) {
        if ((int64_t)mats[v].nonZeroValues.size() != impl->ws.sets[v].geo->nnz)
            throw std::runtime_error("ogmaneo-b200: SparseCoder weight count in stream does not match its geometry");
        impl->ws.sets[v].uploadCSR(*impl->ctx, 0, mats[v].nonZeroValues.data());
    }
    setHiddenState(hc, hcPrev);
}

void SparseCoder::readFromStream(std::istream &is) {
    ComputeSystem cs; // default device
    if (impl) cs.device = impl->ctx->device;
    readFromStream(is, cs);
}

void SparseCoder::setHiddenState(const IntBuffer &hc, const IntBuffer &hcPrev) {
    Impl &m = *impl;
    cudaCheck(cudaSetDevice(m.ctx->device), "cudaSetDevice");
    m.cs_[0].upload(hc.data(), hc.size(), m.ctx->stream);
    m.cs_[1].upload(hcPrev.data(), hcPrev.size(), m.ctx->stream);
    m.ctx->sync();
    m.cur = 0; m.prevAliasesCur = false;
    hiddenCs = hc; hiddenCsPrev = hcPrev; m.csDirty = false;
}

const int *SparseCoder::hiddenCsDevice() const { return impl->cs_[impl->cur].ptr(); }

void SparseCoder::learnRows(int &lo, int &hi) const {
    lo = impl->x0; hi = impl->x1;
    for (const WeightSet &w : impl->ws.sets) { lo = std::min(lo, w.rx0); hi = std::max(hi, w.rx1); }
}
void SparseCoder::visibleRows(int vli, int &lo, int &hi) const { lo = impl->ws.sets[vli].vx0; hi = impl->ws.sets[vli].vx1; }
int SparseCoder::shardRank() const { return impl->rank; }
int SparseCoder::shardWorld() const { return impl->world; }

void SparseCoder::writeDeviceState(std::ostream &os) const {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    const int hdr[4] = {m.cur, m.prevAliasesCur ? 1 : 0, m.x0, m.x1};
    os.write(reinterpret_cast<const char *>(hdr), sizeof(hdr));
    os.write(reinterpret_cast<const char *>(&alpha), sizeof(alpha));
    writeDev(os, ctx, m.cs_[0]); writeDev(os, ctx, m.cs_[1]);
    for (const WeightSet &w : m.ws.sets) { writeDev(os, ctx, w.wf); writeDev(os, ctx, w.wr); writeDev(os, ctx, w.recon); }
}

void SparseCoder::readDeviceState(std::istream &is) {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    int hdr[4] = {0, 0, 0, 0};
    is.read(reinterpret_cast<char *>(hdr), sizeof(hdr));
    if (hdr[2] != m.x0 || hdr[3] != m.x1) throw std::runtime_error("ogmaneo-b200: checkpoint was written by a different shard");
    m.cur = hdr[0]; m.prevAliasesCur = hdr[1] != 0;
    is.read(reinterpret_cast<char *>(&alpha), sizeof(alpha));
    readDev(is, ctx, m.cs_[0]); readDev(is, ctx, m.cs_[1]);
    for (WeightSet &w : m.ws.sets) { readDev(is, ctx, w.wf); readDev(is, ctx, w.wr); readDev(is, ctx, w.recon); }
    m.csDirty = m.reconDirty = true;
    std::fill(m.wDirty.begin(), m.wDirty.end(), 1);
}

int SparseCoder::smallUnits() const {
    const Impl &m = *impl;
    if (m.world > 1 || !ogn_vec_row_ok(hiddenSize.z)) return -1;
    const int pf = Geometry::packFactor(hiddenSize.z);
    int units = hiddenSize.x * ((hiddenSize.y + pf - 1) / pf);
    size_t v = 0;
    while (v < m.ws.sets.size()) {
        size_t e = v + 1;
        while (e < m.ws.sets.size() && m.ws.sets[e].geo == m.ws.sets[v].geo) e++;
        const Geometry &g = *m.ws.sets[v].geo;
        if (!ogn_vec_row_ok(g.Vz) || !ogn_vec_geo_ok(&g.host) || g.host.max_cov <= 0) return -1;
        const int pr = Geometry::packFactor(g.Vz);
        units = std::max(units, (int)(e - v) * g.Vx * ((g.Vy + pr - 1) / pr));
        v = e;
    }
    return units;
}

void SparseCoder::stateSignature(std::vector<int> &key) const {
    key.push_back(impl->cur); key.push_back(impl->prevAliasesCur ? 1 : 0);
    int a;
    std::memcpy(&a, &alpha, sizeof(a));
    key.push_back(a);
}

} // namespace ogmaneo

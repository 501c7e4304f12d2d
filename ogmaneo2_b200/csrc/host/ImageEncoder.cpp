// csrc/host/ImageEncoder.cpp — ogmaneo::ImageEncoder on the device (public interface: include/ogmaneo/ImageEncoder.h).
#include <ogmaneo/ImageEncoder.h>

#include "Device.h"

namespace ogmaneo {
using namespace detail;

struct ImageEncoder::Impl {
    std::shared_ptr<DeviceContext> ctx;
    WeightSetArray ws;                       // F layout only: every kernel walks hidden-column packs
    DevBuf<int> hiddenCs, hcStage;
    DevBuf<float> resources;
    std::vector<DevBuf<float>> recon, inStage;
    bool csDirty = true, resDirty = true, reconDirty = true;
    std::vector<char> wDirty;

    void allocate(ComputeSystem &cs, const Int3 &hidden, const std::vector<VisibleLayerDesc> &vlds) {
        if (cs.shardWorld > 1) throw std::runtime_error("ogmaneo-b200: ImageEncoder layers cannot be sharded");
        if (hidden.z > 32 || hidden.z <= 0) throw std::runtime_error("ogmaneo-b200: ImageEncoder supports up to 32 cells per hidden column");
        if (vlds.size() > OGN_MAX_IENC_VL) throw std::runtime_error("ogmaneo-b200: more than 8 visible layers on one ImageEncoder");
        ctx = cs.sharedContext();
        ws.sets.clear(); ws.sets.resize(vlds.size());
        recon.clear(); recon.resize(vlds.size());
        inStage.clear(); inStage.resize(vlds.size());
        for (size_t v = 0; v < vlds.size(); v++) {
            ws.sets[v].allocate(*ctx, Geometry::get(hidden, vlds[v].size, vlds[v].radius), 1, false, false, 0, hidden.x);
            const size_t n = (size_t)vlds[v].size.x * vlds[v].size.y * vlds[v].size.z;
            recon[v].allocZero(n, ctx->stream);
            inStage[v].alloc(n);
        }
        const size_t ncols = (size_t)hidden.x * hidden.y;
        hiddenCs.allocZero(ncols, ctx->stream);
        hcStage.alloc(ncols);
        resources.alloc(ncols * hidden.z);
        std::vector<float> ones(ncols * hidden.z, 1.0f);
        resources.upload(ones.data(), ones.size(), ctx->stream);
        ctx->sync();
        csDirty = resDirty = reconDirty = true;
        wDirty.assign(vlds.size(), 1);
    }
    ogn_ienc_args args() const {
        ogn_ienc_args a;
        std::memset(&a, 0, sizeof(a));
        int ngeo = 0;
        a.nvl = (int)ws.sets.size();
        for (int v = 0; v < a.nvl; v++) {
            a.vl[v].wf = ws.sets[v].fBase(); a.vl[v].in = inStage[v].ptr(); a.vl[v].recon = recon[v].ptr();
            a.vl[v].geo = geoIndex(a.geo, ngeo, *ws.sets[v].geo);
        }
        return a;
    }
};

ImageEncoder::ImageEncoder() : alpha(0.01f), gamma(0.01f), impl(nullptr) {}
ImageEncoder::~ImageEncoder() { delete impl; }
ImageEncoder::ImageEncoder(const ImageEncoder &o) : alpha(0.01f), gamma(0.01f), impl(nullptr) { *this = o; }

ImageEncoder &ImageEncoder::operator=(const ImageEncoder &o) {
    if (this == &o) return *this;
    alpha = o.alpha; gamma = o.gamma; hiddenSize = o.hiddenSize; visibleLayerDescs = o.visibleLayerDescs;
    hiddenCs = o.hiddenCs; hiddenResources = o.hiddenResources; visibleLayers = o.visibleLayers;
    delete impl; impl = nullptr;
    if (o.impl) {
        impl = new Impl();
        const Impl &s = *o.impl;
        DeviceContext &ctx = *s.ctx;
        cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
        impl->ctx = s.ctx;
        impl->ws.cloneFrom(ctx, s.ws);
        impl->hiddenCs.cloneFrom(s.hiddenCs, ctx.stream); impl->hcStage.alloc(s.hcStage.size());
        impl->resources.cloneFrom(s.resources, ctx.stream);
        impl->recon.resize(s.recon.size()); impl->inStage.resize(s.inStage.size());
        for (size_t v = 0; v < s.recon.size(); v++) { impl->recon[v].cloneFrom(s.recon[v], ctx.stream); impl->inStage[v].alloc(s.inStage[v].size()); }
        impl->csDirty = impl->resDirty = impl->reconDirty = true;
        impl->wDirty.assign(s.wDirty.size(), 1);
        ctx.sync();
    }
    return *this;
}

void ImageEncoder::initRandom(ComputeSystem &cs, const Int3 &hiddenSize_, const std::vector<VisibleLayerDesc> &vlds) {
    hiddenSize = hiddenSize_; hiddenSize.pad = 0;
    visibleLayerDescs = vlds;
    for (auto &d : visibleLayerDescs) d.size.pad = 0;
    visibleLayers.clear(); visibleLayers.resize(vlds.size());
    delete impl; impl = new Impl();
    impl->allocate(cs, hiddenSize, visibleLayerDescs);
    for (size_t v = 0; v < vlds.size(); v++)
        impl->ws.sets[v].initialise(*impl->ctx, 0, WeightInit{&cs, 0.0f, 1.0f, cs.deviceInitMatrix++});
}

void ImageEncoder::stepHost(ComputeSystem &cs, const float *const *inputActs, bool learnEnabled) {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    for (size_t v = 0; v < m.inStage.size(); v++) m.inStage[v].upload(inputActs[v], m.inStage[v].size(), ctx.stream);
    const ogn_ienc_args a = m.args();
    consumeKernel2(cs, hiddenSize.x, hiddenSize.y);
    kernelCheck(ogn_image_encoder_step(&a, hiddenSize.x, hiddenSize.y, hiddenSize.z, learnEnabled ? 1 : 0, alpha, gamma, m.resources.ptr(),
                              m.hiddenCs.ptr(), ctx.stream), "ienc_step");
    ctx.sync(); // the caller's buffers may be pageable: the uploads must have completed before returning
    m.csDirty = true;
    if (learnEnabled) { m.resDirty = true; std::fill(m.wDirty.begin(), m.wDirty.end(), 1); }
}

void ImageEncoder::stepDevice(ComputeSystem &cs, const float *const *inputActsDevice, bool learnEnabled) {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    ogn_ienc_args a = m.args();
    for (int v = 0; v < a.nvl; v++) {
        if (!inputActsDevice[v]) throw std::invalid_argument("ogmaneo-b200: null device input");
        a.vl[v].in = inputActsDevice[v];
    }
    consumeKernel2(cs, hiddenSize.x, hiddenSize.y);
    kernelCheck(ogn_image_encoder_step(&a, hiddenSize.x, hiddenSize.y, hiddenSize.z, learnEnabled ? 1 : 0, alpha, gamma, m.resources.ptr(),
                              m.hiddenCs.ptr(), ctx.stream), "ienc_step");
    m.csDirty = true; // the host mirrors refresh (and wait) on demand
    if (learnEnabled) { m.resDirty = true; std::fill(m.wDirty.begin(), m.wDirty.end(), 1); }
}

const int *ImageEncoder::getHiddenCsDevice() const { return impl ? impl->hiddenCs.ptr() : nullptr; }

void ImageEncoder::step(ComputeSystem &cs, const std::vector<const FloatBuffer *> &inputActs, bool learnEnabled) {
    if (inputActs.size() != visibleLayerDescs.size()) throw std::invalid_argument("ogmaneo-b200: wrong number of dense inputs");
    std::vector<const float *> raw(inputActs.size());
    for (size_t v = 0; v < inputActs.size(); v++) {
        if (inputActs[v]->size() != impl->inStage[v].size()) throw std::invalid_argument("ogmaneo-b200: dense input has the wrong size");
        raw[v] = inputActs[v]->data();
    }
    stepHost(cs, raw.data(), learnEnabled);
}

void ImageEncoder::reconstructHost(ComputeSystem &cs, const int *hcs) {
    Impl &m = *impl;
    DeviceContext &ctx = *m.ctx;
    cudaCheck(cudaSetDevice(ctx.device), "cudaSetDevice");
    m.hcStage.upload(hcs, m.hcStage.size(), ctx.stream);
    const ogn_ienc_args a = m.args();
    for (int v = 0; v < a.nvl; v++) {
        consumeKernel2(cs, visibleLayerDescs[v].size.x, visibleLayerDescs[v].size.y);
        kernelCheck(ogn_image_encoder_reconstruct(&a, v, hiddenSize.x, hiddenSize.y, hiddenSize.z, m.hcStage.ptr(), ctx.stream), "ienc_reconstruct");
    }
    ctx.sync();
    m.reconDirty = true;
}

void ImageEncoder::reconstruct(ComputeSystem &cs, const IntBuffer *hcs) {
    if (hcs->size() != impl->hcStage.size()) throw std::invalid_argument("ogmaneo-b200: hidden CSDR has the wrong number of columns");
    reconstructHost(cs, hcs->data());
}

const IntBuffer &ImageEncoder::getHiddenCs() const {
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

const FloatBuffer &ImageEncoder::getHiddenResources() const {
    Impl &m = *impl;
    if (m.resDirty) {
        cudaCheck(cudaSetDevice(m.ctx->device), "cudaSetDevice");
        hiddenResources.resize(m.resources.size());
        m.resources.download(hiddenResources.data(), hiddenResources.size(), m.ctx->stream);
        m.ctx->sync();
        m.resDirty = false;
    }
    return hiddenResources;
}

void ImageEncoder::getWeights(int i, std::vector<float> &out) const {
    cudaCheck(cudaSetDevice(impl->ctx->device), "cudaSetDevice");
    const WeightSet &w = impl->ws.sets[i];
    out.assign((size_t)w.geo->nnz, 0.0f);
    w.downloadCSR(*impl->ctx, 0, out.data());
}

void ImageEncoder::getReconstruction(int i, FloatBuffer &out) const {
    cudaCheck(cudaSetDevice(impl->ctx->device), "cudaSetDevice");
    out.resize(impl->recon[i].size());
    impl->recon[i].download(out.data(), out.size(), impl->ctx->stream);
    impl->ctx->sync();
}

const ImageEncoder::VisibleLayer &ImageEncoder::getVisibleLayer(int i) const {
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
        for (size_t v = 0; v < visibleLayers.size(); v++) getReconstruction((int)v, visibleLayers[v].reconstructions);
        impl->reconDirty = false;
    }
    return vl;
}

// ImageEncoder.cpp:162-195 (same bytes; Int3::pad written as zero)
void ImageEncoder::writeToStream(std::ostream &os) const {
    os.write(reinterpret_cast<const char *>(&hiddenSize), sizeof(Int3));
    os.write(reinterpret_cast<const char *>(&alpha), sizeof(float));
    os.write(reinterpret_cast<const char *>(&gamma), sizeof(float));
    writeBufferToStream(os, &getHiddenCs());
    writeBufferToStream(os, &getHiddenResources());
    int n = (int)visibleLayerDescs.size();
    os.write(reinterpret_cast<const char *>(&n), sizeof(int));
    for (int v = 0; v < n; v++) {
        const VisibleLayer &vl = getVisibleLayer(v);
        os.write(reinterpret_cast<const char *>(&visibleLayerDescs[v]), sizeof(VisibleLayerDesc));
        writeSMToStream(os, vl.weights);
        writeBufferToStream(os, &vl.reconstructions);
    }
}

// ImageEncoder.cpp:197-225
void ImageEncoder::readFromStream(std::istream &is, ComputeSystem &cs) {
    is.read(reinterpret_cast<char *>(&hiddenSize), sizeof(Int3));
    hiddenSize.pad = 0;
    is.read(reinterpret_cast<char *>(&alpha), sizeof(float));
    is.read(reinterpret_cast<char *>(&gamma), sizeof(float));
    IntBuffer hc; FloatBuffer res;
    readBufferFromStream(is, &hc);
    readBufferFromStream(is, &res);
    int n = 0;
    is.read(reinterpret_cast<char *>(&n), sizeof(int));
    visibleLayerDescs.resize(n);
    std::vector<SparseMatrix> mats(n);
    std::vector<FloatBuffer> recs(n);
    for (int v = 0; v < n; v++) {
        is.read(reinterpret_cast<char *>(&visibleLayerDescs[v]), sizeof(VisibleLayerDesc));
        visibleLayerDescs[v].size.pad = 0;
        readSMFromStream(is, mats[v]);
        readBufferFromStream(is, &recs[v]);
    }
    visibleLayers.clear(); visibleLayers.resize(n);
    delete impl; impl = new Impl();
    impl->allocate(cs, hiddenSize, visibleLayerDescs);
    Impl &m = *impl;
    for (int v = 0; v < n; v++) {
        if ((int64_t)mats[v].nonZeroValues.size() != m.ws.sets[v].geo->nnz)
            throw std::runtime_error("ogmaneo-b200: ImageEncoder weight count in stream does not match its geometry");
        m.ws.sets[v].uploadCSR(*m.ctx, 0, mats[v].nonZeroValues.data());
        if (recs[v].size() == m.recon[v].size()) m.recon[v].upload(recs[v].data(), recs[v].size(), m.ctx->stream);
    }
    if (hc.size() != m.hiddenCs.size() || res.size() != m.resources.size())
        throw std::runtime_error("ogmaneo-b200: ImageEncoder state in stream does not match its shape");
    m.hiddenCs.upload(hc.data(), hc.size(), m.ctx->stream);
    m.resources.upload(res.data(), res.size(), m.ctx->stream);
    m.ctx->sync();
}

void ImageEncoder::readFromStream(std::istream &is) {
    ComputeSystem cs;
    if (impl) cs.device = impl->ctx->device;
    readFromStream(is, cs);
}

} // namespace ogmaneo

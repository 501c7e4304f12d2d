// csrc/host/Device.cpp — CUDA context, geometry tables and weight sets (see Device.h).
#include "Device.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

namespace ogmaneo {
namespace detail {

void cudaCheck(cudaError_t e, const char *what) {
    if (e != cudaSuccess)
        throw std::runtime_error(std::string("ogmaneo-b200: ") + what + ": " + cudaGetErrorString(e) +
                                 " (this build has no CPU fallback)");
}

DeviceContext::DeviceContext(int dev) : device(dev), stream(nullptr) {
    int count = 0;
    cudaCheck(cudaGetDeviceCount(&count), "cudaGetDeviceCount");
    if (count == 0) throw std::runtime_error("ogmaneo-b200: no CUDA device visible (this build has no CPU fallback)");
    if (device < 0) cudaCheck(cudaGetDevice(&device), "cudaGetDevice");
    cudaCheck(cudaSetDevice(device), "cudaSetDevice");
    cudaDeviceProp prop;
    cudaCheck(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (prop.major != 10)
        throw std::runtime_error("ogmaneo-b200: device " + std::to_string(device) + " (" + prop.name +
                                 ") is not compute capability 10.x; the kernels are built for sm_100a only");
    if (const char *g = std::getenv("OGN_L2_FETCH_GRANULARITY")) // experiment knob: 32 / 64 / 128 (a hint to the driver)
        cudaCheck(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)std::atoi(g)), "cudaLimitMaxL2FetchGranularity");
    // OGN_STREAM_PRIO: "main" / "side" gives that stream the device's highest priority and the other the lowest, so that
    // when kernels of both streams have blocks pending the block scheduler always prefers the same one (experiment knob,
    // see profiles/README.md); unset: two ordinary streams
    const char *prio = std::getenv("OGN_STREAM_PRIO");
    if (prio && (!std::strcmp(prio, "main") || !std::strcmp(prio, "side"))) {
        int least = 0, greatest = 0;
        cudaCheck(cudaDeviceGetStreamPriorityRange(&least, &greatest), "cudaDeviceGetStreamPriorityRange");
        const bool mainFirst = !std::strcmp(prio, "main");
        cudaCheck(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, mainFirst ? greatest : least), "cudaStreamCreate");
        cudaCheck(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, mainFirst ? least : greatest), "cudaStreamCreate(side)");
    } else {
        cudaCheck(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
        cudaCheck(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking), "cudaStreamCreate(side)");
    }
    cudaCheck(cudaEventCreateWithFlags(&evFork, cudaEventDisableTiming), "cudaEventCreate");
    cudaCheck(cudaEventCreateWithFlags(&evLearn, cudaEventDisableTiming), "cudaEventCreate");
    cudaCheck(cudaEventCreateWithFlags(&evExchange, cudaEventDisableTiming), "cudaEventCreate");
    cudaCheck(cudaEventCreateWithFlags(&evPredLearn, cudaEventDisableTiming), "cudaEventCreate");
    if (const char *e = std::getenv("OGN_SIDE_STREAM")) { useSide = std::atoi(e) != 0; sideForced = true; } // 0: everything on one stream
    if (const char *e = std::getenv("OGN_GRAPHS")) graphs = std::atoi(e) != 0;
    if (const char *e = std::getenv("OGN_SMALL_KERNEL")) small = std::atoi(e) != 0;
    if (std::getenv("OGN_KERNELS")) small = false; // a forced kernel variant means the per-phase launches are under test
    if (const char *e = std::getenv("OGN_SMALL_MAX_UNITS")) smallMaxUnits = std::atoi(e);
}

// In a dry run (the step's host code replayed for its state transitions while a captured graph does the work) nothing is
// queued: the graph holds the fork/join edges that were captured.
void DeviceContext::forkSide() {
    if (dry) return;
    cudaCheck(cudaEventRecord(evFork, stream), "cudaEventRecord");
    cudaCheck(cudaStreamWaitEvent(side, evFork, 0), "cudaStreamWaitEvent");
}
void DeviceContext::learnQueued() {
    if (dry) return;
    cudaCheck(cudaEventRecord(evLearn, side), "cudaEventRecord");
    learnPending = true;
}
void DeviceContext::exchangeQueued() {
    cudaCheck(cudaEventRecord(evExchange, side), "cudaEventRecord");
    exchangePending = true;
}
void DeviceContext::joinLearn() {
    if (!learnPending) return;
    cudaCheck(cudaStreamWaitEvent(stream, evLearn, 0), "cudaStreamWaitEvent");
    learnPending = false;
}
void DeviceContext::joinExchange() {
    if (!exchangePending) return;
    cudaCheck(cudaStreamWaitEvent(stream, evExchange, 0), "cudaStreamWaitEvent");
    exchangePending = false;
}

// ---- peer arena --------------------------------------------------------------------------------------------------
void DeviceContext::arenaCreate(size_t bytes, int rank, int world) {
    if (world > OGN_MAX_PEERS) throw std::runtime_error("ogmaneo-b200: at most 16 devices can share a peer exchange");
    arena.reset(new PeerArena());
    arena->capInts = std::max<size_t>(bytes / sizeof(int), PeerArena::kHeaderInts * 2);
    arena->rank = rank; arena->world = world;
    cudaCheck(cudaMalloc((void **)&arena->base, arena->capInts * sizeof(int)), "cudaMalloc(peer arena)");
    cudaCheck(cudaMemset(arena->base, 0, arena->capInts * sizeof(int)), "cudaMemset(peer arena)");
    cudaCheck(cudaHostAlloc((void **)&arena->errHost, sizeof(unsigned), cudaHostAllocMapped), "cudaHostAlloc");
    *arena->errHost = 0;
    cudaCheck(cudaHostGetDevicePointer((void **)&arena->errDev, arena->errHost, 0), "cudaHostGetDevicePointer");
    arena->peer[rank] = arena->base;
}

int *DeviceContext::arenaAlloc(size_t nInts) {
    PeerArena &a = *arena;
    const size_t start = (a.usedInts + 63) & ~(size_t)63; // 256-byte aligned: slabs stay 16-byte aligned for 128-bit stores
    if (start + nInts > a.capInts)
        throw std::runtime_error("ogmaneo-b200: peer exchange arena exhausted (raise OGN_PEER_ARENA_MB)");
    a.usedInts = start + nInts;
    return a.base + start;
}

void DeviceContext::arenaExport(void *handle64) const {
    if (!arena) throw std::runtime_error("ogmaneo-b200: no peer arena (create with peer_exchange and shard_world > 1)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    cudaCheck(cudaIpcGetMemHandle(&h, arena->base), "cudaIpcGetMemHandle");
    std::memcpy(handle64, &h, sizeof(h));
}

void DeviceContext::arenaAttach(const void *handles64, int world) {
    if (!arena) throw std::runtime_error("ogmaneo-b200: no peer arena (create with peer_exchange and shard_world > 1)");
    if (world != arena->world) throw std::invalid_argument("ogmaneo-b200: peer attach: world size mismatch");
    cudaCheck(cudaSetDevice(device), "cudaSetDevice");
    for (int r = 0; r < world; r++) {
        if (r == arena->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char *)handles64 + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        cudaCheck(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
        arena->peer[r] = (int *)p;
    }
    arena->attached = true;
}

void DeviceContext::peerAllGather(int *buffer, int countPerRank, bool deferred, unsigned waitMask) {
    PeerArena &a = *arena;
    deferred = deferred && useSide;
    // exchanges complete in the same order on every rank: an earlier deferred one is joined before the next is queued
    cudaStream_t st = stream;
    if (deferred) { forkSide(); st = side; }
    else joinExchange();
    if (!a.attached) throw std::runtime_error("ogmaneo-b200: peer exchange used before ogn_peer_attach");
    if (!inArena(buffer)) throw std::runtime_error("ogmaneo-b200: peer all-gather of a buffer outside the exchange arena");
    ogn_peer_table t;
    std::memset(&t, 0, sizeof(t));
    for (int r = 0; r < a.world; r++) t.base[r] = a.peer[r];
    t.rank = a.rank; t.world = a.world;
    static const unsigned long long timeoutNs = [] {
        const char *e = std::getenv("OGN_PEER_TIMEOUT_MS");
        return (unsigned long long)(e ? std::atoll(e) : 5000) * 1000000ull;
    }();
    const unsigned everybody = (a.world >= 32 ? 0xffffffffu : ((1u << a.world) - 1u)) & ~(1u << a.rank);
    static const bool haloWait = [] { const char *e = std::getenv("OGN_PEER_HALO_WAIT"); return !e || std::atoi(e) != 0; }();
    const unsigned mask = (waitMask && haloWait) ? (waitMask & everybody) : everybody;
    { LaunchScope ls(*this, kExchange, st);
    kernelCheck(ogn_peer_allgather(&t, buffer - a.base, countPerRank, ++a.epoch, mask, timeoutNs, a.errDev, st), "peer_allgather"); }
    if (deferred) exchangeQueued();
}

unsigned long long DeviceContext::peerTakeWaitNs() {
    if (!arena) return 0;
    unsigned long long v = 0;
    cudaCheck(cudaMemcpyAsync(&v, arena->base + 18, sizeof(v), cudaMemcpyDeviceToHost, stream), "D2H");
    cudaCheck(cudaMemsetAsync(arena->base + 18, 0, sizeof(v), stream), "memset");
    sync();
    return v;
}

void DeviceContext::peerCheck() const {
    if (arena && arena->errHost && *(volatile unsigned *)arena->errHost)
        throw std::runtime_error("ogmaneo-b200: a peer all-gather timed out waiting for another rank");
}

DeviceContext::~DeviceContext() {
    if (arena) {
        for (int r = 0; r < arena->world; r++)
            if (r != arena->rank && arena->peer[r]) cudaIpcCloseMemHandle(arena->peer[r]);
        if (arena->base) cudaFree(arena->base);
        if (arena->errHost) cudaFreeHost(arena->errHost);
    }
    for (auto &r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : eventPool) cudaEventDestroy(e);
    if (evFork) cudaEventDestroy(evFork);
    if (evLearn) cudaEventDestroy(evLearn);
    if (evExchange) cudaEventDestroy(evExchange);
    if (evPredLearn) cudaEventDestroy(evPredLearn);
    if (side) cudaStreamDestroy(side);
    if (stream) cudaStreamDestroy(stream);
}

cudaEvent_t DeviceContext::takeEvent() {
    if (!eventPool.empty()) { cudaEvent_t e = eventPool.back(); eventPool.pop_back(); return e; }
    cudaEvent_t e;
    cudaCheck(cudaEventCreate(&e), "cudaEventCreate");
    return e;
}

void DeviceContext::profileCollect() {
    sync();
    peerCheck();
    // OGN_TIMELINE=n: print the first n bracketed launches as a timeline (kind, stream, start relative to the first one,
    // duration; microseconds) to stderr: where a sharded step's time goes across the two streams
    static int timeline = [] { const char *e = std::getenv("OGN_TIMELINE"); return e ? std::atoi(e) : 0; }();
    if (timeline > 0 && !recs.empty()) {
        static const char *names[kKernelKinds] = {"sc_forward", "sc_recon", "pred_step", "actor_fwd", "actor_learn", "sc_update", "hier_small", "exchange"};
        const size_t skip = recs.size() > (size_t)timeline * 2 ? recs.size() / 2 : 0; // a step from the middle of the region
        for (size_t i = skip; i < recs.size() && i < skip + (size_t)timeline; i++) {
            float t0 = 0.0f, dt = 0.0f;
            cudaEventElapsedTime(&t0, recs[skip].a, recs[i].a);
            cudaEventElapsedTime(&dt, recs[i].a, recs[i].b);
            std::fprintf(stderr, "[ogn timeline] dev %d %-11s %s start %9.1f us  dur %7.1f us\n", device, names[recs[i].kind],
                         recs[i].onSide ? "side" : "main", t0 * 1e3f, dt * 1e3f);
        }
        timeline = 0;
    }
    for (auto &r : recs) {
        float ms = 0.0f;
        cudaCheck(cudaEventElapsedTime(&ms, r.a, r.b), "cudaEventElapsedTime");
        kernelMs[r.kind] += ms;
        timedLaunches[r.kind]++;
        eventPool.push_back(r.a); eventPool.push_back(r.b);
    }
    recs.clear();
}

void DeviceContext::profileReset() {
    profileCollect();
    for (int k = 0; k < kKernelKinds; k++) { launches[k] = 0; timedLaunches[k] = 0; kernelMs[k] = 0.0; }
    profileStep = 0;
}

// ---------------------------------------------------------------------------------------------------------------
void Geometry::build(const Int3 &hidden, const Int3 &visible, int r) {
    Hx = hidden.x; Hy = hidden.y; Hz = hidden.z; Vx = visible.x; Vy = visible.y; Vz = visible.z; radius = r;
    // project(): (int)(pos * (float)in / (float)out + 0.5f)   (Helpers.h:223-228, Helpers.cpp:245-246)
    const float sx = (float)Vx / (float)Hx, syf_ = (float)Vy / (float)Hy;
    auto axis = [r](int H, int V, float scale, std::vector<int> &lo, std::vector<int> &n, std::vector<int> &p, std::vector<int> &rlo,
                    std::vector<int> &rhi) {
        lo.resize(H); n.resize(H); p.resize(H);
        rlo.assign(V, H); rhi.assign(V, -1);
        int acc = 0;
        for (int o = 0; o < H; o++) {
            const int c = (int)(o * scale + 0.5f);
            const int a = std::max(0, c - r), b = std::min(V - 1, c + r);
            lo[o] = a; n[o] = std::max(0, b - a + 1); p[o] = acc; acc += n[o];
            for (int i = a; i <= b; i++) { rlo[i] = std::min(rlo[i], o); rhi[i] = std::max(rhi[i], o); }
        }
        return acc;
    };
    sxTotal = axis(Hx, Vx, sx, lox, nx, px, rlox, rhix);
    sy = axis(Hy, Vy, syf_, loy, ny, py, rloy, rhiy);
    nnz = (int64_t)Vz * Hz * (int64_t)sxTotal * sy;

    // packed layouts (include/ogn_kernels_api.h): F packs pf hidden columns along y over the union of their fields,
    // R groups the visible y of each field by pr
    pf = packFactor(Hz); pr = packFactor(Vz);
    const int npk = (Hy + pf - 1) / pf;
    loyf.assign(npk, 0); nyf.assign(npk, 0); pyf.assign(npk, 0);
    syf = 0;
    for (int q = 0; q < npk; q++) {
        int lo = Vy, hi = -1;
        for (int o = q * pf; o < std::min(Hy, (q + 1) * pf); o++)
            if (ny[o] > 0) { lo = std::min(lo, loy[o]); hi = std::max(hi, loy[o] + ny[o] - 1); }
        loyf[q] = hi >= lo ? lo : 0; nyf[q] = hi >= lo ? hi - lo + 1 : 0; pyf[q] = syf; syf += nyf[q];
    }
    ngr.assign(Hy, 0); pyr.assign(Hy, 0);
    syr = 0;
    for (int o = 0; o < Hy; o++) {
        ngr[o] = ny[o] > 0 ? (loy[o] + ny[o] - 1) / pr - loy[o] / pr + 1 : 0;
        pyr[o] = syr; syr += ngr[o];
    }

    std::vector<int> all;
    auto push = [&all](const std::vector<int> &v) { size_t o = all.size(); all.insert(all.end(), v.begin(), v.end()); return o; };
    const size_t o0 = push(lox), o1 = push(nx), o2 = push(px), o3 = push(loy), o4 = push(ny), o5 = push(py), o6 = push(rlox),
                 o7 = push(rhix), o8 = push(rloy), o9 = push(rhiy), o10 = push(loyf), o11 = push(nyf), o12 = push(pyf),
                 o13 = push(ngr), o14 = push(pyr);
    tables_.alloc(all.size());
    cudaCheck(cudaMemcpy(tables_.ptr(), all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice), "geometry H2D");
    const int *t = tables_.ptr();
    host.Hx = Hx; host.Hy = Hy; host.Hz = Hz; host.Vx = Vx; host.Vy = Vy; host.Vz = Vz; host.radius = r; host.sy = sy;
    host.pf = pf; host.pr = pr; host.syf = syf; host.syr = syr;
    host.lox = t + o0; host.nx = t + o1; host.px = t + o2; host.loy = t + o3; host.ny = t + o4; host.py = t + o5;
    host.rlox = t + o6; host.rhix = t + o7; host.rloy = t + o8; host.rhiy = t + o9;
    host.loyf = t + o10; host.nyf = t + o11; host.pyf = t + o12; host.ngr = t + o13; host.pyr = t + o14;
    host.max_nx = nx.empty() ? 0 : *std::max_element(nx.begin(), nx.end());
    host.max_nyf = nyf.empty() ? 0 : *std::max_element(nyf.begin(), nyf.end());
    int mcx = 0, mcy = 0;
    for (int i = 0; i < Vx; i++) mcx = std::max(mcx, rhix[i] - rlox[i] + 1);
    for (int i = 0; i < Vy; i++) mcy = std::max(mcy, rhiy[i] - rloy[i] + 1);
    host.max_cov = std::max(mcx, 0) * std::max(mcy, 0);
    bool tma = Vy % 4 == 0 && !nyf.empty();
    for (size_t q = 0; q < nyf.size() && tma; q++) {
        if (nyf[q] <= 0) tma = false;
        if (q > 0 && (loyf[q] < loyf[q - 1] || loyf[q] + nyf[q] < loyf[q - 1] + nyf[q - 1])) tma = false;
    }
    for (int o = 0; o < Hx && tma; o++)
        if (nx[o] <= 0) tma = false;
    host.tma_ok = tma ? 1 : 0;
}

std::shared_ptr<Geometry> Geometry::get(const Int3 &hidden, const Int3 &visible, int radius) {
    // one table set per (device, shape): layers with temporalHorizon > 1 share it
    typedef std::tuple<int, int, int, int, int, int, int, int> Key;
    static std::mutex mtx;
    static std::map<Key, std::weak_ptr<Geometry>> cache;
    int device = 0;
    cudaCheck(cudaGetDevice(&device), "cudaGetDevice");
    Key k(device, hidden.x, hidden.y, hidden.z, visible.x, visible.y, visible.z, radius);
    std::lock_guard<std::mutex> lock(mtx);
    if (auto sp = cache[k].lock()) return sp;
    auto sp = std::shared_ptr<Geometry>(new Geometry());
    sp->build(hidden, visible, radius);
    cache[k] = sp;
    return sp;
}

void Geometry::visibleRange(int a, int b, int &lo, int &hi) const {
    lo = Vx; hi = 0;
    for (int o = a; o < b; o++)
        if (nx[o] > 0) { lo = std::min(lo, lox[o]); hi = std::max(hi, lox[o] + nx[o]); }
    if (hi < lo) lo = hi = 0;
}
void Geometry::hiddenCover(int a, int b, int &lo, int &hi) const {
    lo = Hx; hi = 0;
    for (int v = a; v < b; v++)
        if (rhix[v] >= rlox[v]) { lo = std::min(lo, rlox[v]); hi = std::max(hi, rhix[v] + 1); }
    if (hi < lo) lo = hi = 0;
}

// ---------------------------------------------------------------------------------------------------------------
void WeightSet::allocate(DeviceContext &ctx, std::shared_ptr<Geometry> g, int batch_, bool hasR_, bool hasRecon, int a, int b) {
    geo = g; batch = batch_; hasR = hasR_; fx0 = a; fx1 = b;
    const bool whole = (a == 0 && b == g->Hx);
    if (whole) { vx0 = 0; vx1 = g->Vx; rx0 = 0; rx1 = g->Hx; }
    else {
        g->visibleRange(a, b, vx0, vx1);     // visible columns whose reconstruction touches an owned hidden column
        g->hiddenCover(vx0, vx1, rx0, rx1);  // every hidden column those reconstructions sum over (replicas)
    }
    wf.allocZero((size_t)(fCount() * batch), ctx.stream); // padding elements of the packed layouts stay zero
    if (hasR) wr.allocZero((size_t)(rCount() * batch), ctx.stream);
    if (hasRecon) recon.allocZero((size_t)g->Vx * g->Vy * g->Vz * batch, ctx.stream);
}

ogn_wset WeightSet::desc() const {
    ogn_wset d;
    std::memset(&d, 0, sizeof(d));
    d.wf = wf.ptr(); d.wr = wr.ptr();
    d.wf_base = geo->xBlockF(fx0); d.wr_base = hasR ? geo->xBlockR(rx0) : 0;
    d.wf_mstride = fCount(); d.wr_mstride = rCount();
    d.fx0 = fx0; d.fx1 = fx1; d.rx0 = rx0; d.rx1 = rx1;
    return d;
}

namespace {
const int64_t kStageFloats = 32ll << 20; // 128 MB staging chunks
}

void WeightSet::fillFromCSRStream(DeviceContext &ctx, int member, const std::function<void(float *, int64_t)> &fill) {
    const Geometry &g = *geo;
    const ogn_wset d = desc();
    const int ncolsAll = g.Hx * g.Hy;
    std::vector<float> host;
    DevBuf<float> stage;
    stage.alloc((size_t)std::min<int64_t>(kStageFloats, std::max<int64_t>(g.nnz, 1)));
    int col = 0;
    while (col < ncolsAll) {
        // largest run of whole columns that fits the staging buffer (at least one column)
        int end = col;
        const int64_t base = g.colBlock(col / g.Hy, col % g.Hy);
        int64_t n = 0;
        while (end < ncolsAll) {
            const int64_t blk = (int64_t)g.slots(end / g.Hy, end % g.Hy) * g.Vz * g.Hz;
            if (n + blk > (int64_t)stage.size() && end > col) break;
            n += blk; end++;
        }
        if (n > (int64_t)stage.size()) stage.alloc((size_t)n);
        host.resize((size_t)n);
        fill(host.data(), n);
        const int xa = col / g.Hy, xb = (end - 1) / g.Hy + 1; // hidden x touched by this chunk
        const bool needF = xa < fx1 && xb > fx0, needR = hasR && xa < rx1 && xb > rx0;
        if (needF || needR) {
            cudaCheck(cudaMemcpyAsync(stage.ptr(), host.data(), (size_t)n * sizeof(float), cudaMemcpyHostToDevice, ctx.stream), "H2D");
            if (needF) {
                const int c0 = std::max(col, fx0 * g.Hy), c1 = std::min(end, fx1 * g.Hy);
                kernelCheck(ogn_weights_from_csr(&d, &g.host, member, 0, c0, c1, stage.ptr(), base, ctx.stream), "weights_from_csr F");
            }
            if (needR) {
                const int c0 = std::max(col, rx0 * g.Hy), c1 = std::min(end, rx1 * g.Hy);
                kernelCheck(ogn_weights_from_csr(&d, &g.host, member, 1, c0, c1, stage.ptr(), base, ctx.stream), "weights_from_csr R");
            }
            ctx.sync(); // `host` and `stage` are reused by the next chunk
        }
        col = end;
    }
}

void WeightSet::uploadCSR(DeviceContext &ctx, int member, const float *csrAll) {
    int64_t pos = 0;
    fillFromCSRStream(ctx, member, [&](float *dst, int64_t n) { std::memcpy(dst, csrAll + pos, (size_t)n * sizeof(float)); pos += n; });
}

void WeightSet::downloadCSR(DeviceContext &ctx, int member, float *out) const {
    const Geometry &g = *geo;
    const ogn_wset d = desc();
    DevBuf<float> stage;
    stage.alloc((size_t)std::min<int64_t>(kStageFloats, std::max<int64_t>(fCount(), 1)));
    int col = fx0 * g.Hy;
    const int colEnd = fx1 * g.Hy;
    while (col < colEnd) {
        int end = col;
        const int64_t base = g.colBlock(col / g.Hy, col % g.Hy);
        int64_t n = 0;
        while (end < colEnd) {
            const int64_t blk = (int64_t)g.slots(end / g.Hy, end % g.Hy) * g.Vz * g.Hz;
            if (n + blk > (int64_t)stage.size() && end > col) break;
            n += blk; end++;
        }
        if (n > (int64_t)stage.size()) stage.alloc((size_t)n);
        kernelCheck(ogn_weights_to_csr(&d, &g.host, member, 0, col, end, stage.ptr(), base, ctx.stream), "weights_to_csr");
        cudaCheck(cudaMemcpyAsync(out + base, stage.ptr(), (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx.stream), "D2H");
        ctx.sync();
        col = end;
    }
}

void WeightSet::initHash(DeviceContext &ctx, int member, uint64_t seed, uint32_t streamId, float lo, float hi) {
    const ogn_wset d = desc();
    kernelCheck(ogn_weights_init_hash(&d, &geo->host, member, seed, streamId, lo, hi, ctx.stream), "weights_init_hash");
}

void WeightSet::initialise(DeviceContext &ctx, int member, const WeightInit &wi) {
    if (wi.cs->deviceInit) {
        initHash(ctx, member, wi.cs->deviceInitSeed, wi.streamId, wi.lo, wi.hi);
        return;
    }
    std::uniform_real_distribution<float> dist(wi.lo, wi.hi);
    std::mt19937 &rng = wi.cs->rng;
    fillFromCSRStream(ctx, member, [&](float *dst, int64_t n) {
        for (int64_t i = 0; i < n; i++) dst[i] = dist(rng);
    });
}

void WeightSet::cloneFrom(DeviceContext &ctx, const WeightSet &o) {
    geo = o.geo; batch = o.batch; hasR = o.hasR;
    fx0 = o.fx0; fx1 = o.fx1; rx0 = o.rx0; rx1 = o.rx1; vx0 = o.vx0; vx1 = o.vx1;
    wf.cloneFrom(o.wf, ctx.stream); wr.cloneFrom(o.wr, ctx.stream); recon.cloneFrom(o.recon, ctx.stream);
}

void WeightSetArray::cloneFrom(DeviceContext &ctx, const WeightSetArray &o) {
    sets.clear();
    sets.resize(o.sets.size());
    for (size_t i = 0; i < sets.size(); i++) sets[i].cloneFrom(ctx, o.sets[i]);
}

void writeDeviceRaw(std::ostream &os, DeviceContext &ctx, const void *dev, size_t bytes) {
    const uint64_t n = bytes;
    os.write(reinterpret_cast<const char *>(&n), sizeof(n));
    if (!bytes) return;
    ctx.joinSide();
    const size_t chunk = std::min<size_t>(bytes, (size_t)64 << 20);
    PinnedBuf<char> stage;
    stage.alloc(chunk);
    for (size_t off = 0; off < bytes; off += chunk) {
        const size_t m = std::min(chunk, bytes - off);
        cudaCheck(cudaMemcpyAsync(stage.ptr(), (const char *)dev + off, m, cudaMemcpyDeviceToHost, ctx.stream), "checkpoint D2H");
        ctx.sync();
        os.write(stage.ptr(), (std::streamsize)m);
    }
    if (!os) throw std::runtime_error("ogmaneo-b200: checkpoint write failed");
}

void readDeviceRaw(std::istream &is, DeviceContext &ctx, void *dev, size_t bytes) {
    uint64_t n = 0;
    is.read(reinterpret_cast<char *>(&n), sizeof(n));
    if (!is || n != bytes)
        throw std::runtime_error("ogmaneo-b200: checkpoint does not match this hierarchy (buffer of " + std::to_string(n) +
                                 " bytes where " + std::to_string(bytes) + " are expected)");
    if (!bytes) return;
    ctx.joinSide();
    const size_t chunk = std::min<size_t>(bytes, (size_t)64 << 20);
    PinnedBuf<char> stage;
    stage.alloc(chunk);
    for (size_t off = 0; off < bytes; off += chunk) {
        const size_t m = std::min(chunk, bytes - off);
        is.read(stage.ptr(), (std::streamsize)m);
        if (!is) throw std::runtime_error("ogmaneo-b200: checkpoint is truncated");
        cudaCheck(cudaMemcpyAsync((char *)dev + off, stage.ptr(), m, cudaMemcpyHostToDevice, ctx.stream), "checkpoint H2D");
        ctx.sync();
    }
}

int geoIndex(ogn_rf *table, int &count, const Geometry &g) {
    for (int i = 0; i < count; i++)
        if (table[i].lox == g.host.lox) return i; // same device tables <=> same Geometry object
    if (count >= OGN_MAX_GEO) throw std::runtime_error("ogmaneo-b200: more than 4 distinct visible-layer shapes on one layer");
    table[count] = g.host;
    return count++;
}

// ---------------------------------------------------------------------------------------------------------------
void consumeKernel1(ComputeSystem &cs, int size) {
    if (!cs.replayRng) return;
    std::uniform_int_distribution<int> seedDist(0, 999999);
    const int batches = (size + cs.batchSize1 - 1) / cs.batchSize1;
    for (int i = 0; i < batches; i++) (void)seedDist(cs.rng);
}

void consumeKernel2(ComputeSystem &cs, int X, int Y, std::vector<int> *seeds) {
    if (!cs.replayRng && !seeds) return;
    std::uniform_int_distribution<int> seedDist(0, 999999);
    const int bx = (X + cs.batchSize2.x - 1) / cs.batchSize2.x, by = (Y + cs.batchSize2.y - 1) / cs.batchSize2.y;
    if (seeds) seeds->resize((size_t)bx * by);
    for (int i = 0; i < bx * by; i++) {
        const int s = seedDist(cs.rng);
        if (seeds) (*seeds)[i] = s;
    }
}

void slabOf(int X, int rank, int world, int &a, int &b) {
    if (world <= 1) { a = 0; b = X; return; }
    if (X % world != 0)
        throw std::runtime_error("ogmaneo-b200: sharding needs every layer's x size to be a multiple of the device count");
    a = rank * (X / world); b = a + X / world;
}

} // namespace detail

// ---------------------------------------------------------------------------------------------------------------
namespace {
int g_numThreads = 1;
}
void ComputeSystem::setNumThreads(int numThreads) { g_numThreads = numThreads; } // host threads are irrelevant on the device path
int ComputeSystem::getNumThreads() { return g_numThreads; }

detail::DeviceContext &ComputeSystem::context() {
    if (!ctx) {
        ctx = std::make_shared<detail::DeviceContext>(device);
        // the side stream pays off when a step is short and has exchanges to hide (sharded layers); on one GPU the kernel
        // chains are bandwidth-bound either way and a single stream keeps the per-kernel timings undisturbed
        if (!std::getenv("OGN_SIDE_STREAM")) ctx->useSide = shardWorld > 1;
        if (peerExchange && shardWorld > 1) {
            const char *e = std::getenv("OGN_PEER_ARENA_MB");
            ctx->arenaCreate((size_t)(e ? std::atoll(e) : 64) << 20, shardRank, shardWorld);
        }
    } else detail::cudaCheck(cudaSetDevice(ctx->device), "cudaSetDevice");
    return *ctx;
}
void ComputeSystem::synchronize() { context().sync(); context().peerCheck(); }
void ComputeSystem::peerExport(void *handle64) { context().arenaExport(handle64); }
void ComputeSystem::peerAttach(const void *handles64, int world) { context().arenaAttach(handles64, world); }
void *ComputeSystem::stream() { return (void *)context().stream; }

} // namespace ogmaneo

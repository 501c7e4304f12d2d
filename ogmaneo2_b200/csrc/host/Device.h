// csrc/host/Device.h — internal plumbing shared by the host classes: CUDA context, device/pinned buffers, the
// closed-form receptive-field geometry and the device-resident weight sets. Not installed; users see include/ogmaneo.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <functional>
#include <istream>
#include <ostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <ogmaneo/ComputeSystem.h>

#include "../../../include/ogn_kernels_api.h"

namespace ogmaneo {
namespace detail {

void cudaCheck(cudaError_t e, const char *what);
inline void kernelCheck(int code, const char *what) { cudaCheck((cudaError_t)code, what); }

enum KernelKind { kScForward = 0, kScLearn = 1, kPredStep = 2, kActorForward = 3, kActorLearn = 4, kScUpdate = 5, kHierSmall = 6, kExchange = 7, kKernelKinds = 8 };

// Phases of one Hierarchy::step recorded instead of launched (DeviceContext::collect): the single-launch small-hierarchy
// kernel runs them from a device-resident array (ogn_hier_step_small). layer / role let Hierarchy::step order them into
// barrier levels.
struct SmallPlan {
    struct Node { ogn_small_phase ph; int layer; };
    std::vector<Node> nodes;
    int curLayer = 0;
    int lists = 0;      // work lists handed out so far
    bool ok = true;     // false: some phase has no 128-bit path (the plan must not be launched)
    void add(const ogn_small_phase &p) { nodes.push_back(Node{p, curLayer}); }
};

// Peer-memory exchange of sharded CSDRs (one process per GPU on one NVLink / NVSwitch node). Every CSDR buffer a
// sharded layer exchanges is carved out of ONE arena per device, in construction order, so it sits at the same offset on
// every rank; the arena is exported with CUDA IPC and each rank maps its peers' arenas. An all-gather is then one kernel
// (ogn_peer_allgather): every rank stores its slab straight into the other ranks' arenas over NVLink, raises a flag in
// each of them and waits for theirs. No NCCL call, no host round trip.
struct PeerArena {
    static constexpr size_t kHeaderInts = 64; // [0..15] arrival flags, [16] blocks done, [18..19] ns waited for peers (total), [20..21] scratch
    int *base = nullptr;
    size_t capInts = 0, usedInts = kHeaderInts;
    int rank = 0, world = 1;
    int *peer[OGN_MAX_PEERS] = {};
    bool attached = false;
    unsigned epoch = 0;
    unsigned *errHost = nullptr, *errDev = nullptr; // mapped pinned word: set by a kernel whose wait timed out
};

// One device + one stream. Created lazily by ComputeSystem::context(); throws if no sm_100 device is usable.
struct DeviceContext {
    int device;
    cudaStream_t stream;
    explicit DeviceContext(int device);
    ~DeviceContext();
    std::unique_ptr<PeerArena> arena;
    void arenaCreate(size_t bytes, int rank, int world);
    int *arenaAlloc(size_t nInts);                       // zero-initialised, 256-byte aligned
    bool inArena(const int *p) const { return arena && p >= arena->base && p < arena->base + arena->capInts; }
    void arenaExport(void *handle64) const;
    void arenaAttach(const void *handles64, int world);
    // in-place all-gather of [world][countPerRank] ints on `stream`; deferred = queue it on the side stream instead (the
    // step's final output: nothing on `stream` needs it before the next exchange, which joins it)
    // waitMask: ranks whose slabs the kernels queued next read (0 = everybody); a partial mask requires that an exchange
    // with the full mask follows before anything reads the other slabs
    void peerAllGather(int *buffer, int countPerRank, bool deferred = false, unsigned waitMask = 0);
    void peerCheck() const;                              // throws if a peer wait timed out
    unsigned long long peerTakeWaitNs();                 // device time spent waiting for other ranks since the last call (syncs)
    // Side stream: SparseCoder learn (reconstruction + update) of a step runs here, concurrently with the Predictors on
    // `stream` (they touch disjoint weights; the two bandwidth-bound kernel chains fill each other's tails), and so does
    // the deferred exchange of the step's output predictions. forkSide() makes the side stream wait for everything
    // queued on `stream` so far; joinLearn()/joinExchange() make `stream` wait for the side work recorded last.
    cudaStream_t side;
    cudaEvent_t evFork, evLearn, evExchange, evPredLearn;
    bool learnPending = false, exchangePending = false, useSide = true;
    bool sideForced = false; // OGN_SIDE_STREAM was given: Hierarchy does not choose
    // graphs: Hierarchy::step may capture its kernel sequence into CUDA graphs (OGN_GRAPHS=0 disables). dry: the step's
    // host code runs for its state transitions only, nothing is queued (the captured graph is launched instead).
    bool graphs = true, dry = false, dryUncounted = false;
    // non-null: layers append the phases they would launch to this plan (and launch nothing), see SmallPlan
    SmallPlan *collect = nullptr;
    bool small = false;     // OGN_SMALL_KERNEL=1: small hierarchies step through the single-launch kernel (ogn_hier_step_small)
    int smallMaxUnits = 8192;
    void forkSide();
    void learnQueued();     // call after queueing learn work on `side`
    void exchangeQueued();  // call after queueing a deferred exchange on `side`
    void joinLearn();
    void joinExchange();
    void joinSide() { joinLearn(); joinExchange(); }
    void sync() const {
        cudaCheck(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
        cudaCheck(cudaStreamSynchronize(side), "cudaStreamSynchronize(side)");
    }

    // Launch accounting: counts are always kept; with profiling on, every launch is bracketed by two CUDA events on
    // `stream` and profileCollect() (after a sync) turns them into per-kind device time.
    bool profiling = false;
    // Bracketing every launch with events costs stream time (about 4 us per kernel on a sharded 0.3 ms step), so only
    // every profilePeriod-th Hierarchy::step is timed; timedLaunches counts the launches kernelMs sums over.
    int profilePeriod = 1;
    long long profileStep = 0;
    bool sampleThis = true;
    void beginStep() { sampleThis = profiling && (profileStep++ % (profilePeriod > 0 ? profilePeriod : 1)) == 0; }
    long long timedLaunches[kKernelKinds] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long launches[kKernelKinds] = {0, 0, 0, 0, 0, 0, 0, 0};
    double kernelMs[kKernelKinds] = {0, 0, 0, 0, 0, 0, 0, 0};
    struct Rec { cudaEvent_t a, b; int kind; int onSide; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> eventPool;
    cudaEvent_t takeEvent();
    void profileCollect();
    void profileReset();
};

// Brackets one kernel launch (see DeviceContext::profiling).
struct LaunchScope {
    DeviceContext &c;
    int kind;
    cudaStream_t st;
    cudaEvent_t a = nullptr;
    LaunchScope(DeviceContext &ctx, int kind_, cudaStream_t stream = nullptr) : c(ctx), kind(kind_), st(stream ? stream : ctx.stream) {
        if (c.dry && c.dryUncounted) return; // the single-launch step replays host state only: nothing of this kind runs
        c.launches[kind]++;
        if (c.profiling && c.sampleThis) { a = c.takeEvent(); cudaCheck(cudaEventRecord(a, st), "cudaEventRecord"); }
    }
    ~LaunchScope() {
        if (a) {
            cudaEvent_t b = c.takeEvent();
            cudaEventRecord(b, st);
            c.recs.push_back(DeviceContext::Rec{a, b, kind, st == c.side ? 1 : 0});
        }
    }
};

template <typename T> class DevBuf {
public:
    DevBuf() {}
    ~DevBuf() { release(); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p_(o.p_), n_(o.n_), owned_(o.owned_) { o.p_ = nullptr; o.n_ = 0; o.owned_ = true; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { release(); p_ = o.p_; n_ = o.n_; owned_ = o.owned_; o.p_ = nullptr; o.n_ = 0; o.owned_ = true; }
        return *this;
    }
    void alloc(size_t n) {
        release();
        n_ = n;
        if (n) cudaCheck(cudaMalloc((void **)&p_, n * sizeof(T)), "cudaMalloc");
    }
    void borrow(T *p, size_t n) { release(); p_ = p; n_ = n; owned_ = false; } // memory owned by someone else (peer arena)
    void allocZero(size_t n, cudaStream_t s) {
        alloc(n);
        if (n) cudaCheck(cudaMemsetAsync(p_, 0, n * sizeof(T), s), "cudaMemsetAsync");
    }
    void release() {
        if (p_ && owned_) cudaFree(p_);
        p_ = nullptr; n_ = 0; owned_ = true;
    }
    void cloneFrom(const DevBuf &o, cudaStream_t s) {
        alloc(o.n_);
        if (n_) cudaCheck(cudaMemcpyAsync(p_, o.p_, n_ * sizeof(T), cudaMemcpyDeviceToDevice, s), "clone D2D");
    }
    void upload(const T *h, size_t n, cudaStream_t s, size_t offset = 0) {
        if (n) cudaCheck(cudaMemcpyAsync(p_ + offset, h, n * sizeof(T), cudaMemcpyHostToDevice, s), "H2D");
    }
    void download(T *h, size_t n, cudaStream_t s, size_t offset = 0) const {
        if (n) cudaCheck(cudaMemcpyAsync(h, p_ + offset, n * sizeof(T), cudaMemcpyDeviceToHost, s), "D2H");
    }
    T *ptr() const { return p_; }
    size_t size() const { return n_; }

private:
    T *p_ = nullptr;
    size_t n_ = 0;
    bool owned_ = true;
};

template <typename T> class PinnedBuf {
public:
    PinnedBuf() {}
    ~PinnedBuf() { release(); }
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    void alloc(size_t n) {
        release();
        n_ = n;
        if (n) cudaCheck(cudaMallocHost((void **)&p_, n * sizeof(T)), "cudaMallocHost");
    }
    void release() {
        if (p_) cudaFreeHost(p_);
        p_ = nullptr; n_ = 0;
    }
    T *ptr() const { return p_; }
    size_t size() const { return n_; }

private:
    T *p_ = nullptr;
    size_t n_ = 0;
};

// initSMLocalRF (Helpers.cpp:236-313) in closed form; host tables + their device copies.
struct Geometry {
    int Hx = 0, Hy = 0, Hz = 0, Vx = 0, Vy = 0, Vz = 0, radius = 0, sy = 0;
    int pf = 1, pr = 1, syf = 0, syr = 0, sxTotal = 0;
    std::vector<int> lox, nx, px, loy, ny, py, rlox, rhix, rloy, rhiy;
    std::vector<int> loyf, nyf, pyf; // per hidden-y pack (F layout)
    std::vector<int> ngr, pyr;       // per hidden y (R layout)
    int64_t nnz = 0;                 // weights (CSR order, == the reference's nonZeroValues.size())
    ogn_rf host{};                   // the by-value kernel descriptor; its table pointers are device pointers

    static std::shared_ptr<Geometry> get(const Int3 &hidden, const Int3 &visible, int radius);
    // pack factor of a layout whose rows hold z cells: must match Pack<G>::P in ogn_kernels.cu
    static int packFactor(int z) {
        int g = 1;
        while (g < z && g < 32) g <<= 1;
        return g >= 8 ? 32 / g : 4;
    }
    // CSR-order block offsets (reference value order)
    int64_t colBlock(int ox, int oy) const { return (int64_t)Vz * Hz * ((int64_t)px[ox] * sy + (int64_t)nx[ox] * py[oy]); }
    int64_t xBlock(int ox) const { return ox >= Hx ? nnz : colBlock(ox, 0); } // offset of column (ox, 0)
    // device layouts: offset of the first element of hidden x = ox (ox == Hx: total size incl. padding)
    int64_t xBlockF(int ox) const { return (int64_t)Vz * Hz * pf * (int64_t)(ox >= Hx ? sxTotal : px[ox]) * syf; }
    int64_t xBlockR(int ox) const { return (int64_t)Vz * Hz * pr * (int64_t)(ox >= Hx ? sxTotal : px[ox]) * syr; }
    int slots(int ox, int oy) const { return nx[ox] * ny[oy]; }
    // visible x range [lo, hi) touched by hidden x in [a, b); hidden x range [lo, hi) covering visible x in [a, b)
    void visibleRange(int a, int b, int &lo, int &hi) const;
    void hiddenCover(int a, int b, int &lo, int &hi) const;

private:
    DevBuf<int> tables_;
    void build(const Int3 &hidden, const Int3 &visible, int radius);
};

// How the weights of a new layer get their initial values.
struct WeightInit {
    ComputeSystem *cs;   // mt19937 order (reference parity) when cs->deviceInit == false
    float lo, hi;        // uniform range
    uint32_t streamId;   // distinguishes matrices in the counter-hash mode
};

// Weights of one visible layer of one layer object, resident in HBM.
struct WeightSet {
    std::shared_ptr<Geometry> geo;
    int batch = 1;
    bool hasR = false;
    int fx0 = 0, fx1 = 0, rx0 = 0, rx1 = 0, vx0 = 0, vx1 = 0;
    DevBuf<float> wf, wr, recon;

    void allocate(DeviceContext &ctx, std::shared_ptr<Geometry> g, int batch, bool hasR, bool hasRecon, int fx0, int fx1);
    ogn_wset desc() const;
    // "virtual bases": storage pointer minus the block offset of the first stored column (see ogn_kernels_api.h)
    float *fBase() const { return wf.ptr() - geo->xBlockF(fx0); }
    float *rBase() const { return hasR ? wr.ptr() - geo->xBlockR(rx0) : nullptr; }
    int64_t fCount() const { return geo->xBlockF(fx1) - geo->xBlockF(fx0); } // floats stored per member (incl. padding)
    int64_t rCount() const { return hasR ? geo->xBlockR(rx1) - geo->xBlockR(rx0) : 0; }
    // fill(dst, n) must produce the next n values of the matrix in the reference's CSR value order (whole matrix,
    // from index 0 to nnz). Values of columns this device does not store are drawn and dropped.
    void fillFromCSRStream(DeviceContext &ctx, int member, const std::function<void(float *, int64_t)> &fill);
    void uploadCSR(DeviceContext &ctx, int member, const float *csrAll);
    // Values of the stored F range written at their global CSR positions in out (size nnz); other entries untouched.
    void downloadCSR(DeviceContext &ctx, int member, float *out) const;
    void initHash(DeviceContext &ctx, int member, uint64_t seed, uint32_t streamId, float lo, float hi);
    void initialise(DeviceContext &ctx, int member, const WeightInit &wi);
    void cloneFrom(DeviceContext &ctx, const WeightSet &o);
};

struct WeightSetArray {
    std::vector<WeightSet> sets;
    void cloneFrom(DeviceContext &ctx, const WeightSetArray &o);
};

// Raw device buffer <-> stream, 64-bit element count first, through a pinned staging buffer (checkpoints).
void writeDeviceRaw(std::ostream &os, DeviceContext &ctx, const void *dev, size_t bytes);
void readDeviceRaw(std::istream &is, DeviceContext &ctx, void *dev, size_t bytes);
template <typename T> void writeDev(std::ostream &os, DeviceContext &ctx, const DevBuf<T> &b) {
    writeDeviceRaw(os, ctx, b.ptr(), b.size() * sizeof(T));
}
template <typename T> void readDev(std::istream &is, DeviceContext &ctx, DevBuf<T> &b) {
    readDeviceRaw(is, ctx, b.ptr(), b.size() * sizeof(T));
}

// Geometry table of one launch: returns the index of g in `table` (by-value ogn_rf copies), appending if new.
int geoIndex(ogn_rf *table, int &count, const Geometry &g);

// RNG consumption of the reference's launchers (Helpers.cpp:15-72). seeds != nullptr forces the draws.
void consumeKernel1(ComputeSystem &cs, int size);
void consumeKernel2(ComputeSystem &cs, int X, int Y, std::vector<int> *seeds = nullptr);

// x-slab owned by `rank` of `world` for a grid with X rows (requires X % world == 0 when world > 1)
void slabOf(int X, int rank, int world, int &a, int &b);

} // namespace detail
} // namespace ogmaneo

// include/ogmaneo/ComputeSystem.h — B200-native drop-in for the reference's <ogmaneo/ComputeSystem.h> (:17-41).
// Same public members (batchSize1/2/3, rng, setNumThreads/getNumThreads). In addition it owns the CUDA context the
// layers run on: one device, one stream. There is no CPU fallback: creating the context throws std::runtime_error
// if no sm_100 device is usable.
#pragma once

#include "Helpers.h"

#include <cstdint>
#include <memory>
#include <random>

namespace ogmaneo {
namespace detail {
struct DeviceContext;
}

class ComputeSystem {
public:
    // Tile sizes of the reference's runKernel1/2/3. They no longer shape the execution (CUDA grids do), but they
    // still define how many seeds each "kernel" draws from rng, which the Actor's sampling depends on.
    int batchSize1;
    Int2 batchSize2;
    Int3 batchSize3;

    std::mt19937 rng;

    // The reference draws one seed per tile from `rng` for every runKernelN launch (Helpers.cpp:22-31,46-61), although
    // only Actor::forward ever uses the per-tile generators. When replayRng is true every layer replays those draws on
    // the host so that cs.rng (and therefore Actor sampling) stays bit-identical to the reference run with one thread.
    // When false the draws of layers that ignore their generators are skipped: cs.rng then diverges from the
    // reference's, which no output can observe unless an Actor exists. Default false;
    // Hierarchy::initRandom switches it on when any input is an action input. A 512x512 layer would otherwise spend
    // more host time drawing seeds than the GPU spends on the step.
    bool replayRng;

    // CUDA device ordinal to run on (-1: the current device).
    int device;

    // Weight initialisation of initRandom(): false = draw every weight from `rng` in the reference's order (bit-identical
    // initial weights, ~4 ns per weight on the host); true = counter-based hash on the device keyed by
    // (deviceInitSeed, matrix, CSR index), for sizes the reference cannot construct.
    bool deviceInit;
    uint64_t deviceInitSeed;
    uint32_t deviceInitMatrix; // running matrix counter of the hash mode (identical on every rank)

    // Spatial sharding of every layer by hidden-column x-slab over shardWorld devices (one process per device). After
    // a layer produced its slab of a CSDR, allGather(user, buffer, countPerRank, stream) must make `buffer`
    // ([shardWorld][countPerRank] int32, this rank's slab already in place) complete on every rank, ordered on `stream`
    // (e.g. ncclAllGather in place). Weights never move. shardWorld == 1: no exchange.
    typedef void (*AllGatherFn)(void *user, int *buffer, int countPerRank, void *stream);
    int shardRank, shardWorld;
    AllGatherFn allGather;
    void *allGatherUser;
    // Alternative to the callback for one process per GPU on one NVLink node: set peerExchange before the first layer is
    // created, exchange the 64-byte handles of peerExport() between the ranks (any transport) and hand all of them to
    // peerAttach(). The all-gathers then run as one peer-memory kernel each on this system's stream (no NCCL, no host).
    bool peerExchange;
    void peerExport(void *handle64);
    void peerAttach(const void *handles64, int world);

    ComputeSystem()
        : batchSize1(512), batchSize2(2, 2), batchSize3(2, 2, 2), replayRng(false), device(-1), deviceInit(false),
          deviceInitSeed(0), deviceInitMatrix(0), shardRank(0), shardWorld(1), allGather(nullptr), allGatherUser(nullptr),
          peerExchange(false) {}

    static void setNumThreads(int numThreads);
    static int getNumThreads();

    // Internal: lazily created device context (device ordinal + stream), shared with the layers created through it.
    detail::DeviceContext &context();
    std::shared_ptr<detail::DeviceContext> sharedContext() { context(); return ctx; }
    // Run on `other`'s device and stream (call before the first layer is created through this system): objects made through
    // two systems that share a context queue their kernels in program order on one stream, so one can consume the device
    // buffers the other just produced without any synchronisation (ImageEncoder -> Hierarchy), while each keeps its own rng.
    void shareContextWith(ComputeSystem &other) { ctx = other.sharedContext(); device = other.device; }
    // Block until all work queued on this system's stream has finished.
    void synchronize();
    // The cudaStream_t (as void*) the layers launch on.
    void *stream();

private:
    std::shared_ptr<detail::DeviceContext> ctx;
};
} // namespace ogmaneo

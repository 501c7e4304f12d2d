// csrc/host/LayerImpl.h — device-side state behind the public layer classes (pimpl bodies). Internal.
#pragma once

#include <ogmaneo/Hierarchy.h>

#include "Device.h"

namespace ogmaneo {

struct SparseCoder::Impl {
    std::shared_ptr<detail::DeviceContext> ctx;
    int batch = 1;          // ensemble members stepped in lockstep
    int x0 = 0, x1 = 0;     // hidden x-slab owned by this device
    int world = 1, rank = 0;
    detail::WeightSetArray ws;
    // hiddenCs / hiddenCsPrev ping-pong: cs_[cur] is hiddenCs; hiddenCsPrev is the same buffer after a step (the
    // reference copies hiddenCs -> hiddenCsPrev at the end of step()) or the other one after setState().
    detail::DevBuf<int> cs_[2];
    int cur = 0;
    bool prevAliasesCur = true;
    detail::DevBuf<unsigned long long> upd;
    // work list of the two-kernel learn: visible columns whose reconstruction missed (8-byte entries) + {entries, warps done}
    detail::DevBuf<unsigned long long> work;
    detail::DevBuf<unsigned> workCounters;
    // one list region per run of equally-shaped visible layers: run r owns entries [batch * runOffset[r], batch *
    // runOffset[r + 1]) of `work`, member m of an ensemble the m-th slice of it (single-launch kernel)
    std::vector<size_t> runOffset;
    std::vector<detail::DevBuf<int>> inStage; // standalone step(): uploaded inputs
    // host-mirror validity
    bool csDirty = true, reconDirty = true;
    std::vector<char> wDirty;

    void allocate(ComputeSystem &cs, const Int3 &hidden, const std::vector<VisibleLayerDesc> &vlds, int batch);
    void refreshCs(IntBuffer &hc, IntBuffer &hcPrev);
};

struct Predictor::Impl {
    std::shared_ptr<detail::DeviceContext> ctx;
    int batch = 1;
    int x0 = 0, x1 = 0;
    int world = 1;
    detail::WeightSetArray ws;
    detail::DevBuf<float> acts;
    detail::DevBuf<int> hiddenCs;
    // inputCsPrev ping-pong per visible layer: prev_[p][v] with p = prevCur is what the next learn() reads
    std::vector<detail::DevBuf<int>> prev_[2];
    int prevCur = 0;
    std::vector<detail::DevBuf<int>> inStage;
    detail::DevBuf<int> targetStage;
    detail::PinnedBuf<int> csPinned; // staging of hiddenCs downloads (pageable destinations are several times slower)
    bool csDirty = true, actDirty = true, prevDirty = true;
    std::vector<char> wDirty;

    void allocate(ComputeSystem &cs, const Int3 &hidden, const std::vector<VisibleLayerDesc> &vlds, int batch);
};

struct Actor::Impl {
    std::shared_ptr<detail::DeviceContext> ctx;
    detail::WeightSetArray vw, aw; // value (Hz = 1) and action weights
    detail::DevBuf<float> acts, values, u01;
    detail::DevBuf<int> hiddenCs;
    detail::PinnedBuf<float> u01Host;
    // history ring (CircleBuffer<HistorySample>, Actor.h:63): slot-major device buffers + host rewards
    int capacity = 0, start = 0, historySize = 0;
    std::vector<detail::DevBuf<int>> histInput; // [vl] -> [capacity][visible columns]
    detail::DevBuf<int> histTarget;             // [capacity][hidden columns]
    detail::DevBuf<float> histValues;           // [capacity][hidden columns]
    std::vector<float> rewards;                 // [capacity]
    std::vector<detail::DevBuf<int>> inStage;
    detail::DevBuf<int> targetStage;
    static constexpr int kU01Slots = 4;         // pinned u01 staging ring: the host may queue this many steps ahead
    cudaEvent_t u01Free[kU01Slots] = {};        // slot k may be rewritten after u01Free[k]
    int u01Slot = 0;
    bool csDirty = true, valDirty = true;
    std::vector<char> wDirty;

    ~Impl() {
        for (auto e : u01Free)
            if (e) cudaEventDestroy(e);
    }
    int slot(int logical) const { return (start + logical) % capacity; }
    void allocate(ComputeSystem &cs, const Int3 &hidden, int historyCapacity, const std::vector<VisibleLayerDesc> &vlds);
};

} // namespace ogmaneo

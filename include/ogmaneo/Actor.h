// include/ogmaneo/Actor.h — B200-native drop-in for the reference's <ogmaneo/Actor.h> (:15-181).
// Same public interface. Value/action weights, the history ring and all per-column state live in HBM; step()
// launches K9 (forward + sampling) and K11 (learn) from include/ogn_kernels_api.h. Sampling is bit-identical to the
// reference run with one thread: the per-tile std::mt19937 sub-generators are replayed on the host with the real
// libstdc++ classes and one canonical float per action column is handed to the kernel.
#pragma once

#include "ComputeSystem.h"

namespace ogmaneo {
class Actor {
public:
    struct VisibleLayerDesc {
        Int3 size;
        int radius;
        VisibleLayerDesc() : size(4, 4, 16), radius(2) {}
    };

    struct VisibleLayer {
        SparseMatrix valueWeights;  // host mirror (on demand)
        SparseMatrix actionWeights; // host mirror (on demand)
    };

    struct HistorySample { // host mirror form of one history entry (used by save/load)
        std::vector<IntBuffer> inputCs;
        IntBuffer hiddenTargetCsPrev;
        FloatBuffer hiddenValuesPrev;
        float reward;
    };

    struct Impl;

    float alpha;      // value learning rate
    float beta;       // action learning rate
    float gamma;      // discount
    int minSteps;     // minimum distance into the past of a learning sample
    int historyIters; // learning iterations per step

    Actor();
    Actor(const Actor &other);
    Actor &operator=(const Actor &other);
    ~Actor();

    // Actor.cpp:187-246
    void initRandom(ComputeSystem &cs, const Int3 &hiddenSize, int historyCapacity,
                    const std::vector<VisibleLayerDesc> &visibleLayerDescs);
    // Actor.cpp:248-313
    void step(ComputeSystem &cs, const std::vector<const IntBuffer *> &inputCs, const IntBuffer *hiddenTargetCsPrev,
              float reward, bool learnEnabled, bool mimic);

    void writeToStream(std::ostream &os) const; // Actor.cpp:315-369
    void readFromStream(std::istream &is);      // Actor.cpp:371-438

    int getNumVisibleLayers() const { return (int)visibleLayerDescs.size(); }
    const VisibleLayer &getVisibleLayer(int i) const;
    const VisibleLayerDesc &getVisibleLayerDesc(int i) const { return visibleLayerDescs[i]; }
    const IntBuffer &getHiddenCs() const;
    const Int3 &getHiddenSize() const { return hiddenSize; }

    // --- extensions ---
    const FloatBuffer &getHiddenValues() const; // private in the reference (Actor.h:60)
    void getWeights(int i, int which /*0 value, 1 action*/, std::vector<float> &out) const;
    // step() on device-resident CSDRs. Asynchronous on cs' stream.
    void stepDevice(ComputeSystem &cs, const int *const *inputCsDevice, const int *hiddenTargetCsPrevDevice, float reward,
                    bool learnEnabled, bool mimic);
    const int *hiddenCsDevice() const;
    void readFromStream(std::istream &is, ComputeSystem &cs);
    // raw device state (weights in their device layout, history ring, hyper-parameters) of an already constructed layer of
    // the same shape: the Actor part of Hierarchy::writeCheckpoint / readCheckpoint (Actor.cpp:315-438 is the reference's
    // own save format, which writeToStream keeps)
    void writeDeviceState(std::ostream &os) const;
    void readDeviceState(std::istream &is);

private:
    Int3 hiddenSize;
    std::vector<VisibleLayerDesc> visibleLayerDescs;
    mutable IntBuffer hiddenCs;
    mutable FloatBuffer hiddenValues;
    mutable std::vector<VisibleLayer> visibleLayers;
    Impl *impl;

    friend class Hierarchy;
};
} // namespace ogmaneo

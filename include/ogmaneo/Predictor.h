// include/ogmaneo/Predictor.h — B200-native drop-in for the reference's <ogmaneo/Predictor.h> (:15-150).
// Same public interface; state and weights live in HBM; learn()/activate() launch the fused sm_100a kernel
// K6+K7+K8 (include/ogn_kernels_api.h). Host members returned by getters are lazily refreshed mirrors.
#pragma once

#include <istream>
#include <ostream>

#include "ComputeSystem.h"

namespace ogmaneo {
class Predictor {
public:
    struct VisibleLayerDesc {
        Int3 size;
        int radius;
        VisibleLayerDesc() : size(4, 4, 16), radius(2) {}
    };

    struct VisibleLayer {
        SparseMatrix weights;  // host mirror (on demand)
        IntBuffer inputCsPrev; // host mirror: the inputs seen by the previous activate()
    };

    struct Impl;

    float alpha; // learning rate

    Predictor();
    Predictor(const Predictor &other);
    Predictor &operator=(const Predictor &other);
    ~Predictor();

    // Predictor.cpp:72-109 — weights U(-0.01,0.01)
    void initRandom(ComputeSystem &cs, const Int3 &hiddenSize, const std::vector<VisibleLayerDesc> &visibleLayerDescs);
    // Predictor.cpp:111-127
    void activate(ComputeSystem &cs, const std::vector<const IntBuffer *> &inputCs);
    // Predictor.cpp:129-135
    void learn(ComputeSystem &cs, const IntBuffer *hiddenTargetCs);

    void writeToStream(std::ostream &os) const; // Predictor.cpp:137-162
    void readFromStream(std::istream &is);      // Predictor.cpp:164-192

    int getNumVisibleLayers() const { return (int)visibleLayerDescs.size(); }
    const VisibleLayer &getVisibleLayer(int i) const;
    const VisibleLayerDesc &getVisibleLayerDesc(int i) const { return visibleLayerDescs[i]; }
    const IntBuffer &getHiddenCs() const;
    const Int3 &getHiddenSize() const { return hiddenSize; }

    // --- extensions ---
    const FloatBuffer &getHiddenActivations() const; // private in the reference (Predictor.h:42)
    void getWeights(int i, std::vector<float> &out) const;
    const IntBuffer &getInputCsPrev(int i) const;
    // weights of ONE hidden column (ox, oy) towards visible layer i in the reference's CSR order [hz][slot][vz]
    void getColumnWeights(int i, int ox, int oy, std::vector<float> &out) const;
    int smallUnits() const; // see SparseCoder::smallUnits
    void stateSignature(std::vector<int> &key) const; // everything that selects this layer's kernel arguments (graph cache key)
    void hiddenRows(int &lo, int &hi) const; // hidden x rows [lo,hi) this device owns
    void visibleRows(int vli, int &lo, int &hi) const; // x rows of visible layer vli this device's slab reads
    int copyHiddenCs(int *out) const; // getHiddenCs() straight into out (pinned staging); returns the column count
    void initRandomMember(ComputeSystem &cs, const Int3 &hiddenSize, const std::vector<VisibleLayerDesc> &visibleLayerDescs,
                          int batch, int member);
    // learn (optional) + activate (optional) as ONE kernel launch on device-resident CSDRs (int32 [member][columns]).
    // deferExchange (sharded + peer exchange only): queue the all-gather of the new hiddenCs on the side stream; the next
    // exchange, getHiddenCs() and ComputeSystem::synchronize() wait for it
    // onSideStream: queue a learn-only pass on the ComputeSystem's side stream (the caller orders it)
    void stepDevice(ComputeSystem &cs, bool learn, const int *targetCsDevice, bool activate, const int *const *inputCsDevice,
                    bool deferExchange = false, bool onSideStream = false);
    const int *hiddenCsDevice() const;
    // raw device state (weights in their device layouts, CSDRs, parities) of an already constructed layer of the same
    // shape: the building block of Hierarchy::writeCheckpoint / readCheckpoint
    void writeDeviceState(std::ostream &os) const;
    void readDeviceState(std::istream &is);
    void readFromStream(std::istream &is, ComputeSystem &cs);
    void setState(const IntBuffer &hiddenCs, const std::vector<IntBuffer> &inputCsPrev);

private:
    Int3 hiddenSize;
    std::vector<VisibleLayerDesc> visibleLayerDescs;
    mutable FloatBuffer hiddenActivations;
    mutable IntBuffer hiddenCs;
    mutable std::vector<VisibleLayer> visibleLayers;
    Impl *impl;

    friend class Hierarchy;
};
} // namespace ogmaneo

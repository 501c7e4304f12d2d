// include/ogmaneo/ImageEncoder.h — B200-native drop-in for the reference's <ogmaneo/ImageEncoder.h> (:18-154): the dense
// float -> CSDR pre-encoder that sits immediately before the Sparse Predictive Hierarchy for image inputs.
// Same public interface (VisibleLayerDesc, VisibleLayer, alpha, gamma, initRandom, step, reconstruct, write/readFromStream,
// getters). Weights, resources and reconstructions live in HBM; step() and reconstruct() launch sm_100a kernels
// (ogn_image_encoder_step / ogn_image_encoder_reconstruct, include/ogn_kernels_api.h). Hidden columns may hold up to 32 cells.
#pragma once

#include <istream>
#include <ostream>

#include "ComputeSystem.h"

namespace ogmaneo {
class ImageEncoder {
public:
    struct VisibleLayerDesc { // ImageEncoder.h:21-33
        Int3 size;  // size of the dense input
        int radius; // receptive-field radius onto the input
        VisibleLayerDesc() : size(4, 4, 16), radius(2) {}
    };

    struct VisibleLayer { // ImageEncoder.h:36-40
        SparseMatrix weights;        // host mirror in the reference's CSR form (filled on demand)
        FloatBuffer reconstructions; // host mirror
    };

    struct Impl; // device state (csrc/host/ImageEncoder.cpp)

    float alpha; // resource depletion rate (ImageEncoder.h:94)
    float gamma; // rank ("gas") falloff (ImageEncoder.h:95)

    ImageEncoder();
    ImageEncoder(const ImageEncoder &other);
    ImageEncoder &operator=(const ImageEncoder &other);
    ~ImageEncoder();

    // ImageEncoder.cpp:96-137 — weights U(0,1) drawn from cs.rng in the reference's order, resources 1
    void initRandom(ComputeSystem &cs, const Int3 &hiddenSize, const std::vector<VisibleLayerDesc> &visibleLayerDescs);
    // ImageEncoder.cpp:139-148 — activate (nearest weight vector per column) and, if learnEnabled, learn
    void step(ComputeSystem &cs, const std::vector<const FloatBuffer *> &inputActs, bool learnEnabled);
    // ImageEncoder.cpp:150-160 — dense reconstruction of every visible layer from a hidden CSDR
    void reconstruct(ComputeSystem &cs, const IntBuffer *hiddenCs);

    void writeToStream(std::ostream &os) const; // ImageEncoder.cpp:162-195
    void readFromStream(std::istream &is);      // ImageEncoder.cpp:197-225

    int getNumVisibleLayers() const { return (int)visibleLayerDescs.size(); }
    const VisibleLayer &getVisibleLayer(int i) const;
    const VisibleLayerDesc &getVisibleLayerDesc(int i) const { return visibleLayerDescs[i]; }
    const IntBuffer &getHiddenCs() const;
    const Int3 &getHiddenSize() const { return hiddenSize; }

    // --- extensions (not in the reference) ---
    const FloatBuffer &getHiddenResources() const; // private upstream (ImageEncoder.h:41)
    void getWeights(int i, std::vector<float> &out) const; // nonZeroValues order, without the index arrays
    void getReconstruction(int i, FloatBuffer &out) const;
    void stepHost(ComputeSystem &cs, const float *const *inputActs, bool learnEnabled); // raw pointers (what the C ABI calls)
    void reconstructHost(ComputeSystem &cs, const int *hiddenCs);
    // device-resident step: inputActs[v] are DEVICE pointers (dense fp32, same layout); nothing is copied and nothing waits.
    // The hidden CSDR is then at getHiddenCsDevice() on the system's stream, e.g. as an input of Hierarchy::stepDevice of a
    // hierarchy created through the same ComputeSystem (or one that shares its context): image -> CSDR -> prediction
    // without a host round trip.
    void stepDevice(ComputeSystem &cs, const float *const *inputActsDevice, bool learnEnabled);
    const int *getHiddenCsDevice() const;
    void readFromStream(std::istream &is, ComputeSystem &cs); // choose the device to load onto

private:
    Int3 hiddenSize;
    std::vector<VisibleLayerDesc> visibleLayerDescs;
    mutable IntBuffer hiddenCs;
    mutable FloatBuffer hiddenResources;
    mutable std::vector<VisibleLayer> visibleLayers;
    Impl *impl;
};
} // namespace ogmaneo

// include/ogmaneo/SparseCoder.h — B200-native drop-in for the reference's <ogmaneo/SparseCoder.h> (:15-151).
// Same public interface (VisibleLayerDesc, VisibleLayer, alpha, initRandom, step, write/readFromStream, getters).
// State and weights live in HBM; step() launches the sm_100a kernels K2/K3 (include/ogn_kernels_api.h). The host
// members returned by the getters are mirrors, refreshed lazily (a getter may synchronise the stream).
#pragma once

#include <istream>
#include <ostream>

#include "ComputeSystem.h"

namespace ogmaneo {
class SparseCoder {
public:
    struct VisibleLayerDesc {
        Int3 size;  // visible (input) size
        int radius; // receptive-field radius onto the input
        VisibleLayerDesc() : size(4, 4, 16), radius(2) {}
    };

    struct VisibleLayer {
        SparseMatrix weights;        // host mirror in the reference's CSR form (filled on demand)
        FloatBuffer reconstructions; // host mirror
    };

    struct Impl; // device state (csrc/host/SparseCoder.cpp)

    float alpha; // weight learning rate

    SparseCoder();
    SparseCoder(const SparseCoder &other);
    SparseCoder &operator=(const SparseCoder &other);
    ~SparseCoder();

    // SparseCoder.cpp:85-125 — weights U(-1,0) drawn from cs.rng in the reference's order
    void initRandom(ComputeSystem &cs, const Int3 &hiddenSize, const std::vector<VisibleLayerDesc> &visibleLayerDescs);

    // SparseCoder.cpp:127-147 — forward, (learn), hiddenCs -> hiddenCsPrev. inputCs are host CSDRs.
    void step(ComputeSystem &cs, const std::vector<const IntBuffer *> &inputCs, bool learnEnabled);

    void writeToStream(std::ostream &os) const; // SparseCoder.cpp:149-171
    void readFromStream(std::istream &is);      // SparseCoder.cpp:173-203

    int getNumVisibleLayers() const { return (int)visibleLayerDescs.size(); }
    const VisibleLayer &getVisibleLayer(int i) const;
    const VisibleLayerDesc &getVisibleLayerDesc(int i) const { return visibleLayerDescs[i]; }
    const IntBuffer &getHiddenCs() const;
    const IntBuffer &getHiddenCsPrev() const;
    const Int3 &getHiddenSize() const { return hiddenSize; }

    // --- extensions (not in the reference) ---
    // Raw weight values of visible layer i in the reference's nonZeroValues order, without building the index arrays.
    void getWeights(int i, std::vector<float> &out) const;
    void getReconstructions(int i, FloatBuffer &out) const;
    // (hidden column, visible column) pairs rewritten by learn since the counter was last read (algorithmic write bytes).
    unsigned long long takeUpdateCount();
    // Test helpers: number of weights of visible layer i whose two device copies (F and R layout) differ, and the
    // weights of one hidden column in the reference's CSR order [hz][slot][vz].
    unsigned long long debugFRMismatch(int i) const;
    void getColumnWeights(int i, int ox, int oy, std::vector<float> &out) const;

    // Ensemble support: member 0 allocates room for `batch` lock-stepped copies, every member draws its own weights.
    void initRandomMember(ComputeSystem &cs, const Int3 &hiddenSize, const std::vector<VisibleLayerDesc> &visibleLayerDescs,
                          int batch, int member);
    // step() on inputs that already live in device memory (int32 [member][columns]); copyOut (optional) also receives
    // the new hiddenCs. Asynchronous on cs' stream.
    // work units per member the largest phase of this layer needs in the single-launch small-hierarchy step (one per pack of
    // the forward pass / per (visible layer, pack) of the reconstruction), or -1 if some phase has no 128-bit kernel path
    int smallUnits() const;
    void stateSignature(std::vector<int> &key) const; // everything that selects this layer's kernel arguments (graph cache key)
    void stepDevice(ComputeSystem &cs, const int *const *inputCsDevice, bool learnEnabled, int *copyOut, unsigned exchangeWaitMask = 0);
    // hidden x rows [lo, hi) of this layer's CSDR that this device's learn pass reads (owned slab + replica rows)
    void learnRows(int &lo, int &hi) const;
    const int *hiddenCsDevice() const;
    // raw device state (weights in their device layouts, CSDRs, parities) of an already constructed layer of the same
    // shape: the building block of Hierarchy::writeCheckpoint / readCheckpoint
    void visibleRows(int vli, int &lo, int &hi) const; // visible x rows [lo,hi) this device's shard of the layer reads
    int shardRank() const;
    int shardWorld() const;
    void writeDeviceState(std::ostream &os) const;
    void readDeviceState(std::istream &is);
    void readFromStream(std::istream &is, ComputeSystem &cs); // choose the device / stream to load onto
    void setHiddenState(const IntBuffer &hiddenCs, const IntBuffer &hiddenCsPrev);

private:
    Int3 hiddenSize;
    std::vector<VisibleLayerDesc> visibleLayerDescs;
    mutable IntBuffer hiddenCs, hiddenCsPrev;        // host mirrors
    mutable std::vector<VisibleLayer> visibleLayers; // host mirrors
    Impl *impl;

    friend class Hierarchy;
};
} // namespace ogmaneo

// include/ogmaneo/Hierarchy.h — B200-native drop-in for the reference's <ogmaneo/Hierarchy.h> (:15-217).
// Same public interface: InputType, State, LayerDesc, initRandom, step, getState/setState, write/readFromStream and
// all getters. The whole step is device-resident: input CSDRs are copied into an HBM history ring, the exponential-
// memory tick schedule is evaluated on the host (it is data-independent) and each ticking layer is 2 + (#predictors)
// kernel launches on the ComputeSystem's stream. step() returns without waiting for the GPU; getPredictionCs() (and
// every other getter) waits for the data it returns.
#pragma once

#include "Actor.h"
#include "Predictor.h"
#include "SparseCoder.h"

#include <memory>

namespace ogmaneo {
enum InputType { none = 0, prediction = 1, action = 2 };

// Snapshot of the CSDR state (no weights), Hierarchy.h:26-36
struct State {
    std::vector<IntBuffer> hiddenCs;
    std::vector<IntBuffer> hiddenCsPrev;
    std::vector<std::vector<std::vector<IntBuffer>>> predInputCsPrev;
    std::vector<std::vector<IntBuffer>> predHiddenCs;
    std::vector<std::vector<CircleBuffer<IntBuffer>>> histories;
    std::vector<char> updates;
    std::vector<int> ticks;
};

class Hierarchy {
public:
    struct LayerDesc {
        Int3 hiddenSize;
        int ffRadius;        // SparseCoder radius
        int pRadius;         // Predictor radius
        int ticksPerUpdate;  // ticks of the layer below per update of this layer
        int temporalHorizon; // number of past inputs the layer sees (>= ticksPerUpdate)
        int aRadius;         // Actor radius (first layer only)
        int historyCapacity; // Actor history length (first layer only)
        LayerDesc()
            : hiddenSize(4, 4, 16), ffRadius(2), pRadius(2), ticksPerUpdate(2), temporalHorizon(4), aRadius(2),
              historyCapacity(32) {}
    };

    struct Impl;

    Hierarchy();
    Hierarchy(const Hierarchy &other);
    const Hierarchy &operator=(const Hierarchy &other);
    ~Hierarchy();

    // Hierarchy.cpp:16-147
    void initRandom(ComputeSystem &cs, const std::vector<Int3> &inputSizes, const std::vector<InputType> &inputTypes,
                    const std::vector<LayerDesc> &layerDescs);
    // Hierarchy.cpp:192-285
    void step(ComputeSystem &cs, const std::vector<const IntBuffer *> &inputCs, bool learnEnabled = true, float reward = 0.0f,
              bool mimic = false);

    void getState(State &state) const;          // Hierarchy.cpp:434-466
    void setState(const State &state);          // Hierarchy.cpp:468-493
    void writeToStream(std::ostream &os) const; // Hierarchy.cpp:287-344
    void readFromStream(std::istream &is);      // Hierarchy.cpp:346-432

    int getNumLayers() const { return (int)scLayers.size(); }
    const IntBuffer &getPredictionCs(int i) const {
        if (aLayers[i] != nullptr) return aLayers[i]->getHiddenCs();
        return pLayers.front()[i]->getHiddenCs();
    }
    bool getUpdate(int l) const { return updates[l]; }
    int getTicks(int l) const { return ticks[l]; }
    int getTicksPerUpdate(int l) const { return ticksPerUpdate[l]; }
    const std::vector<Int3> &getInputSizes() const { return inputSizes; }
    SparseCoder &getSCLayer(int l) { return scLayers[l]; }
    const SparseCoder &getSCLayer(int l) const { return scLayers[l]; }
    std::vector<std::unique_ptr<Predictor>> &getPLayers(int l) { return pLayers[l]; }
    const std::vector<std::unique_ptr<Predictor>> &getPLayers(int l) const { return pLayers[l]; }
    std::vector<std::unique_ptr<Actor>> &getALayers() { return aLayers; }
    const std::vector<std::unique_ptr<Actor>> &getALayers() const { return aLayers; }

    // --- extensions (not in the reference) ---
    // step() on raw host pointers (input i holds inputSizes[i].x * inputSizes[i].y * ensemble ints): what the C ABI calls,
    // without building IntBuffers first
    void stepHost(ComputeSystem &cs, const int *const *inputCs, bool learnEnabled = true, float reward = 0.0f, bool mimic = false);
    // getPredictionCs(i) copied straight into out (pinned staging, no host mirror refresh); returns the column count
    // (out == nullptr: only the count)
    int copyPredictionCs(int i, int *out) const;
    // Same as step() but the inputs are already in device memory (int32, [member][columns] each) and nothing is copied
    // back: the whole step is asynchronous on cs' stream.
    void stepDevice(ComputeSystem &cs, const std::vector<const int *> &inputCsDevice, bool learnEnabled = true,
                    float reward = 0.0f, bool mimic = false);
    // Device pointer to the prediction CSDR of input i (valid until the hierarchy is destroyed).
    const int *getPredictionCsDevice(int i) const;
    // Pipelined readback for streaming callers: queue() copies x rows [x0, x1) of the CURRENT prediction CSDR of input i
    // (x1 <= x0: all of it) to pinned memory on a copy stream, behind the work queued so far, and returns a ticket; finish()
    // waits for that copy only and hands out the ints. Two tickets may be outstanding, so step t + 1 can be queued before
    // the prediction of step t is collected. A sharded rank that asks for rows of its own slab does not wait for the
    // exchange with the other ranks.
    int queuePredictionReadback(int i, int x0, int x1);
    int finishPredictionReadback(int ticket, int *out);
    // An ensemble of seeds.size() independent hierarchies of identical shape stepped in lockstep by the same kernel
    // launches (BASELINE config 5). Member k is initialised exactly like initRandom() with cs.rng.seed(seeds[k]); every
    // CSDR passed to / returned from the hierarchy then holds the members back to back ([member][columns]).
    void initRandomEnsemble(ComputeSystem &cs, const std::vector<unsigned> &seeds, const std::vector<Int3> &inputSizes,
                            const std::vector<InputType> &inputTypes, const std::vector<LayerDesc> &layerDescs);
    int getEnsembleSize() const;
    // Checkpoint of hierarchies the reference's save format cannot hold (more than 2^31 weights per matrix, ensembles,
    // x-slab shards: SURVEY.md §8f-1): 64-bit counts, weights streamed in their device layouts through pinned staging,
    // one file per rank. Layout: "OGNCKPT1", the initRandom arguments, ensemble size and shard (rank, world), tick state,
    // history rings, then SparseCoder / Predictor device state layer by layer. readCheckpoint rebuilds the hierarchy from
    // the header (device-side init, then overwritten).
    // Actor layers are covered (weights, history ring, hyper-parameters); ComputeSystem::rng is the caller's to carry over
    // (ogn_checkpoint_save stores it next to the hierarchy).
    void writeCheckpoint(std::ostream &os, const std::vector<InputType> &inputTypes, const std::vector<LayerDesc> &layerDescs) const;
    void readCheckpoint(std::istream &is, ComputeSystem &cs, std::vector<InputType> *inputTypesOut = nullptr);
    void readFromStream(std::istream &is, ComputeSystem &cs); // choose the device / stream to load onto

private:
    std::vector<SparseCoder> scLayers;
    std::vector<std::vector<std::unique_ptr<Predictor>>> pLayers;
    std::vector<std::unique_ptr<Actor>> aLayers;
    std::vector<char> updates;
    std::vector<int> ticks;
    std::vector<int> ticksPerUpdate;
    std::vector<Int3> inputSizes;
    Impl *impl; // device history rings + staging

    void initImpl(ComputeSystem &cs, const std::vector<Int3> &inputSizes, const std::vector<InputType> &inputTypes,
                  const std::vector<LayerDesc> &layerDescs, int batch, int member);
    bool smallEligible(ComputeSystem &cs);
    unsigned haloWaitMask(ComputeSystem &cs, int l);
    void stepImpl(ComputeSystem &cs, const int *const *hostIn, const std::vector<const int *> *devIn,
                  bool learnEnabled, float reward, bool mimic);
};
} // namespace ogmaneo

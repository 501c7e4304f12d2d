/* ogn_hierarchy_api.h — the flat C ABI of the Sparse Predictive Hierarchy step path.
 *
 * The reference (ogmacorp/OgmaNeo2) has no FFI of its own: consumers include <ogmaneo/Hierarchy.h> and link
 * libOgmaNeo. This header is the C-ABI a binding (ctypes / cgo / JNI / SWIG as in PyOgmaNeo2) would bind for the
 * step path; every entry point names the reference interface it replaces (file:line under source/ogmaneo/).
 * Plain pointers and sizes only. All buffers passed here are HOST buffers unless the name says `_dev`.
 *
 * The same set of functions is exported three times with different prefixes, so that one test harness can drive
 * all of them on identical inputs:
 *     ogn_    the product: B200-native CUDA path (libogn_b200.so). No CPU fallback: every call fails with a
 *             non-zero code if no sm_100 device / kernel image is available.
 *     oport_  oracle/: the CPU restatement of the reference algorithm (test infrastructure only).
 *     oref_   oracle/_ref/: the unmodified reference sources compiled with -Dogmaneo=ogmaneo_ref (test
 *             infrastructure only; exists only where /root/reference was present at build time).
 *
 * Conventions: functions returning int return 0 on success and a non-zero code on failure (P##last_error() gives
 * the message) unless documented to return a count. CSDRs are int32, one active-cell index per column, column
 * (x, y) at index y + x*size_y (Helpers.h:240-245). Weight dumps are fp32 in the reference's CSR order
 * (Helpers.cpp:236-313): row (ox,oy,oz), then ix, iy, iz ascending inside the clamped receptive field.
 */
#ifndef OGN_HIERARCHY_API_H
#define OGN_HIERARCHY_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Hierarchy::LayerDesc, Hierarchy.h:42-66 (same fields, same defaults: 4,4,16, 2, 2, 2, 4, 2, 32). */
typedef struct ogn_layer_desc {
    int hidden_x, hidden_y, hidden_z; /* Int3 hiddenSize */
    int ff_radius;                    /* ffRadius */
    int p_radius;                     /* pRadius */
    int ticks_per_update;             /* ticksPerUpdate */
    int temporal_horizon;             /* temporalHorizon */
    int a_radius;                     /* aRadius */
    int history_capacity;             /* historyCapacity */
} ogn_layer_desc;

/* InputType, Hierarchy.h:19-23 */
enum { OGN_INPUT_NONE = 0, OGN_INPUT_PREDICTION = 1, OGN_INPUT_ACTION = 2 };

/* Arguments of Hierarchy::initRandom, Hierarchy.h:102-107. input_sizes holds 3 ints (x,y,z) per input. */
typedef struct ogn_hierarchy_desc {
    int num_inputs;
    const int *input_sizes;
    const int *input_types;
    int num_layers;
    const ogn_layer_desc *layers;
} ogn_hierarchy_desc;

#define OGN_DECLARE_HIERARCHY_API(P)                                                                               \
    /* ComputeSystem cs; cs.rng.seed(seed); Hierarchy::initRandom(cs, ...)  — ComputeSystem.h:17-41,             \
       Hierarchy.cpp:16-147. The handle owns both the ComputeSystem and the Hierarchy. */                          \
    int P##create(const ogn_hierarchy_desc *desc, uint32_t seed, void **out);                                      \
    void P##destroy(void *h);                                                                                      \
    const char *P##last_error(void);                                                                               \
    /* ComputeSystem::setNumThreads, ComputeSystem.h:34-36 (host OpenMP threads; no effect on the device path) */  \
    int P##set_num_threads(int n);                                                                                 \
    /* Hierarchy::step, Hierarchy.cpp:192-285. input_cs[i] -> num columns of input i ints. */                      \
    int P##step(void *h, const int *const *input_cs, int learn_enabled, float reward, int mimic);                  \
    /* Hierarchy::getNumLayers / getUpdate / getTicks / getTicksPerUpdate, Hierarchy.h:139-172 */                  \
    int P##num_layers(void *h);                                                                                    \
    int P##get_update(void *h, int l);                                                                             \
    int P##get_ticks(void *h, int l);                                                                              \
    int P##get_ticks_per_update(void *h, int l);                                                                   \
    /* Hierarchy::getPredictionCs, Hierarchy.h:144-151. Returns the number of ints written (columns of input i),  \
       or -1 if input i has neither predictor nor actor. */                                                        \
    int P##get_prediction_cs(void *h, int i, int *out);                                                            \
    /* SparseCoder::getHiddenCs / getHiddenCsPrev / getNumVisibleLayers, SparseCoder.h:117-144 */                  \
    int P##sc_hidden_cs(void *h, int l, int *out);                                                                 \
    int P##sc_hidden_cs_prev(void *h, int l, int *out);                                                            \
    int P##sc_num_visible(void *h, int l);                                                                         \
    /* SparseCoder::getVisibleLayer(vli).weights.nonZeroValues, SparseCoder.h:121-126. Returns the value count;   \
       writes at most cap floats if out != NULL. */                                                                \
    int64_t P##sc_weights(void *h, int l, int vli, float *out, int64_t cap);                                       \
    /* SparseCoder::getVisibleLayer(vli).reconstructions */                                                        \
    int P##sc_reconstructions(void *h, int l, int vli, float *out);                                                \
    /* Hierarchy::getPLayers(l)[p], Hierarchy.h:194-206; Predictor getters, Predictor.h:113-142.                   \
       pred_count = pLayers[l].size(); pred_exists = (pLayers[l][p] != nullptr). */                                \
    int P##pred_count(void *h, int l);                                                                             \
    int P##pred_exists(void *h, int l, int p);                                                                     \
    int P##pred_hidden_cs(void *h, int l, int p, int *out);                                                        \
    int P##pred_activations(void *h, int l, int p, float *out);                                                    \
    int P##pred_num_visible(void *h, int l, int p);                                                                \
    int64_t P##pred_weights(void *h, int l, int p, int vli, float *out, int64_t cap);                              \
    int P##pred_input_cs_prev(void *h, int l, int p, int vli, int *out);                                           \
    /* Hierarchy::getALayers()[p], Hierarchy.h:209-215; Actor getters, Actor.h:150-178. which: 0 = valueWeights,   \
       1 = actionWeights. */                                                                                       \
    int P##actor_exists(void *h, int p);                                                                           \
    int P##actor_hidden_cs(void *h, int p, int *out);                                                              \
    int P##actor_values(void *h, int p, float *out);                                                               \
    int P##actor_num_visible(void *h, int p);                                                                      \
    int64_t P##actor_weights(void *h, int p, int vli, int which, float *out, int64_t cap);                         \
    /* public hyper-parameter members: SparseCoder::alpha (SparseCoder.h:83), Predictor::alpha (Predictor.h:82),   \
       Actor::{alpha,beta,gamma,minSteps,historyIters} (Actor.h:108-112) */                                        \
    int P##set_sc_alpha(void *h, int l, float alpha);                                                              \
    int P##set_pred_alpha(void *h, int l, int p, float alpha);                                                     \
    int P##set_actor_params(void *h, int p, float alpha, float beta, float gamma, int min_steps,                   \
                            int history_iters);                                                                    \
    /* ... and reading them back (`h.getSCLayer(l).alpha` in user code), plus the layer shapes: SparseCoder /      \
       Predictor / Actor::getHiddenSize and getVisibleLayerDesc (SparseCoder.h:129-147, Predictor.h:128-146,       \
       Actor.h:160-178). kind: 0 SparseCoder of layer l, 1 + p Predictor p of layer l, -1 - p Actor of input p     \
       (l ignored). hidden_size writes 3 ints; visible_desc writes {size.x, size.y, size.z, radius}. */            \
    float P##get_sc_alpha(void *h, int l);                                                                         \
    float P##get_pred_alpha(void *h, int l, int p);                                                                \
    int P##get_actor_params(void *h, int p, float *alpha, float *beta, float *gamma, int *min_steps,               \
                            int *history_iters);                                                                   \
    int P##layer_hidden_size(void *h, int kind, int l, int *out3);                                                 \
    int P##layer_visible_desc(void *h, int kind, int l, int vli, int *out4);                                       \
    /* Hierarchy::writeToStream / readFromStream on a file, Hierarchy.cpp:287-432 (byte layout: SURVEY.md App. B). \
       save returns bytes written or -1. load creates a new handle whose ComputeSystem rng is seeded with seed. */ \
    int64_t P##save(void *h, const char *path);                                                                    \
    int P##load(const char *path, uint32_t seed, void **out);                                                      \
    /* Hierarchy::getState / setState, Hierarchy.cpp:434-493 (struct State, Hierarchy.h:26-36): a snapshot of every  \
       CSDR of the hierarchy (hidden states, predictor outputs and previous inputs, history rings, ticks, updates;  \
       no weights) as an opaque object owned by the caller; set_state rewinds the hierarchy to it. */              \
    int P##get_state(void *h, void **state_out);                                                                   \
    int P##set_state(void *h, const void *state);                                                                  \
    void P##free_state(void *state);                                                                               \
    /* Next raw output of a COPY of cs.rng (ComputeSystem.h:24): lets tests compare RNG consumption without        \
       disturbing it. */                                                                                           \
    uint32_t P##rng_peek(void *h);

OGN_DECLARE_HIERARCHY_API(ogn_)

#ifdef __cplusplus
}
#endif
#endif /* OGN_HIERARCHY_API_H */

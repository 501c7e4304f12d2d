/* ogn_b200.h — product-only entry points of libogn_b200.so, on top of the flat hierarchy ABI
 * (ogn_hierarchy_api.h). These have no counterpart in the reference: they expose what a device-resident
 * implementation can do beyond Hierarchy::step on host buffers — stepping on inputs already in HBM, ensembles of
 * lock-stepped hierarchies (BASELINE config 5), x-slab sharding of large layers across GPUs (config 4) and a
 * device-side weight init for sizes the reference cannot construct. All functions return 0 / non-zero like the flat
 * ABI; ogn_last_error() has the message.
 */
#ifndef OGN_B200_H
#define OGN_B200_H

#include "ogn_hierarchy_api.h"
#include "ogn_kernels_api.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Called after a layer wrote its x-slab of a CSDR into `buffer` ([shard_world][count_per_rank] int32, this rank's slab
 * in place): must complete `buffer` on every rank, ordered on `stream` (cudaStream_t) — e.g. ncclAllGather in place,
 * torch.distributed.all_gather_into_tensor on that stream, or a peer-memory kernel. */
typedef void (*ogn_all_gather_fn)(void *user, int *buffer, int count_per_rank, void *stream);

typedef struct ogn_create_options {
    int device;            /* CUDA device ordinal, -1 = current */
    int device_init;       /* 0: weights drawn from std::mt19937 in the reference's order (bit-identical to the reference);
                              1: counter-hash init on the device keyed by (seed, matrix, CSR index) */
    int replay_rng;        /* 1: replay every runKernelN seed draw so ComputeSystem::rng tracks the reference exactly
                              (always on when the hierarchy has an Actor) */
    int shard_rank, shard_world; /* x-slab sharding of every layer over shard_world devices (one process per device) */
    ogn_all_gather_fn all_gather;
    void *all_gather_user;
    int ensemble;          /* number of independent lock-stepped hierarchies (>= 1) */
    const uint32_t *ensemble_seeds; /* optional [ensemble]; default seed + k */
} ogn_create_options;

/* ogn_create with options. ogn_create(desc, seed, out) == ogn_create_ex with {device -1, device_init 0, replay_rng 1}. */
int ogn_create_ex(const ogn_hierarchy_desc *desc, uint32_t seed, const ogn_create_options *opt, void **out);

/* Hierarchy::step on inputs that already live in device memory (int32 [member][columns] per input). Fully asynchronous
 * on the hierarchy's stream: nothing is copied to or from the host. */
int ogn_step_device(void *h, const int *const *input_cs_dev, int learn_enabled, float reward, int mimic);
/* Device pointer of the prediction CSDR of input i ([member][columns]); valid for the life of the handle. */
const int *ogn_prediction_cs_device(void *h, int i);
/* The cudaStream_t all work of this hierarchy is queued on / wait for it to drain. */
void *ogn_stream(void *h);
int ogn_synchronize(void *h);
int ogn_ensemble_size(void *h);
/* (hidden column, visible column) pairs whose weights SparseCoder l rewrote since the last call: the data-dependent
 * term U of the algorithmic byte count. Synchronises. */
uint64_t ogn_take_sc_update_count(void *h, int l);
/* Launch accounting. Kernel kinds, in array order: 0 SparseCoder forward (K2), 1 SparseCoder learn (K3), 2 Predictor
 * learn+activate (K6+K7+K8), 3 Actor forward (K9), 4 Actor learn (K11). enable(on) resets the counters; with on != 0
 * every launch is bracketed by CUDA events on the hierarchy's stream. read() synchronises and returns the accumulated
 * device milliseconds and launch counts per kind (arrays of OGN_KERNEL_KINDS). */
#define OGN_KERNEL_KINDS 5
int ogn_profile_enable(void *h, int on);
int ogn_profile_read(void *h, double *ms, int64_t *launches);
int ogn_device_count(void);
/* Test helpers. fr_mismatch: number of SparseCoder weights whose F and R copies differ (must be 0; synchronises).
 * column_weights: the weights of ONE hidden column (ox, oy) of SparseCoder l (kind 0), visible layer vli, in the
 * reference's CSR order [hz][slot][vz]; returns the count. Lets a host check sizes the oracle cannot build. */
int64_t ogn_debug_sc_fr_mismatch(void *h, int l, int vli);
int64_t ogn_debug_column_weights(void *h, int kind, int l, int vli, int ox, int oy, float *out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* OGN_B200_H */

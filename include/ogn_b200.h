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
    int peer_exchange;     /* 1 (with shard_world > 1, one process per GPU on one NVLink node): exchange the CSDR slabs with
                              the library's own peer-memory kernel instead of the all_gather callback; call ogn_peer_export,
                              pass every rank's handle to ogn_peer_attach, then step */
} ogn_create_options;

/* ogn_create with options. ogn_create(desc, seed, out) == ogn_create_ex with {device -1, device_init 0, replay_rng 1}. */
int ogn_create_ex(const ogn_hierarchy_desc *desc, uint32_t seed, const ogn_create_options *opt, void **out);

/* Peer-memory exchange set-up (peer_exchange = 1). export: the 64-byte CUDA IPC handle of this rank's exchange arena.
 * attach: handles = world x 64 bytes in rank order (this rank's own entry is ignored); maps the peers' arenas. The
 * transport of the handles between the processes is the caller's (torch.distributed, MPI, a file ...). */
#define OGN_PEER_HANDLE_BYTES 64
int ogn_peer_export(void *h, void *handle_out);
int ogn_peer_attach(void *h, const void *handles, int world);
/* Nanoseconds of device time this rank's exchange kernels spent waiting for the other ranks since the last call (load
 * imbalance + NVLink latency; synchronises). */
uint64_t ogn_peer_take_wait_ns(void *h);

/* The x-slab plan of one (hidden grid, visible grid, radius) pairing for `rank` of `world` (host only, no device
 * needed): out[0..1] = hidden rows [x0,x1) this rank owns (forward, Predictor, F copy); out[2..3] = visible rows
 * [vx0,vx1) whose reconstruction / learn it runs; out[4..5] = hidden rows [rx0,rx1) whose R copy it keeps (owner +
 * replicas). Returns non-zero if hidden_x is not a multiple of world. */
int ogn_shard_plan(int hidden_x, int visible_x, int radius, int rank, int world, int out[6]);

/* Checkpoints for hierarchies the reference's save format cannot hold (more than 2^31 weights per matrix, ensembles,
 * shards): 64-bit counts, weights streamed in their device layouts, one file per rank (Hierarchy::writeCheckpoint).
 * save returns the bytes written or -1. load rebuilds the hierarchy from the file's header; opt supplies the device and,
 * for a sharded checkpoint, the same shard_rank / shard_world (and exchange settings) the file was written with. */
int64_t ogn_checkpoint_save(void *h, const char *path);
int ogn_checkpoint_load(const char *path, const ogn_create_options *opt, void **out);

/* Hierarchy::step on inputs that already live in device memory (int32 [member][columns] per input). Fully asynchronous
 * on the hierarchy's stream: nothing is copied to or from the host. */
int ogn_step_device(void *h, const int *const *input_cs_dev, int learn_enabled, float reward, int mimic);
/* Device pointer of the prediction CSDR of input i ([member][columns]); valid for the life of the handle. */
const int *ogn_prediction_cs_device(void *h, int i);
/* Pipelined readback of the prediction CSDR of input i for streaming callers (Hierarchy::getPredictionCs, Hierarchy.h:144-151,
 * without stalling the queue): queue() copies x rows [x0, x1) of the CURRENT prediction (x1 <= x0: all of it) to pinned
 * memory on a copy stream, behind the work queued so far, and returns a ticket (0 or 1; -1 on error); finish() waits for
 * that copy only, writes the ints to out and returns their number. Two tickets may be outstanding: queue step t + 1 before
 * collecting the prediction of step t. A shard that asks for rows of its own slab does not wait for the other ranks. */
int ogn_prediction_readback_queue(void *h, int i, int x0, int x1);
int ogn_prediction_readback_finish(void *h, int ticket, int *out);
/* The cudaStream_t all work of this hierarchy is queued on / wait for it to drain. */
void *ogn_stream(void *h);
int ogn_synchronize(void *h);
int ogn_ensemble_size(void *h);
/* (hidden column, visible column) pairs whose weights SparseCoder l rewrote since the last call: the data-dependent
 * term U of the algorithmic byte count. Synchronises. */
uint64_t ogn_take_sc_update_count(void *h, int l);
/* Launch accounting. Kernel kinds, in array order: 0 SparseCoder forward (K2), 1 SparseCoder learn (K3: the
 * reconstruction when learn runs as two kernels, else the fused learn), 2 Predictor learn+activate (K6+K7+K8), 3 Actor
 * forward (K9), 4 Actor learn (K11), 5 SparseCoder learn part 2 (work-list update), 6 the single-launch small-hierarchy step
 * (every phase of a step in one kernel, ogn_hier_step_small), 7 the peer-memory CSDR exchange of sharded layers. enable(on) resets the counters; with on != 0
 * the launches of every on-th Hierarchy::step (on = 1: every step) are bracketed by CUDA events on the stream they are
 * queued on. read() synchronises and returns, per kind (arrays of OGN_KERNEL_KINDS): the device milliseconds summed
 * over the bracketed launches, the number of launches, and the number of bracketed launches. */
#define OGN_KERNEL_KINDS 8
int ogn_profile_enable(void *h, int on);
int ogn_profile_read(void *h, double *ms, int64_t *launches, int64_t *timed_launches);
int ogn_device_count(void);
/* Test helpers. fr_mismatch: number of SparseCoder weights whose F and R copies differ (must be 0; synchronises).
 * column_weights: the weights of ONE hidden column (ox, oy) of SparseCoder l (kind 0) or of Predictor kind-1 of layer l
 * (kind >= 1), visible layer vli, in the reference's CSR order [hz][slot][vz]; returns the count. Lets a host check sizes
 * the oracle cannot build, and lets bench.py hash a weight sample on whichever rank stores the column. */
int64_t ogn_debug_sc_fr_mismatch(void *h, int l, int vli);
int64_t ogn_debug_column_weights(void *h, int kind, int l, int vli, int ox, int oy, float *out, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* OGN_B200_H */

/* ogn_kernels_api.h — the thin extern "C" layer between the C++ host classes (include/ogmaneo/) and the
 * hand-written sm_100a kernels (ogmaneo2_b200/csrc/ogn_kernels.cu).
 *
 * Every launcher replaces one runKernelN call site of the reference; the reference
 * file:line is cited at each declaration. All pointers are DEVICE pointers unless the name ends in _host.
 * Launchers are asynchronous on `stream` (a cudaStream_t passed as void*), allocate nothing, never synchronise,
 * and return 0 or a cudaError_t value (ogn_cuda_error_string() decodes it).
 *
 * Data layout in HBM:
 *   CSDR            int32 [member][x][y]                      (member = ensemble index, 1 for a single hierarchy)
 *   activations     fp32  [member][x][y][z]
 *   weights: every gathered "row" (the Z cells selected by one column, one receptive-field slot and one active input
 *   cell) is contiguous, and P = 32/pow2ceil(Z) (at most 4) y-adjacent columns are PACKED so that the rows a warp
 *   needs for one visible column form one full 128-byte line:
 *     "F" (row gather: SparseCoder forward, Predictor, Actor)
 *         [hidden x][hidden-y pack q][dx][dyU][visible cell vz][p = oy - q*pf][hidden cell hz]
 *         dyU runs over the UNION of the pack's fields along y (loyf/nyf); slots outside a column's own field are padding
 *     "R" (SparseCoder reconstruction; SparseCoder only)
 *         [hidden x][hidden y][dx][visible-y group g = iy/pr][hidden cell hz][iy % pr][visible cell vz]
 *   With Z = 32 (P = 1) both reduce to the reference's CSR slot order with the cell index moved innermost. Block
 *   offsets are closed-form products of per-axis prefix sums (no index arrays) and always 64-bit.
 */
#ifndef OGN_KERNELS_API_H
#define OGN_KERNELS_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OGN_MAX_VL 32 /* visible layers per launch (inputs x temporalHorizon of one SparseCoder) */

/* Receptive-field geometry of one (hidden grid, visible grid, radius) pairing: initSMLocalRF in closed form
 * (Helpers.cpp:236-313, project() Helpers.h:223-228). Passed to kernels BY VALUE; the tables it points to live in
 * device memory. */
typedef struct ogn_rf {
    int Hx, Hy, Hz;           /* hidden (output) grid */
    int Vx, Vy, Vz;           /* visible (input) grid */
    int radius;
    int sy;                   /* sum of ny[] over all hidden y (CSR order) */
    int pf, pr;               /* pack factors of the F layout (hidden y) and the R layout (visible y) */
    int syf;                  /* sum of nyf[] over all hidden-y packs */
    int syr;                  /* sum of ngr[] over all hidden y */
    const int *lox, *nx, *px; /* [Hx] first visible x of the field, extent, exclusive prefix sum of nx */
    const int *loy, *ny, *py; /* [Hy] same along y */
    const int *loyf, *nyf, *pyf; /* [ceil(Hy/pf)] F packs: first visible y of the union field, extent, prefix sum */
    const int *ngr, *pyr;     /* [Hy] R: number of visible-y groups (iy/pr) the field touches, prefix sum */
    const int *rlox, *rhix;   /* [Vx] first / last hidden x whose field covers visible x (rlox > rhix: none) */
    const int *rloy, *rhiy;   /* [Vy] */
    int max_nx, max_nyf;      /* largest nx[] / nyf[] entry: bounds the slot descriptors of the 128-bit kernels */
    int max_cov;              /* largest number of hidden columns covering one visible column */
    int tma_ok;               /* 1: CSDR windows of this pairing can be staged by bulk copies (Vy % 4 == 0, every union field
                                 non-empty, field starts and ends non-decreasing along y) */
} ogn_rf;

/* ---- per-launch descriptors, all passed BY VALUE (kernel parameter space): a launch never chases a descriptor
 * pointer through HBM before it can issue its first weight load. Weight pointers are "virtual bases": the address of
 * the (possibly not stored) first value of the whole matrix, i.e. storage pointer minus the block offset of the first
 * stored column, so kernels index with global block offsets. ---- */
#define OGN_MAX_GEO 4 /* distinct (visible shape, radius) pairings per launch */

typedef struct ogn_fwd_vl {       /* one visible layer of a row-gather launch (SparseCoder forward) */
    const float *wf;              /* F layout virtual base */
    const int *cs;                /* visible CSDR [member][Vx][Vy] */
    int64_t mstride;              /* floats between consecutive ensemble members */
    int geo;                      /* index into the launch's geometry table */
    int pad;
} ogn_fwd_vl;

typedef struct ogn_fwd_args {
    ogn_rf geo[OGN_MAX_GEO];
    ogn_fwd_vl vl[OGN_MAX_VL];
    int nvl;
} ogn_fwd_args;

typedef struct ogn_learn_vl {     /* one visible layer of a SparseCoder learn launch */
    float *wf, *wr;               /* F and R layout virtual bases */
    float *recon;                 /* reconstructions [member][Vx][Vy][Vz] (SparseCoder.h:35) */
    const int *cs;                /* visible CSDR = reconstruction target */
    int64_t f_mstride, r_mstride;
} ogn_learn_vl;

typedef struct ogn_learn_args {   /* visible layers of one launch share the geometry */
    ogn_rf geo;
    ogn_learn_vl vl[OGN_MAX_VL];
    int nvl;
    int fx0, fx1;                 /* hidden x range whose F copy is stored on this device */
    /* work list of the two-kernel learn (reconstruction, then update of the listed visible columns): room for
     * batch * nvl * (vx1 - vx0) * Vy entries of 8 bytes, and two zero-initialised counters {entries, warps done} that
     * the update kernel leaves at zero again. NULL selects the single fused kernel. */
    void *work;
    unsigned *work_counters;
} ogn_learn_args;

#define OGN_MAX_PRED_VL 8
typedef struct ogn_pred_vl {
    float *wf;                    /* F layout virtual base */
    const int *cs;                /* current inputs (Predictor::activate argument) */
    const int *cs_prev;           /* VisibleLayer::inputCsPrev (Predictor.h:36) */
    int *cs_prev_out;             /* where this launch stores cs for the next learn (ping-pong buffer) or NULL */
    int64_t mstride;
    int geo, pad;
} ogn_pred_vl;

typedef struct ogn_pred_args {
    ogn_rf geo[OGN_MAX_GEO];
    ogn_pred_vl vl[OGN_MAX_PRED_VL];
    int nvl;
} ogn_pred_args;

typedef struct ogn_actor_vl {
    float *wv, *wa;               /* value (Hz = 1) and action weights, F layout virtual bases */
    const int *cs;                /* inputs (forward) or the history sample's inputs (learn) */
    int geo_v, geo_a;             /* geometry indices of the value matrix (Hx,Hy,1) and the action matrix (Hx,Hy,Hz) */
} ogn_actor_vl;

typedef struct ogn_actor_args {
    ogn_rf geo[OGN_MAX_GEO];
    ogn_actor_vl vl[OGN_MAX_PRED_VL];
    int nvl;
} ogn_actor_args;

/* Storage description used by the (re)layout / init utilities (host-side struct, not a kernel parameter). */
typedef struct ogn_wset {
    float *wf;                /* F layout storage, hidden x in [fx0, fx1) */
    float *wr;                /* R layout storage, hidden x in [rx0, rx1); NULL unless SparseCoder */
    int64_t wf_base, wr_base; /* F / R block offset of hidden x = fx0 / rx0 */
    int64_t wf_mstride, wr_mstride; /* floats between consecutive ensemble members */
    int fx0, fx1, rx0, rx1;
} ogn_wset;

/* History samples of one Actor::learn launch (Actor.cpp:280-309: historyIters samples per step, applied in order). */
#define OGN_MAX_ACTOR_SAMPLES 16
typedef struct ogn_actor_sample {
    const int *cs[OGN_MAX_PRED_VL]; /* sPrev.inputCs of every visible layer */
    const int *target_prev;         /* s.hiddenTargetCsPrev */
    const float *values_prev;       /* sPrev.hiddenValues */
    float q, g;                     /* discounted reward sum and gamma^(index+1) */
} ogn_actor_sample;
typedef struct ogn_actor_batch {
    ogn_actor_sample it[OGN_MAX_ACTOR_SAMPLES];
    int n;
} ogn_actor_batch;

/* One phase of the single-launch small-hierarchy step (ogn_hier_step_small): the arguments one of the launchers below
 * would have been given, plus the scheduling flags of the in-kernel phase loop. An array of these lives in DEVICE memory
 * (the union makes an element ~1.8 KB; a step of a 4-layer hierarchy has at most ~20 phases). */
#define OGN_PHASE_SC_FORWARD 0
#define OGN_PHASE_SC_RECON 1
#define OGN_PHASE_SC_UPDATE 2
#define OGN_PHASE_PRED 3
#define OGN_SMALL_MAX_LISTS 30  /* work lists (reconstruction/update pairs) per step */
#define OGN_SMALL_COUNTERS (2 + OGN_SMALL_MAX_LISTS) /* per member: barrier {arrivals, generation} + list fills */
typedef struct ogn_small_phase {
    int kind;                 /* OGN_PHASE_* */
    int z;                    /* row width of the phase's 128-bit kernels: Hz (forward, Predictor) or Vz (reconstruction, update) */
    int unitShift;            /* first of the launch's unit slots (warps, wide mapping) this phase starts on: the phases between two
                                 barriers are independent, and disjoint slot ranges make them run side by side */
    int barrierAfter;         /* 1: the blocks of a member meet before the next phase starts */
    int counter;              /* work-list slot (reconstruction / update pairs share one) */
    int Hx, Hy;               /* grid the packs tile: hidden grid (forward, Predictor), visible grid (reconstruction) */
    int npacks, ppr;          /* packs per member and per grid row */
    int learn, activate;      /* Predictor */
    int updVer, chunks;       /* update: kernel version (1 | 2) and work units per list entry */
    int updBatches;           /* update version 2: batches of 4 pairs per work unit */
    float alpha;
    int *out0, *out1;         /* forward: hiddenCs + optional copy; Predictor: hiddenCs */
    const int *target;        /* Predictor learn target */
    float *acts;              /* Predictor activations */
    const int *hcs, *hcsPrev; /* reconstruction / update: SparseCoder hiddenCs and hiddenCsPrev */
    void *work;               /* work list base; member m uses [m * workStride, (m + 1) * workStride) entries of 8 bytes */
    int64_t workStride;
    unsigned long long *updCount;
    union {
        ogn_fwd_args fwd;
        ogn_learn_args learn;
        ogn_pred_args pred;
    } u;
} __attribute__((aligned(16))) ogn_small_phase; /* copied into shared memory with 16-byte loads */

const char *ogn_cuda_error_string(int code);

/* 1 if rows of z cells / this geometry have a 128-bit kernel path (rows of 8, 16 or 32 cells; slot descriptors fit their bit
 * fields): what a caller planning a single-launch step needs to know. */
int ogn_vec_row_ok(int z);
int ogn_vec_geo_ok(const ogn_rf *rf_host);
/* update kernel version the library would use for this geometry (2 needs Hz % 8 == 0; OGN_SC_UPDATE=1|2 overrides), and
 * the work units per list entry for that version */
int ogn_sc_update_version(const ogn_learn_args *args_host);
int ogn_sc_update_chunks(const ogn_learn_args *args_host, int version);
int ogn_sc_update_batches(void); /* version 2: batches of 4 pairs one work unit rewrites (OGN_UPD2_BATCHES, default 4) */

/* The whole step of a small hierarchy in ONE launch (SURVEY.md 8(b2) ogn_hier_step_small; replaces the ~54 runKernelN
 * fork/joins of Hierarchy::step, Hierarchy.cpp:219-284, Helpers.cpp:15-72): the phases (device array) run in order
 * inside one kernel, the blocks of a member meeting at a barrier wherever barrierAfter is set. wide = 1: one warp per pack
 * (latency mapping: a single small hierarchy), 0: one octet per pack (bandwidth mapping: ensembles). grid =
 * (blocks_per_member, members); blocks_per_member > 1 launches cooperatively (all blocks co-resident) and fails with
 * cudaErrorCooperativeLaunchTooLarge if they do not fit. counters: members x OGN_SMALL_COUNTERS zero-initialised words,
 * left at zero. max_blocks_per_member(threads) tells how many blocks of `threads` threads can be co-resident. */
/* K1 inside the same launch (Hierarchy.cpp:205-212): before the first phase every member copies count[i] ints of src[i]
 * (device-resident input CSDR, [member][count]) to dst[i] (the front slot of the layer-0 history ring). */
#define OGN_SMALL_MAX_INPUTS 4
typedef struct ogn_small_inputs {
    const int *src[OGN_SMALL_MAX_INPUTS];
    int *dst[OGN_SMALL_MAX_INPUTS];
    int count[OGN_SMALL_MAX_INPUTS];
    int n;
} ogn_small_inputs;
int ogn_hier_step_small(const ogn_small_phase *phases_dev, int nphases, const ogn_small_inputs *inputs_host, int members,
                        int blocks_per_member, int threads, int wide, unsigned *counters_dev, void *stream);
int ogn_hier_small_max_blocks(int threads);

/* K2 — SparseCoder::forward (SparseCoder.cpp:13-43) + multiplyOHVs (SparseMatrix.cpp:260-276) over hidden x in
 * [x0,x1): gather-sum and first-max arg-max. hidden_cs2 (optional) receives a second copy: the next layer's history
 * slot (Hierarchy.cpp:239-246, K5). */
int ogn_sc_forward(const ogn_fwd_args *args_host, int Hx, int Hy, int Hz, int x0, int x1, int batch, int *hidden_cs,
                   int *hidden_cs2, void *stream);

/* K3 — SparseCoder::learn (SparseCoder.cpp:45-83) + multiplyOHVsT/countT/deltaChangedOHVsT (SparseMatrix.cpp:278-294,
 * 215-221, 572-590) for the visible layers in args (same geometry) and visible x in [vx0,vx1). update_count (optional)
 * is incremented by the number of (hidden column, visible column) pairs whose Vz weights were rewritten.
 * With Vz in {8,16,32} and a work list in args this is two launches: the reconstruction + arg-max over every visible
 * column (a pure 128-bit gather), then the delta rule over the listed columns only.
 * phase selects what is queued: OGN_LEARN_ALL, or the two halves separately (so that a caller can time them). When
 * ogn_sc_learn_is_split() is 0 the fused kernel runs in the RECONSTRUCT phase and UPDATE queues nothing. */
#define OGN_LEARN_ALL 0
#define OGN_LEARN_RECONSTRUCT 1
#define OGN_LEARN_UPDATE 2
int ogn_sc_learn(const ogn_learn_args *args_host, const int *hidden_cs, const int *hidden_cs_prev, int vx0, int vx1,
                 float alpha, int batch, unsigned long long *update_count, int phase, void *stream);
int ogn_sc_learn_is_split(const ogn_learn_args *args_host);

/* K6+K7+K8 — Predictor::learn (Predictor.cpp:49-70, deltaOHVs SparseMatrix.cpp:488-501), then Predictor::forward
 * (Predictor.cpp:13-47), then inputCs -> inputCsPrev (Predictor.cpp:118-126), over hidden x in [x0,x1).
 * learn != 0 requires target_cs. activations: [member][Hx][Hy][Hz] in/out (read by learn, rewritten by forward).
 * activate: 0 = learn only; 1 = activate and copy the inputs (all of them, whatever [x0,x1) is) to cs_prev_out;
 * 2 = activate without the copy (a launch over part of the rows whose sibling launch does the copy). */
int ogn_pred_step(const ogn_pred_args *args_host, int Hx, int Hy, int Hz, int x0, int x1, int batch, int learn, int activate,
                  const int *target_cs, float alpha, float *activations, int *hidden_cs, void *stream);

/* K9 — Actor::forward (Actor.cpp:13-90). u01[col] is the host-replayed std::generate_canonical<float,24> draw of the
 * column's tile sub-generator. */
int ogn_actor_forward(const ogn_actor_args *args_host, int Hx, int Hy, int Hz, const float *u01, float *hidden_values,
                      float *hidden_activations, int *hidden_cs, void *stream);

/* K11 — Actor::learn (Actor.cpp:92-185) on batch->n history pairs, applied one after the other inside one launch (each
 * action column owns its weights, so the per-sample launches of the reference collapse without changing the result).
 * args.vl[v].cs is ignored; the samples carry the CSDR pointers. */
int ogn_actor_learn(const ogn_actor_args *args_host, const ogn_actor_batch *batch_host, int Hx, int Hy, int Hz,
                    const float *hidden_values, float alpha, float beta, int mimic, float *hidden_activations, void *stream);

/* ImageEncoder (ImageEncoder.cpp:19-94): dense fp32 inputs, weights in the F layout (every visible cell of
 * every slot is read, so a column pack streams its block front to back). */
#define OGN_MAX_IENC_VL 8
typedef struct ogn_ienc_vl {
    float *wf;                    /* F layout virtual base */
    const float *in;              /* dense input [Vx][Vy][Vz] (address3, Helpers.h:247-252) */
    float *recon;                 /* reconstructions [Vx][Vy][Vz] */
    int geo, pad;
} ogn_ienc_vl;
typedef struct ogn_ienc_args {
    ogn_rf geo[OGN_MAX_GEO];
    ogn_ienc_vl vl[OGN_MAX_IENC_VL];
    int nvl;
} ogn_ienc_args;
/* ImageEncoder::forward for every hidden column (ImageEncoder.cpp:19-74): activation = -sum of squared distances
 * (SparseMatrix::distance2 :122-137), first-max arg-max -> hidden_cs; with learn != 0 the cells of a column are ranked by
 * activation (descending; equal activations in the order libstdc++'s std::sort leaves, ogn_stdsort.h), strength = expf(-rank^2 * gamma / max(0.001, resource)) * resource, resource -= alpha *
 * strength, and every weight of the cell moves towards its input by strength (SparseMatrix::hebb :692-701). Hz <= 32. */
int ogn_image_encoder_step(const ogn_ienc_args *args_host, int Hx, int Hy, int Hz, int learn, float alpha, float gamma, float *resources,
                  int *hidden_cs, void *stream);
/* ImageEncoder::backward for every cell of visible layer vli (ImageEncoder.cpp:76-94): the weights of the active cells of the
 * covering hidden columns, summed in ascending column order, divided by max(1, number of covering columns). */
int ogn_image_encoder_reconstruct(const ogn_ienc_args *args_host, int vli, int Hx, int Hy, int Hz, const int *hidden_cs, void *stream);

/* Peer-memory all-gather of a sharded CSDR (multi-GPU, one process per GPU): base[r] is rank r's exchange arena as mapped
 * in THIS process (CUDA IPC); the buffer lives at the same offset in every arena. One launch: this rank's slab
 * [rank*count, (rank+1)*count) is stored into every peer's arena with 128-bit NVLink stores, the last block to finish
 * raises this rank's arrival flag (arena word `rank`, value epoch) in every peer with a system-scope release store and
 * waits until the flag of every peer in wait_mask (bit r = rank r) reached epoch in the local arena (acquire loads, bounded
 * by timeout_ns; on timeout *err_flag, a host-mapped word, is set to 1). Kernels queued behind it on the stream see the
 * slabs of the ranks waited for; with a partial mask (the x-neighbours whose halo rows the next kernels read) the other
 * slabs are guaranteed complete only after a later exchange with the full mask: flags only grow, and a rank raises
 * exchange k + 1 after it delivered exchange k. */
#define OGN_MAX_PEERS 16
typedef struct ogn_peer_table {
    int *base[OGN_MAX_PEERS];
    int rank, world;
} ogn_peer_table;
int ogn_peer_allgather(const ogn_peer_table *table_host, int64_t offset_ints, int count_per_rank, unsigned epoch, unsigned wait_mask,
                       unsigned long long timeout_ns, unsigned *err_flag, void *stream);

/* Weight (re)layout between the reference's CSR value order and the device layouts, for hidden columns
 * [col0, col1) (col = oy + ox*Hy) of one ensemble member. csr is a DEVICE staging buffer holding CSR-order values;
 * csr[0] is the value with global CSR index csr_base (normally the CSR block offset of column col0). which: 0 = F,
 * 1 = R. Padding elements of the packed layouts are never written (storage is zero-initialised).
 * wset_host / rf_host are HOST structs (rf_host's table pointers are device pointers). */
int ogn_weights_from_csr(const ogn_wset *wset_host, const ogn_rf *rf_host, int member, int which, int col0, int col1,
                         const float *csr, int64_t csr_base, void *stream);
int ogn_weights_to_csr(const ogn_wset *wset_host, const ogn_rf *rf_host, int member, int which, int col0, int col1,
                       float *csr, int64_t csr_base, void *stream);

/* Device-side counter-based init for sizes the reference cannot construct: every weight is
 * lo + (hi-lo)*u, u = hash(seed, stream_id, CSR index of the weight)/2^24, written to the F (and R) layout. */
int ogn_weights_init_hash(const ogn_wset *wset_host, const ogn_rf *rf_host, int member, uint64_t seed, uint32_t stream_id,
                          float lo, float hi, void *stream);

/* Number of elements where the F and R copies of a SparseCoder weight set disagree (debug / tests). */
int ogn_weights_fr_mismatch(const ogn_wset *wset_host, const ogn_rf *rf_host, int member, unsigned long long *count_dev,
                            void *stream);

/* Launch counts per (kernel, variant) since the library was loaded: out[k*8 + c], k = 0 SparseCoder forward, 1 SparseCoder
 * reconstruction, 2 Predictor step; c = 0 generic one-cell-per-lane, 1 vec1, 2 vec2, 3 pipe4, 4 pipe6, 5 tma, 6 wide, 7 unused.
 * Tests use it to assert which kernel the default selector ran at a given size. */
void ogn_debug_variant_launches(long long out[24]);
/* debug timeline of the tile-staged ImageEncoder kernel: see ogn_kernels.cu */
int ogn_debug_ienc_trace(int blocks, long long *out);
/* debug timeline of the single-launch step (ogn_hier_step_small): see ogn_kernels.cu */
int ogn_debug_small_trace(int arm, unsigned long long *out);

/* Device replays of glibc expf/tanhf (csrc/ogn_libm.h), exposed for tests: out[i] = f(in[i]). */
int ogn_test_expf(const float *in, float *out, int64_t n, void *stream);
int ogn_test_tanhf(const float *in, float *out, int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OGN_KERNELS_API_H */

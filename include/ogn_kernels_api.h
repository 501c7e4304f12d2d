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
 *   weights, "F"    fp32  [hidden column][slot][visible cell vz][hidden cell hz]   (row gather: hz contiguous)
 *   weights, "R"    fp32  [hidden column][slot][hidden cell hz][visible cell vz]   (SparseCoder reconstruction:
 *                                                                                   vz contiguous); SparseCoder only
 *   slot = (ix - lox)*ny + (iy - loy) inside the clamped receptive field, exactly the reference's CSR order
 *   (Helpers.cpp:236-313); a hidden column's block starts at Vz*Hz*(px[ox]*sy + nx[ox]*py[oy]) in all layouts, so no
 *   index arrays are stored and every offset is 64-bit.
 */
#ifndef OGN_KERNELS_API_H
#define OGN_KERNELS_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OGN_MAX_VL 32 /* visible layers per launch (inputs x temporalHorizon of one SparseCoder) */

/* Receptive-field tables of one (hidden grid, visible grid, radius) pairing: initSMLocalRF in closed form
 * (Helpers.cpp:236-313, project() Helpers.h:223-228). The struct and the tables it points to live in device memory. */
typedef struct ogn_rf {
    int Hx, Hy, Hz;           /* hidden (output) grid */
    int Vx, Vy, Vz;           /* visible (input) grid */
    int radius;
    int sy;                   /* sum of ny[] over all hidden y */
    const int *lox, *nx, *px; /* [Hx] first visible x of the field, extent, exclusive prefix sum of nx */
    const int *loy, *ny, *py; /* [Hy] same along y */
    const int *rlox, *rhix;   /* [Vx] first / last hidden x whose field covers visible x (rlox > rhix: none) */
    const int *rloy, *rhiy;   /* [Vy] */
} ogn_rf;

/* The weights of one visible layer of one layer object (static after init; an array of these lives in device memory). */
typedef struct ogn_wset {
    const ogn_rf *rf;         /* device pointer */
    float *wf;                /* F layout, hidden x in [fx0, fx1) */
    float *wr;                /* R layout, hidden x in [rx0, rx1); NULL unless SparseCoder */
    float *recon;             /* SparseCoder: reconstructions [member][Vx][Vy][Vz] (SparseCoder.h:35); else NULL */
    int64_t wf_base, wr_base; /* block offset of column (fx0,0) / (rx0,0): subtracted from global block offsets */
    int64_t wf_mstride, wr_mstride; /* floats between consecutive ensemble members */
    int fx0, fx1, rx0, rx1;
    int vx0, vx1;             /* SparseCoder learn: visible x range this device is responsible for */
} ogn_wset;

/* Per-launch CSDR pointers, passed by value. */
typedef struct ogn_cs_args {
    const int *cs[OGN_MAX_VL];      /* visible CSDR of visible layer v */
} ogn_cs_args;
typedef struct ogn_pred_args {
    const int *cs[8];               /* current inputs (Predictor::activate argument) */
    const int *cs_prev[8];          /* VisibleLayer::inputCsPrev (Predictor.h:36) */
    int *cs_prev_out[8];            /* where this launch stores cs for the next learn (ping-pong buffer) or NULL */
} ogn_pred_args;

const char *ogn_cuda_error_string(int code);

/* K2 — SparseCoder::forward (SparseCoder.cpp:13-43) + multiplyOHVs (SparseMatrix.cpp:260-276) over hidden x in
 * [x0,x1): gather-sum and first-max arg-max. hidden_cs2 (optional) receives a second copy: the next layer's history
 * slot (Hierarchy.cpp:239-246, K5). */
int ogn_sc_forward(const ogn_wset *wsets, int nvl, const ogn_cs_args *cs_host, int Hx, int Hy, int Hz, int x0, int x1,
                   int batch, int *hidden_cs, int *hidden_cs2, void *stream);

/* K3 — SparseCoder::learn (SparseCoder.cpp:45-83) + multiplyOHVsT/countT/deltaChangedOHVsT (SparseMatrix.cpp:278-294,
 * 215-221, 572-590) for visible layers [v0, v0+nv) which must share (Vx,Vy,Vz). update_count (optional) is incremented
 * by the number of (hidden column, visible column) pairs whose Vz weights were rewritten. */
int ogn_sc_learn(const ogn_wset *wsets, int v0, int nv, const ogn_cs_args *cs_host, const int *hidden_cs,
                 const int *hidden_cs_prev, int Hx, int Hy, int Hz, int Vx, int Vy, int Vz, int vx0, int vx1, float alpha,
                 int batch, unsigned long long *update_count, void *stream);

/* K6+K7+K8 — Predictor::learn (Predictor.cpp:49-70, deltaOHVs SparseMatrix.cpp:488-501), then Predictor::forward
 * (Predictor.cpp:13-47), then inputCs -> inputCsPrev (Predictor.cpp:118-126), over hidden x in [x0,x1).
 * learn != 0 requires target_cs. activations: [member][Hx][Hy][Hz] in/out (read by learn, rewritten by forward). */
int ogn_pred_step(const ogn_wset *wsets, int nvl, const ogn_pred_args *args_host, int Hx, int Hy, int Hz, int x0, int x1,
                  int batch, int learn, int activate, const int *target_cs, float alpha, float *activations,
                  int *hidden_cs, void *stream);

/* K9 — Actor::forward (Actor.cpp:13-90). value_wsets/action_wsets: nvl entries each. u01[col] is the host-replayed
 * std::generate_canonical<float,24> draw of the column's tile sub-generator. */
int ogn_actor_forward(const ogn_wset *value_wsets, const ogn_wset *action_wsets, int nvl, const ogn_pred_args *args_host,
                      int Hx, int Hy, int Hz, const float *u01, float *hidden_values, float *hidden_activations,
                      int *hidden_cs, void *stream);

/* K11 — Actor::learn (Actor.cpp:92-185) on one history pair. args.cs = sPrev.inputCs. */
int ogn_actor_learn(const ogn_wset *value_wsets, const ogn_wset *action_wsets, int nvl, const ogn_pred_args *args_host,
                    int Hx, int Hy, int Hz, const int *target_cs_prev, const float *hidden_values_prev,
                    const float *hidden_values, float q, float g, float alpha, float beta, int mimic,
                    float *hidden_activations, void *stream);

/* Weight (re)layout between the reference's CSR value order and the device layouts, for hidden columns
 * [col0, col1) (col = oy + ox*Hy) of one ensemble member. csr is a DEVICE staging buffer holding CSR-order values;
 * csr[0] is the value with global CSR index csr_base (normally the block offset of column col0). which: 0 = F, 1 = R.
 * wset_host / rf_host are HOST copies of the descriptors (their table pointers are device pointers). */
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

/* Device replays of glibc expf/tanhf (csrc/ogn_libm.h), exposed for tests: out[i] = f(in[i]). */
int ogn_test_expf(const float *in, float *out, int64_t n, void *stream);
int ogn_test_tanhf(const float *in, float *out, int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OGN_KERNELS_API_H */

/* ogn_image_encoder_api.h — flat C ABI of the ImageEncoder pre-encoder: the step immediately BEFORE
 * the Sparse Predictive Hierarchy path for dense float inputs (images): every hidden column picks the cell whose weight
 * vector is nearest (squared distance) to the dense input patch in its receptive field, and learns with a rank-gated,
 * resource-limited Hebbian rule; reconstruct() maps a hidden CSDR back to dense inputs.
 *
 * Replaces ogmaneo::ImageEncoder (source/ogmaneo/ImageEncoder.h:18-154, ImageEncoder.cpp:19-225). Exported three times like
 * the hierarchy ABI: ogn_ (CUDA product, no CPU fallback), oport_ (oracle port) and oref_ (compiled reference); the last
 * two are test infrastructure. Plain pointers and sizes; all buffers are HOST buffers. Dense inputs are fp32, cell
 * (x, y, z) of a visible layer at index z + size_z * (y + size_y * x) (address3, Helpers.h:247-252). Functions return 0 on
 * success unless documented to return a count.
 */
#ifndef OGN_IMAGE_ENCODER_API_H
#define OGN_IMAGE_ENCODER_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ImageEncoder::VisibleLayerDesc, ImageEncoder.h:21-33 (defaults 4,4,16 / 2) */
typedef struct ogn_ienc_visible_desc {
    int size_x, size_y, size_z;
    int radius;
} ogn_ienc_visible_desc;

#define OGN_DECLARE_IMAGE_ENCODER_API(P)                                                                            \
    /* ComputeSystem cs; cs.rng.seed(seed); ImageEncoder::initRandom(cs, hiddenSize, descs) — ImageEncoder.cpp:96-137 */ \
    int P##ienc_create(int hidden_x, int hidden_y, int hidden_z, int num_visible, const ogn_ienc_visible_desc *descs,   \
                       uint32_t seed, void **out);                                                                      \
    void P##ienc_destroy(void *e);                                                                                      \
    const char *P##ienc_last_error(void);                                                                               \
    /* ImageEncoder::step, ImageEncoder.cpp:139-148 (forward :19-74): input_acts[v] holds the dense input of visible layer v */ \
    int P##ienc_step(void *e, const float *const *input_acts, int learn_enabled);                                       \
    /* ImageEncoder::reconstruct, ImageEncoder.cpp:150-160 (backward :76-94): hidden_cs = one cell index per hidden column */ \
    int P##ienc_reconstruct(void *e, const int *hidden_cs);                                                             \
    /* getHiddenCs ImageEncoder.h:140-142; hiddenResources is private upstream (ImageEncoder.h:41), exposed for parity;  \
       getVisibleLayer(v).reconstructions / .weights.nonZeroValues ImageEncoder.h:123-137. Return the element count. */  \
    int P##ienc_hidden_cs(void *e, int *out);                                                                           \
    int P##ienc_hidden_resources(void *e, float *out);                                                                  \
    int P##ienc_reconstruction(void *e, int vli, float *out);                                                           \
    int64_t P##ienc_weights(void *e, int vli, float *out, int64_t cap);                                                 \
    /* public members alpha (resource depletion rate) and gamma (rank falloff), ImageEncoder.h:94-104 */               \
    int P##ienc_set_params(void *e, float alpha, float gamma);                                                          \
    /* ImageEncoder::writeToStream / readFromStream on a file, ImageEncoder.cpp:162-225; save returns bytes or -1 */     \
    int64_t P##ienc_save(void *e, const char *path);                                                                    \
    int P##ienc_load(const char *path, void **out);

OGN_DECLARE_IMAGE_ENCODER_API(ogn_)

/* ---- device-resident extensions (CUDA product only; absent upstream) ----
 * ogn_ienc_create_on: like ogn_ienc_create, but the encoder runs on the device and stream of `hierarchy` (a handle of
 * include/ogn_hierarchy_api.h / ogn_b200.h, which must outlive the encoder); it keeps its own generator (seed).
 * ogn_ienc_step_device: ImageEncoder::step on dense inputs that already live in device memory; asynchronous on that stream.
 * ogn_ienc_hidden_cs_device: device pointer of the hidden CSDR (int32 per hidden column), valid for the life of the handle —
 * hand it to ogn_step_device of the same hierarchy: image -> CSDR -> Hierarchy::step without a host round trip.
 * ogn_ienc_synchronize: wait for the encoder's stream. */
int ogn_ienc_create_on(void *hierarchy, int hidden_x, int hidden_y, int hidden_z, int num_visible, const ogn_ienc_visible_desc *descs,
                       uint32_t seed, void **out);
int ogn_ienc_step_device(void *e, const float *const *input_acts_dev, int learn_enabled);
const int *ogn_ienc_hidden_cs_device(void *e);
int ogn_ienc_synchronize(void *e);

#ifdef __cplusplus
}
#endif
#endif /* OGN_IMAGE_ENCODER_API_H */

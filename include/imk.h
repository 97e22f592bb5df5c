/*
 * imk.h -- C ABI of the B200-native Inconsistency-Mask pseudo-labelling hot path.
 *
 * The reference (MichaelVorndran/InconsistencyMasks) is pure Python and has no FFI
 * of its own; the boundary this library sits behind is the set of Python helpers
 * in the reference's functions.py.  Every entry point below names the reference
 * interface it replaces (file:line relative to the reference repository root).
 * inconsistencymasks_b200/functions.py binds these symbols with ctypes and keeps
 * the reference's helper names / argument meaning; INTEGRATION.md shows the stub
 * a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no exceptions across the boundary.
 *   - every function returns 0 on success, a negative IMK_E* code otherwise;
 *     imk_last_error() returns a thread-local description of the last failure.
 *   - the caller owns every buffer passed in.  Pointers named *_dev are CUDA
 *     device pointers on the current device, *_host are host pointers (pinned
 *     memory gives asynchronous copies; pageable works but serialises).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     Device-buffer calls are asynchronous on that stream; *_host calls return
 *     after their results are in the host buffers.
 *   - layouts follow the reference: images uint8 NHWC [N,H,W,c] exactly as
 *     cv2.imread returns them (BGR on disk order for c == 3), probabilities
 *     float32 NHWC [N,H,W,K], labels / IM uint8 [N,H,W], sizes int64 [N].
 *   - there is no CPU fallback: without a CUDA device every compute call fails
 *     with IMK_ECUDA.
 */
#ifndef IMK_H_
#define IMK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMK_VERSION 200

#define IMK_OK        0
#define IMK_EINVAL   -1   /* bad argument (shape, NULL pointer, unsupported K / M) */
#define IMK_ECUDA    -2   /* a CUDA runtime call or kernel launch failed          */
#define IMK_ENOMEM   -3   /* device or host allocation failed                     */
#define IMK_ESTATE   -4   /* handle used in a way its state does not allow        */

#define IMK_ACT_SIGMOID 0   /* config.ini ACTIFU_OUTPUT = sigmoid (ISIC_2018, HELA) */
#define IMK_ACT_SOFTMAX 1   /* config.ini ACTIFU_OUTPUT = softmax (SUIM, CITYSCAPES) */

#define IMK_IN_U8   0       /* model.predict on the uint8 array of functions.py:2852 */
#define IMK_IN_F32  1       /* model.predict on a float32 array (benchmark_hela, functions.py:1199) */

#define IMK_MAX_MODELS 16
#define IMK_MAX_CLASSES 256

int         imk_version(void);
const char *imk_last_error(void);
/* Number of kernel launches this thread issued through the library so far. */
int64_t     imk_launch_count(void);
/* Images per internal trunk pass (the workspace of a model is sized for this many; env IMK_CHUNK overrides). */
int64_t     imk_max_chunk(void);
/* Change it (1..1024; 0 restores the default).  Workspaces grow on demand; results do not depend on it. */
int         imk_set_max_chunk(int64_t n);
/* 1 when a CUDA device is usable from this process, else 0 (never throws). */
int         imk_device_available(void);

/* Per-kernel device timing for the roofline report (bench.py).  Between begin and end
 * every kernel the calling thread launches through the library is bracketed by CUDA
 * events on its own stream; end synchronises the device and returns one entry per
 * (kernel name, tag) with the launch count and the summed device time.  `tag` is the
 * index of the U-Net layer (0..23, creation order of unet.py) for convolution kernels,
 * -1 otherwise.  n_out receives the number of distinct entries (may exceed cap). */
typedef struct imk_profile_entry {
    char    name[48];
    int     tag;
    int64_t launches;
    double  total_ms;
} imk_profile_entry;
int imk_profile_begin(void);
int imk_profile_end(imk_profile_entry *out, int cap, int *n_out);

/* ------------------------------------------------------------------------- *
 *  Row a5 / a6 : pred_masks_to_im_binary  (functions.py:3104-3120)
 *                pred_masks_to_im_multiclass (functions.py:3123-3137)
 *  masks_dev: int64 [M, P] (P = pixels of ONE image), any integer values --
 *  the binary helper is a SUM over models, the multiclass one an all-equal test.
 *  Outputs: label/im uint8 [P]; sizes_dev int64 [2] = {im_size, pred_size}
 *  (multiclass writes im_size only).
 * ------------------------------------------------------------------------- */
int imk_masks_to_im_binary(const int64_t *masks_dev, int M, int64_t P,
                           uint8_t *label_dev, uint8_t *im_dev, int64_t *sizes_dev,
                           void *stream);
int imk_masks_to_im_multiclass(const int64_t *masks_dev, int M, int64_t P,
                               uint8_t *label_dev, uint8_t *im_dev, int64_t *sizes_dev,
                               void *stream);

/* ------------------------------------------------------------------------- *
 *  Rows a2 / a3 + a7 / a9 : get_im_prediction_binary (functions.py:3140-3162),
 *  get_im_prediction_hela (functions.py:3165-3202) with model.predict replaced
 *  by its output, fused with the blanking of functions.py:2867-2874 / 2968-2974.
 *
 *  probs_dev   host array of M device pointers, each float32 [N,H,W,K], K = 1 (ISIC)
 *              or 3 (HeLa heads alive, dead, position).
 *  strict_gt   1: prob >  thr (functions.py:3157)   0: prob >= thr (functions.py:3187-3189)
 *  img_dev     uint8 [N,H,W,c] to blank (may be NULL when block_in == 0 or img_out_dev == NULL)
 *  labels_dev  uint8 [K][N,H,W] (head-major planes): 255 where all models fire
 *  im_dev      uint8 [N,H,W]: 255 where models disagree (K == 3: max over heads)
 *  im_size_dev int64 [N]: K == 1 pixels of the IM; K == 3 the SUM of the three head
 *              IM sizes (functions.py:3200), counted before blanking
 *  pred_size_dev int64 [K][N] or NULL: pixels where all models fire
 *  block_in    write img_out = im ? 0 : img          block_out: label = im ? 0 : label
 *              (K == 1: a no-op without morphology, the label is already 0 there;
 *              K == 3: clears alive / dead where ANY head disagrees, functions.py:2972-2973;
 *              head 2 (position) is always returned raw -- the reference blanks the circle
 *              image the host draws from it, functions.py:2953-2965, 2974)
 * ------------------------------------------------------------------------- */
int imk_im_binary(const float *const *probs_dev, int M, int64_t N, int H, int W, int K,
                  float thr, int strict_gt,
                  const uint8_t *img_dev, int c, int block_in, int block_out,
                  uint8_t *img_out_dev, uint8_t *labels_dev, uint8_t *im_dev,
                  int64_t *im_size_dev, int64_t *pred_size_dev, void *stream);

/* ------------------------------------------------------------------------- *
 *  Rows a4 + a8 : get_im_prediction_multiclass (functions.py:3206-3238) fused with
 *  the blanking of functions.py:3054-3061.  argmax = first index of the maximum,
 *  NaN counts as the maximum (np.argmax, functions.py:3225).
 *  label_dev   uint8 [N,H,W] class id where all models agree else 0
 *  im_dev      uint8 [N,H,W] 255 where they do not
 *  lists_equal_dev uint8 [N] or NULL: 1 when every model predicts the same SET of
 *              classes over the image (functions.py:3226-3234); needs K <= 64.
 * ------------------------------------------------------------------------- */
int imk_im_multiclass(const float *const *probs_dev, int M, int64_t N, int H, int W, int K,
                      const uint8_t *img_dev, int c, int block_in, int block_out,
                      uint8_t *img_out_dev, uint8_t *label_dev, uint8_t *im_dev,
                      int64_t *im_size_dev, uint8_t *lists_equal_dev, void *stream);

/* ------------------------------------------------------------------------- *
 *  Optional morphology of rows a7-a9 (off in config.ini: ERODE_KERNEL = DILATE_KERNEL = 0).
 *  cv2.erode / cv2.dilate(mask, ones((k,k)), iterations=1) on uint8 [N,H,W]
 *  (functions.py:2858-2864); dilate_mask == 3x3 dilate of the label map
 *  (functions.py:3075-3100).  src and dst must not alias.
 * ------------------------------------------------------------------------- */
int imk_erode_u8(const uint8_t *src_dev, uint8_t *dst_dev, int64_t N, int H, int W, int k, void *stream);
int imk_dilate_u8(const uint8_t *src_dev, uint8_t *dst_dev, int64_t N, int H, int W, int k, void *stream);
/* image[im > 0] = 0 over c channels and n_labels label planes [n_labels][N,H,W]
 * (functions.py:2867-2874, 2968-2974, 3054-3061).  In place.  Either may be NULL. */
int imk_blank(const uint8_t *im_dev, int64_t N, int H, int W,
              uint8_t *img_dev, int c, uint8_t *labels_dev, int n_labels, void *stream);

/* ------------------------------------------------------------------------- *
 *  Row a1 : get_unet (unet.py:46-67) forward, i.e. model.predict
 *  (functions.py:3157, 3184, 3224).
 *
 *  weights: the 104 float32 arrays of Keras model.get_weights() in creation order
 *  (conv kernel HWIO + bias; BatchNormalization gamma, beta, moving_mean,
 *  moving_variance), SURVEY.md appendix C.  Host pointers; copied and repacked.
 * ------------------------------------------------------------------------- */
typedef struct imk_unet imk_unet_t;

typedef struct imk_unet_desc {
    int   height, width;      /* unet.py:47 i_height, i_width (multiples of 16)        */
    int   in_channels;        /* i_channels: 1 or 3                                     */
    int   num_outputmasks;    /* K                                                      */
    float alpha;              /* width multiplier, channels = int(k * alpha)            */
    int   ks;                 /* 3 (unet.py:46 default; 1 and 3 supported)              */
    int   act_out;            /* IMK_ACT_SIGMOID / IMK_ACT_SOFTMAX                      */
    int   swap_rb;            /* 1: feed channel 2-i of the image buffer (the reference
                                 feeds cvtColor(BGR2RGB) of the array it later blanks,
                                 functions.py:2846-2852); 0: feed as is                 */
} imk_unet_desc;

int  imk_unet_create(const imk_unet_desc *desc, const float *const *weights_host,
                     const int64_t *weight_sizes, int n_weights, imk_unet_t **out);
void imk_unet_destroy(imk_unet_t *net);
int  imk_unet_param_count(const imk_unet_t *net, int64_t *count);
/* Selects the convolution engine for the layers that have a tensor-core path:
 * 0 = shared-memory-tiled direct convolutions everywhere, 1 = tcgen05 implicit GEMM
 * where Cin is wide enough (default).  Both are CUDA; used for A/B measurements. */
int  imk_unet_set_engine(imk_unet_t *net, int engine);
/* Changes desc.swap_rb after creation.  The flag belongs to imk_unet_forward / imk_unet_predict_host only;
 * the ensemble and host-pipeline calls below take their own per-call swap_rb and never touch it. */
int  imk_unet_set_swap_rb(imk_unet_t *net, int swap_rb);

/* probs_dev float32 [N,H,W,K].  images_dev: uint8 or float32 [N,H,W,c] per in_dtype. */
int imk_unet_forward(imk_unet_t *net, const void *images_dev, int in_dtype, int64_t N,
                     float *probs_dev, void *stream);
/* Same with host buffers (what model.predict does): copies in, runs, copies out. */
int imk_unet_predict_host(imk_unet_t *net, const void *images_host, int in_dtype, int64_t N,
                          float *probs_host);

/* ------------------------------------------------------------------------- *
 *  Fused fast path: ensemble forward + a2/a3/a4 + a7/a8/a9 with the fp32
 *  probability maps never written to HBM.  Same outputs as imk_im_binary /
 *  imk_im_multiclass fed with imk_unet_forward's probabilities, bit for bit.
 *  All models must share height, width, in_channels, num_outputmasks, act_out.
 *  swap_rb (per call, in_channels == 3): 1 = the models are fed channel 2-i of images_dev,
 *  i.e. cv2.cvtColor(image, COLOR_BGR2RGB) of functions.py:2847-2852, while img_out keeps the
 *  buffer's own (BGR) order like the array the reference blanks (functions.py:2867-2874).
 * ------------------------------------------------------------------------- */
int imk_ensemble_im_binary(imk_unet_t *const *nets, int M, const uint8_t *images_dev, int64_t N, int swap_rb,
                           float thr, int strict_gt, int block_in, int block_out,
                           uint8_t *img_out_dev, uint8_t *labels_dev, uint8_t *im_dev,
                           int64_t *im_size_dev, int64_t *pred_size_dev, void *stream);
int imk_ensemble_im_multiclass(imk_unet_t *const *nets, int M, const uint8_t *images_dev, int64_t N, int swap_rb,
                               int block_in, int block_out,
                               uint8_t *img_out_dev, uint8_t *label_dev, uint8_t *im_dev,
                               int64_t *im_size_dev, uint8_t *lists_equal_dev, void *stream);

/* ------------------------------------------------------------------------- *
 *  The per-directory loops of create_pseudo_labels_im_ISIC_2018 / _hela /
 *  _multiclass (functions.py:2844-2887, 2932-2980, 3020-3066) minus PNG I/O, on
 *  HOST buffers: images are streamed to the device in chunks on two streams so
 *  that copies overlap compute, results are streamed back.  erode_kernel ==
 *  dilate_kernel == 0 only (the config.ini defaults); use the device-buffer calls
 *  for morphology.  Outputs as in the device calls above; any output may be NULL.
 *  chunk > 0: images per chunk, as given.  chunk <= 0: the library picks -- up to
 *  imk_max_chunk() images, at least four chunks, and a quarter and a half chunk at
 *  both ends (the first upload and the last download are the only copies nothing
 *  overlaps).  Results do not depend on the chunking.
 * ------------------------------------------------------------------------- */
int imk_pseudo_label_binary_host(imk_unet_t *const *nets, int M, const uint8_t *images_host, int64_t N, int swap_rb,
                                 float thr, int strict_gt, int block_in, int block_out,
                                 uint8_t *img_out_host, uint8_t *labels_host, uint8_t *im_host,
                                 int64_t *im_size_host, int64_t *pred_size_host, int64_t chunk);
int imk_pseudo_label_multiclass_host(imk_unet_t *const *nets, int M, const uint8_t *images_host, int64_t N, int swap_rb,
                                     int block_in, int block_out,
                                     uint8_t *img_out_host, uint8_t *label_host, uint8_t *im_host,
                                     int64_t *im_size_host, uint8_t *lists_equal_host, int64_t chunk);

/* Opt-in compact result layout of the two calls above (same arguments, same statistics): the 0/255 planes -- the K label
 * planes of the binary / HeLa path and the IM -- come back as bits (imk_pack_bits: 8 pixels per byte, pixel i of a group
 * in bit i; plane k of the labels starts at byte k*N*H*W/8), a multiclass label keeps its class-id bytes.  img_out_host
 * may be NULL: blanking is `image[im > 0] = 0` (functions.py:2867), which a host that still holds the image it uploaded can
 * apply itself -- then 1/8 of a byte per mask pixel crosses PCIe instead of c + planes + 1 bytes.  H*W % 8 == 0. */
int imk_pseudo_label_binary_host_packed(imk_unet_t *const *nets, int M, const uint8_t *images_host, int64_t N, int swap_rb,
                                        float thr, int strict_gt, int block_in, int block_out,
                                        uint8_t *img_out_host, uint8_t *label_bits_host, uint8_t *im_bits_host,
                                        int64_t *im_size_host, int64_t *pred_size_host, int64_t chunk);
int imk_pseudo_label_multiclass_host_packed(imk_unet_t *const *nets, int M, const uint8_t *images_host, int64_t N, int swap_rb,
                                            int block_in, int block_out,
                                            uint8_t *img_out_host, uint8_t *label_host, uint8_t *im_bits_host,
                                            int64_t *im_size_host, uint8_t *lists_equal_host, int64_t chunk);

/* ------------------------------------------------------------------------- *
 *  Adjacent components, on data that is already in HBM.
 * ------------------------------------------------------------------------- */

/* 8f-2: integer confusion counts behind benchmark_ISIC2018 / _hela / _multiclass (functions.py:1078-1339); the
 * host forms get_IoU_binary (functions.py:1767), dice_score_numpy_binary (:1837), get_IoU_multi_unique (:1790) and
 * pixel_accuracy (:1819) from them with the reference's own expressions.
 * pred / gt: uint8 [N][hw] planes.  counts int64 [N][5] = |gt!=0 & pred!=0|, |gt!=0 | pred!=0|, |gt>=128 & pred>=128|,
 * |gt>=128|, |pred>=128|.  hist int64 [N][3][256] = per value v: |gt==v|, |pred==v|, |gt==v & pred==v|. */
int imk_seg_counts_binary(const uint8_t *pred_dev, const uint8_t *gt_dev, int64_t N, int64_t hw, int64_t *counts_dev, void *stream);
int imk_seg_counts_multiclass(const uint8_t *pred_dev, const uint8_t *gt_dev, int64_t N, int64_t hw, int64_t *hist_dev, void *stream);

/* Opt-in compact result layout for 0/255 planes (labels of the binary / HeLa paths, the IM): 8 pixels per byte,
 * pixel i of each group of 8 in bit i.  n_bytes (multiple of 8) input bytes -> n_bytes / 8 output bytes.
 * np.unpackbits(bits, bitorder="little") * 255 restores the planes exactly. */
int imk_pack_bits(const uint8_t *planes_dev, int64_t n_bytes, uint8_t *bits_dev, void *stream);

/* 8f-1: augment_image_and_mask(s) of the IM+ / IM++ scripts (functions.py:2725-2828, primitives :1463-1506) for a batch.
 * The host chooses the operations (the reference's unseeded random / np.random draws); the device computes the pixels:
 * flips + rotation on image and masks, convertScaleAbs, GaussianBlur 3/5/7 (sigma 0) and uniform noise on the image.
 * Everything but the noise is bit-exact against OpenCV for the same parameters. */
typedef struct imk_aug_params {
    int flip_v, flip_h;             /* cv2.flip(x, 0) / cv2.flip(x, 1), applied in that order */
    int rot;                        /* 0 none, 1 cv2.ROTATE_90_CLOCKWISE, 2 ROTATE_180, 3 ROTATE_90_COUNTERCLOCKWISE (1, 3: H == W) */
    int scale_on;                   /* cv2.convertScaleAbs(image, alpha, beta) */
    float alpha, beta;
    int blur_k;                     /* 0, 3, 5, 7: cv2.GaussianBlur(image, (k, k), 0) */
    int noise_max;                  /* > 0: image + randint(-noise_max, noise_max), clipped to 0..255 */
    uint64_t seed;                  /* of the noise generator (counter-based: same seed -> same pixels) */
} imk_aug_params;
/* img uint8 [N,H,W,c] (or NULL: masks only), masks uint8 [mask_planes][N,H,W] (mask_planes may be 0).
 * scratch_dev: N*H*W*c bytes, needed when any image has blur_k or noise_max > 0.  params_host: N entries. */
int imk_augment_u8(const uint8_t *img_dev, const uint8_t *masks_dev, int64_t N, int H, int W, int c, int mask_planes,
                   const imk_aug_params *params_host, uint8_t *img_out_dev, uint8_t *masks_out_dev,
                   uint8_t *scratch_dev, void *stream);

/* 8f-4: EvalNet forward of the IM++ scripts (evalnet.py:24-73; called functions.py:5733-5737, 6010-6017).
 * n_heads 1 = get_evalnet (Dense(1)), 2 = get_evalnet_miou (Dense(b_channels) 'iou' + 'detection').
 * b_onehot: input B is given as a uint8 class map [N,H,W] standing for its one-hot encoding over b_channels classes
 * (functions.py:6004-6006), else as uint8 [N,H,W,b_channels].  int(16 * alpha) must be a multiple of 16.
 * Weight order: see imk_evalnet_create in csrc/imk_evalnet.cu (layer creation order of evalnet.py). */
typedef struct imk_evalnet imk_evalnet_t;
typedef struct imk_evalnet_desc {
    int height, width, a_channels, b_channels;
    float alpha;
    int ks;                         /* 3 */
    int normalize_a, normalize_b;   /* x / 255 in the input block (evalnet.py:5-6) */
    int b_onehot;
    int n_heads;
} imk_evalnet_desc;
int  imk_evalnet_create(const imk_evalnet_desc *desc, const float *const *weights_host, const int64_t *weight_sizes,
                        int n_weights, imk_evalnet_t **out);
void imk_evalnet_destroy(imk_evalnet_t *net);
int  imk_evalnet_param_count(const imk_evalnet_t *net, int64_t *count);
/* out0 / out1: float32 [N][1 or b_channels]; out1 (the 'detection' head) only with n_heads == 2. */
int  imk_evalnet_forward(imk_evalnet_t *net, const uint8_t *a_dev, const uint8_t *b_dev, int64_t N, int swap_rb_a,
                         float *out0_dev, float *out1_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* IMK_H_ */

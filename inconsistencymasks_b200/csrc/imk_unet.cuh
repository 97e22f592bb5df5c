// Internal description of one packed U-Net (reference unet.py:46-67) on the device.
#pragma once
#include <vector>
#include "imk_common.cuh"
#include "imk_block_tc.cuh"

namespace imk {

constexpr int kChanPad = 16;            // activations are fp16 NHWC with channels padded to 16
inline int pad_ch(int c) { return (c + kChanPad - 1) / kChanPad * kChanPad; }

// One Conv2D(+ReLU)(+BatchNormalization) of the plan, packed for the kernels.
struct ConvLayer {
    int ks = 1, cin = 0, cout = 0, cin_p = 0, cout_p = 0;
    bool has_bn = false;
    __half *w_direct = nullptr;   // [ks*ks][cin_p][cout_p] fp16  (direct convolution engine)
    __half *w_umma = nullptr;     // tcgen05 operand-B image, see imk_conv_tc.cu
    float *w_f32 = nullptr;       // [ks*ks][cin][cout] fp32 (first / last layer only)
    float *bias = nullptr;        // [cout_p] fp32, zero in the padding
    float *bn_scale = nullptr;    // [cout_p] gamma / sqrt(var + eps)   (0 in the padding)
    float *bn_shift = nullptr;    // [cout_p] beta - mean * scale        (0 in the padding)
};

struct Level {                    // buffers of one resolution level, sized for `cap_n` images
    __half *skip = nullptr, *a = nullptr, *b = nullptr;
    int h = 0, w = 0, ch_p = 0;
};

}  // namespace imk

struct imk_unet {
    imk_unet_desc desc{};
    int widths[5] = {0, 0, 0, 0, 0};           // int(16a), int(32a), int(64a), int(128a), int(256a)
    std::vector<imk::ConvLayer> conv;           // 24 layers in creation order (unet.py:49-63)
    int64_t n_params = 0;
    int engine = 2;                             // 0 direct, 1 layer-wise tcgen05, 2 block-fused tcgen05 (falls back per block)
    imk::FusedBlock fb_enc[5];                  // [0] = FRONT (in + enc1), [1..3] = enc2..4, [4] = bottleneck (conv3 + conv1, no pool)
    imk::FusedBlock fb_dec[4];                  // decoder block that OUTPUTS level l
    imk::FusedBlock fb_front_u8;                // FRONT for uint8 images: input block on the loader warps, chain of two (kind 3)
    imk::FusedBlock fb_head;                    // level-0 decoder + output layer + activation / decision (kind 4, K <= 3 on <= 32 channels)
    bool c8 = false;                            // fused engine: maps of <= 8 channels are one 16-byte plane per pixel (c1, p1, c8, c9 at alpha = 0.5)
    float *w_out8 = nullptr;                    // output layer [K][8] fp32 for that layout (int(16a) <= 8 only)
    std::vector<void *> owned;                  // device allocations freed at destroy
    // workspace (grown on demand), for `cap_n` images
    int64_t cap_n = 0;
    imk::Level lvl[5];
    uint8_t *dec = nullptr;                     // per-pixel decision byte of the head stage (votes / class id), cap_n images
    cudaEvent_t last_use = nullptr;             // recorded after the last consumer of the workspace; waited by the next trunk
    void *ws = nullptr;
    size_t ws_bytes = 0;
    void *stage_in = nullptr;                   // staging for *_host calls
    size_t stage_in_bytes = 0;
    float *stage_probs = nullptr;
    size_t stage_probs_bytes = 0;
};

namespace imk {

// Runs the 23 hidden layers for images [n0, n0+n) and leaves c9 (decoder level-0
// output, fp16 [n,H,W,C1p]) in net->lvl[0].a.  `images` points at image n0.
// With `head` (and unet_has_head(net)) the last kernel also runs the output layer and writes what `head` asks for
// instead of c9.
int unet_trunk(imk_unet *net, const void *images, int in_dtype, int swap_rb, int64_t n, cudaStream_t stream, const HeadOut *head = nullptr);
inline bool unet_has_head(const imk_unet *net) { return net->engine == 2 && net->fb_head.ok; }
// channels per pixel of c9 as the trunk leaves it (lvl[0].a) and the output layer's weights [K][that many] fp32
inline int unet_c9_channels(const imk_unet *net) { return (net->engine == 2 && net->c8) ? 8 : net->conv.back().cin_p; }
inline const float *unet_out_weights(const imk_unet *net) { return (net->engine == 2 && net->c8) ? net->w_out8 : net->conv.back().w_f32; }
// Call after the last kernel that reads the model's workspace has been enqueued on `stream`.
int unet_mark_used(imk_unet *net, cudaStream_t stream);
int unet_reserve(imk_unet *net, int64_t n);
int64_t max_chunk();                    // images per trunk pass (bounds the workspace; IMK_CHUNK overrides the default)
#define kMaxChunk (imk::max_chunk())

// Building blocks shared with the EvalNet forward (imk_evalnet.cu): one Conv2D (+ReLU)(+BN) packed for every engine and
// launched on the engine that fits (1 / 2: layer-wise tcgen05 when a strip plan exists, else the direct kernel).
// bn: {gamma, beta, moving_mean, moving_variance} or null.  first: the 1x1 layer that reads the image (fp32 weights).
int conv_layer_pack(ConvLayer &L, int ks, int cin, int cout, const float *hwio, const float *bias, const float *const bn[4],
                    bool first, std::vector<void *> &owned);
int conv_layer_launch(const ConvLayer &L, int tag, int engine, const __half *in, const __half *in_lo, __half *out,
                      int64_t n, int h, int w, cudaStream_t stream);
int in_conv_launch(const ConvLayer &L, const void *img, int in_dtype, int c, int swap_rb, int normalize, __half *out, int64_t px,
                   cudaStream_t stream);
int maxpool_launch(const __half *in, __half *out, int64_t n, int h, int w, int cp, cudaStream_t stream);

// imk_conv_tc.cu: tcgen05 implicit-GEMM engine.  Returns IMK_OK when it handled the layer.
bool conv_tc_supported(const ConvLayer &L);
bool conv_tc_fits(const ConvLayer &L, int h, int w);   // a strip configuration exists for this resolution
int conv_tc_pack(ConvLayer &L, const float *hwio, std::vector<void *> &owned);
int conv_tc_launch(const ConvLayer &L, const __half *in, const __half *in_lo /*upsample+add source or null*/,
                   __half *out, __half *pool_out, int64_t n, int h, int w, cudaStream_t stream);

}  // namespace imk

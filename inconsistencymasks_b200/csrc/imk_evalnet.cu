// EvalNet forward: the two-input scoring CNN of the IM++ scripts, reference evalnet.py:4-73.
//
//   get_evalnet        evalnet.py:24-47    A: image, B: mask (x/255 when normalize_B) -> Dense(1, sigmoid)      (ISIC)
//   get_evalnet_miou   evalnet.py:49-73    A: image, B: one-hot class map (not normalised) -> two Dense(K, sigmoid)
//                                          heads 'iou' and 'detection'                                  (HeLa / SUIM / Cityscapes)
//
//   input_block  x[/255] -> Conv2D 1x1 (16a) + act -> BN                                   evalnet.py:4-11
//   conv_block   Conv2D ks x ks + act -> Conv2D 1x1 + act -> BN -> MaxPooling2D(2,2)       evalnet.py:14-21
//   a = conv_block(input_block(A)); b = conv_block(input_block(B)); c = concat(a, b);      evalnet.py:28-38
//   conv_block x 5 (16a, 32a, 64a, 128a, 256a) -> GlobalAvgPool2D -> Dense                 evalnet.py:40-45
//
// Built from the U-Net engines: the image-reading 1x1 layers run the fp32 first-layer kernel, every other convolution
// is a ConvLayer on the layer-wise tcgen05 engine (imk_conv_tc.cu) or the direct kernel where no strip plan fits (the
// 8x8 / 4x4 maps at the end), fp16 NHWC activations with fp32 accumulation as in the U-Net.  For the one-hot input the
// first layer is a table look-up: Conv2D 1x1 of a one-hot vector selects one row of its kernel.
// int(16 * alpha) must be a multiple of 16 (config.ini: ALPHA_EVALNET = 1 or 2) so that the concatenation of the two
// branches is dense in the 16-channel-padded layout.
#include <algorithm>
#include "imk_unet.cuh"

struct imk_evalnet {
    imk_evalnet_desc desc{};
    imk::ConvLayer in_a, in_b;                  // input blocks (fp32 1x1 on the raw input)
    imk::ConvLayer a3, a1, b3, b1;              // branch conv_blocks
    imk::ConvLayer c3[5], c1[5];                // trunk conv_blocks
    float *w_onehot = nullptr;                  // B as a class map: kernel rows of in_b, [K][cout_p] fp32 (bias added in the kernel)
    float *dense_w[2] = {nullptr, nullptr};     // [out][C] fp32
    float *dense_b[2] = {nullptr, nullptr};
    int n_heads = 1, n_out = 1, width[5] = {0, 0, 0, 0, 0};
    int64_t n_params = 0;
    std::vector<void *> owned;
    void *ws = nullptr;
    size_t ws_bytes = 0;
};

namespace imk {

// input block for a class-id map standing for its one-hot encoding (functions.py:6004-6006): one thread = pixel x 8 channels
__global__ void __launch_bounds__(256)
in_onehot_kernel(const uint8_t *__restrict__ cls, int K, const float *__restrict__ w /*[K][cout_p]*/, const float *__restrict__ bias,
                 const float *__restrict__ bn_scale, const float *__restrict__ bn_shift, int cout, int cout_p,
                 __half *__restrict__ out, int64_t total_px) {
    const int chunks = cout_p / 8;
    const int64_t total = total_px * chunks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t px = i / chunks;
        const int co0 = (int)(i % chunks) * 8;
        const int c = cls[px];
        __align__(16) __half o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int co = co0 + j;
            float v = 0.f;
            if (co < cout) {
                v = bias[co];
                if (c < K) v = __fadd_rn(v, w[(int64_t)c * cout_p + co]);      // 1 * w + b; classes >= K: an all-zero one-hot row
                v = fmaxf(v, 0.f);
                v = __fmaf_rn(v, bn_scale[co], bn_shift[co]);
            }
            o[j] = __float2half_rn(v);
        }
        *reinterpret_cast<uint4 *>(out + px * cout_p + co0) = *reinterpret_cast<const uint4 *>(o);
    }
}

// channel concatenation of two fp16 NHWC maps (tf.keras.layers.concatenate, evalnet.py:38): 128-bit vectors
__global__ void __launch_bounds__(256)
concat_kernel(const __half *__restrict__ a, int ca, const __half *__restrict__ b, int cb, __half *__restrict__ out, int64_t px) {
    const int va = ca / 8, vb = cb / 8, vt = va + vb;
    const int64_t total = px * vt;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / vt;
        const int v = (int)(i - p * vt);
        const uint4 x = v < va ? *reinterpret_cast<const uint4 *>(a + p * ca + v * 8) : *reinterpret_cast<const uint4 *>(b + p * cb + (v - va) * 8);
        *reinterpret_cast<uint4 *>(out + p * (ca + cb) + v * 8) = x;
    }
}

// GlobalAvgPool2D + Dense + sigmoid (evalnet.py:43-45 / 69-71): one CTA per image; fp32 sums in pixel order per channel
__global__ void __launch_bounds__(256)
gap_dense_kernel(const __half *__restrict__ x, int hw, int cp, int C, const float *__restrict__ w0, const float *__restrict__ b0,
                 const float *__restrict__ w1, const float *__restrict__ b1, int n_out, float *__restrict__ out0, float *__restrict__ out1) {
    extern __shared__ float mean[];
    const int64_t n = blockIdx.x;
    const __half *xi = x + n * (int64_t)hw * cp;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < hw; ++p) s += __half2float(xi[(int64_t)p * cp + c]);
        mean[c] = s / (float)hw;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 2 * n_out; o += blockDim.x) {
        const int head = o / n_out, k = o - head * n_out;
        const float *w = head ? w1 : w0, *b = head ? b1 : b0;
        if (!w) continue;
        float acc = b[k];
        for (int c = 0; c < C; ++c) acc = __fmaf_rn(mean[c], w[(int64_t)k * C + c], acc);
        (head ? out1 : out0)[n * n_out + k] = sigmoid_f32(acc);
    }
}

static int grid_for(int64_t items) {
    int64_t b = (items + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename T>
static int upload_vec(std::vector<void *> &owned, const std::vector<T> &host, T **dev) {
    void *p = nullptr;
    if (cudaMalloc(&p, host.size() * sizeof(T) + 16) != cudaSuccess) { cudaGetLastError(); set_error("evalnet: cudaMalloc(%zu) failed", host.size() * sizeof(T)); return IMK_ENOMEM; }
    owned.push_back(p);
    IMK_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dev = reinterpret_cast<T *>(p);
    return IMK_OK;
}

}  // namespace imk

using namespace imk;

// weights, in the order evalnet.py CREATES the layers (branch A, branch B, trunk, heads) -- NOT model.get_weights()
// order, which for a two-branch functional model interleaves the branches by graph depth:
//   in_a conv (kernel, bias), in_a BN (gamma, beta, mean, var), a conv3, a conv1, a BN, in_b conv, in_b BN, b conv3, b conv1,
//   b BN, then per trunk block conv3, conv1, BN, then Dense 'iou' (kernel [C][out], bias) and, with two heads, Dense 'detection'.
// inconsistencymasks_b200.evalnet.weights_from_keras(model) collects them by layer from a Keras model.
extern "C" int imk_evalnet_create(const imk_evalnet_desc *desc, const float *const *weights_host, const int64_t *weight_sizes,
                                  int n_weights, imk_evalnet_t **out) {
    IMK_REQUIRE(desc && weights_host && weight_sizes && out, "imk_evalnet_create: NULL argument");
    const imk_evalnet_desc &d = *desc;
    IMK_REQUIRE(d.height > 0 && d.width > 0 && d.height % 2 == 0 && d.width % 2 == 0, "imk_evalnet_create: bad size %dx%d", d.height, d.width);
    IMK_REQUIRE((d.height >> 6) >= 1 && (d.width >> 6) >= 1, "imk_evalnet_create: six 2x2 poolings need at least 64x64 inputs");
    IMK_REQUIRE(d.a_channels >= 1 && d.a_channels <= 4 && d.b_channels >= 1 && d.b_channels <= 255, "imk_evalnet_create: channels");
    IMK_REQUIRE(d.ks == 3, "imk_evalnet_create: ksi = 3 only");
    IMK_REQUIRE(d.n_heads == 1 || d.n_heads == 2, "imk_evalnet_create: n_heads is 1 (get_evalnet) or 2 (get_evalnet_miou)");
    IMK_REQUIRE(d.b_onehot == 0 || d.b_channels >= 1, "imk_evalnet_create: one-hot B needs its class count");
    IMK_REQUIRE(d.b_onehot || d.b_channels <= 4, "imk_evalnet_create: a dense B input has at most 4 channels");
    if (!imk_device_available()) { set_error("imk_evalnet_create: no CUDA device (there is no CPU fallback)"); return IMK_ECUDA; }
    const int f0 = (int)(16 * (double)d.alpha);
    IMK_REQUIRE(f0 >= 16 && f0 % 16 == 0, "imk_evalnet_create: int(16 * alpha) must be a multiple of 16 (config.ini: ALPHA_EVALNET = 1 or 2), got %d", f0);
    const int n_out = d.n_heads == 1 ? 1 : d.b_channels;
    const int want = 2 * 14 + 5 * 8 + 2 * d.n_heads;            // per branch: in conv 2 + BN 4, conv3 2, conv1 2 + BN 4; per trunk block 2 + 2 + 4
    IMK_REQUIRE(n_weights == want, "imk_evalnet_create: expected %d weight arrays (Keras get_weights order), got %d", want, n_weights);

    imk_evalnet *net = new imk_evalnet();
    net->desc = d; net->n_heads = d.n_heads; net->n_out = n_out;
    const int base[5] = {16, 32, 64, 128, 256};
    for (int i = 0; i < 5; ++i) net->width[i] = (int)(base[i] * (double)d.alpha);
    int wi = 0, rc = IMK_OK;
    auto fail = [&](int code) { imk_evalnet_destroy(net); return code; };
    auto take_conv = [&](ConvLayer &L, int ks, int cin, int cout, bool bn, bool first) -> int {
        const int64_t wsz = (int64_t)ks * ks * cin * cout;
        if (weight_sizes[wi] != wsz || weight_sizes[wi + 1] != cout) {
            set_error("imk_evalnet_create: weight %d: expected kernel %dx%dx%dx%d (+bias %d), got sizes %lld, %lld", wi, ks, ks, cin, cout, cout,
                      (long long)weight_sizes[wi], (long long)weight_sizes[wi + 1]);
            return IMK_EINVAL;
        }
        const float *k = weights_host[wi], *b = weights_host[wi + 1];
        wi += 2;
        net->n_params += wsz + cout;
        const float *bnp[4] = {nullptr, nullptr, nullptr, nullptr};
        if (bn) {
            for (int j = 0; j < 4; ++j) {
                if (weight_sizes[wi + j] != cout) { set_error("imk_evalnet_create: weight %d: BatchNormalization vector of %d expected", wi + j, cout); return IMK_EINVAL; }
                bnp[j] = weights_host[wi + j];
            }
            wi += 4;
            net->n_params += 4 * (int64_t)cout;
        }
        return conv_layer_pack(L, ks, cin, cout, k, b, bn ? bnp : nullptr, first, net->owned);
    };
    // NOTE on order: a conv1's BN follows it, a conv3 has none -> take_conv(conv3, no BN) then take_conv(conv1, BN)
    if ((rc = take_conv(net->in_a, 1, d.a_channels, f0, true, true))) return fail(rc);
    if ((rc = take_conv(net->a3, 3, f0, f0, false, false))) return fail(rc);
    if ((rc = take_conv(net->a1, 1, f0, f0, true, false))) return fail(rc);
    {
        const int wb = wi;
        if ((rc = take_conv(net->in_b, 1, d.b_channels, f0, true, true))) return fail(rc);
        if (d.b_onehot) {                                        // kernel rows [K][cout_p] for the look-up form
            std::vector<float> rows((size_t)d.b_channels * net->in_b.cout_p, 0.f);
            for (int c = 0; c < d.b_channels; ++c)
                for (int co = 0; co < f0; ++co) rows[(size_t)c * net->in_b.cout_p + co] = weights_host[wb][(size_t)c * f0 + co];
            if ((rc = upload_vec(net->owned, rows, &net->w_onehot))) return fail(rc);
        }
    }
    if ((rc = take_conv(net->b3, 3, f0, f0, false, false))) return fail(rc);
    if ((rc = take_conv(net->b1, 1, f0, f0, true, false))) return fail(rc);
    int cin = 2 * f0;
    for (int i = 0; i < 5; ++i) {
        if ((rc = take_conv(net->c3[i], 3, cin, net->width[i], false, false))) return fail(rc);
        if ((rc = take_conv(net->c1[i], 1, net->width[i], net->width[i], true, false))) return fail(rc);
        cin = net->width[i];
    }
    for (int hd = 0; hd < d.n_heads; ++hd) {
        const int C = net->width[4];
        if (weight_sizes[wi] != (int64_t)C * n_out || weight_sizes[wi + 1] != n_out) {
            set_error("imk_evalnet_create: weight %d: Dense kernel %dx%d (+bias %d) expected", wi, C, n_out, n_out);
            return fail(IMK_EINVAL);
        }
        std::vector<float> wt((size_t)n_out * C), bb(weights_host[wi + 1], weights_host[wi + 1] + n_out);
        for (int c = 0; c < C; ++c)
            for (int k = 0; k < n_out; ++k) wt[(size_t)k * C + c] = weights_host[wi][(size_t)c * n_out + k];     // Keras [in][out] -> [out][in]
        if ((rc = upload_vec(net->owned, wt, &net->dense_w[hd]))) return fail(rc);
        if ((rc = upload_vec(net->owned, bb, &net->dense_b[hd]))) return fail(rc);
        net->n_params += (int64_t)C * n_out + n_out;
        wi += 2;
    }
    *out = net;
    return IMK_OK;
}

extern "C" void imk_evalnet_destroy(imk_evalnet_t *net) {
    if (!net) return;
    for (void *p : net->owned) cudaFree(p);
    if (net->ws) cudaFree(net->ws);
    delete net;
}

extern "C" int imk_evalnet_param_count(const imk_evalnet_t *net, int64_t *count) {
    IMK_REQUIRE(net && count, "imk_evalnet_param_count: NULL argument");
    *count = net->n_params;
    return IMK_OK;
}

// a_dev: uint8 [N,H,W,a_channels]; b_dev: uint8 [N,H,W,b_channels] (dense input) or uint8 class map [N,H,W] (b_onehot).
// out0 / out1: float32 [N][n_out] (out1: the 'detection' head of get_evalnet_miou, else NULL).
extern "C" int imk_evalnet_forward(imk_evalnet_t *net, const uint8_t *a_dev, const uint8_t *b_dev, int64_t N, int swap_rb_a,
                                   float *out0_dev, float *out1_dev, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IMK_REQUIRE(net && a_dev && b_dev && out0_dev && N >= 0, "imk_evalnet_forward: NULL argument");
    IMK_REQUIRE(net->n_heads == 1 || out1_dev, "imk_evalnet_forward: the two-head model needs out1");
    if (N == 0) return IMK_OK;
    const imk_evalnet_desc &d = net->desc;
    const int H = d.height, W = d.width, f0p = net->in_a.cout_p;
    const int64_t chunk_max = 64;
    // workspace: three maps of the largest size (level 0, f0p channels) for a chunk
    const size_t map0 = (size_t)chunk_max * H * W * f0p * sizeof(__half);
    if (net->ws_bytes < 3 * map0) {
        if (net->ws) { IMK_CUDA(cudaStreamSynchronize(stream)); cudaFree(net->ws); net->ws = nullptr; net->ws_bytes = 0; }
        if (cudaMalloc(&net->ws, 3 * map0) != cudaSuccess) { cudaGetLastError(); set_error("imk_evalnet_forward: cudaMalloc(%zu) failed", 3 * map0); return IMK_ENOMEM; }
        net->ws_bytes = 3 * map0;
    }
    __half *X = reinterpret_cast<__half *>(net->ws), *Y = reinterpret_cast<__half *>((char *)net->ws + map0), *Z = reinterpret_cast<__half *>((char *)net->ws + 2 * map0);
    int rc;
    for (int64_t n0 = 0; n0 < N; n0 += chunk_max) {
        const int64_t n = std::min<int64_t>(chunk_max, N - n0);
        const int64_t px = n * H * W;
        const int h2 = H / 2, w2 = W / 2;
        // branch A: input block -> X, conv3 -> Y, conv1 + BN -> X, pool -> first half of the concat staging (Z holds A|B pooled maps)
        __half *pa = Z, *pb = Z + (size_t)n * h2 * w2 * f0p;
        if ((rc = in_conv_launch(net->in_a, a_dev + n0 * H * W * d.a_channels, IMK_IN_U8, d.a_channels, swap_rb_a, d.normalize_a, X, px, stream))) return rc;
        if ((rc = conv_layer_launch(net->a3, 101, 1, X, nullptr, Y, n, H, W, stream))) return rc;
        if ((rc = conv_layer_launch(net->a1, 102, 1, Y, nullptr, X, n, H, W, stream))) return rc;
        if ((rc = maxpool_launch(X, pa, n, H, W, f0p, stream))) return rc;
        // branch B
        if (d.b_onehot) {
            IMK_PROFILE("in_onehot", 0, stream);
            in_onehot_kernel<<<grid_for(px * (f0p / 8)), 256, 0, stream>>>(b_dev + n0 * H * W, d.b_channels, net->w_onehot, net->in_b.bias, net->in_b.bn_scale,
                                                                           net->in_b.bn_shift, net->in_b.cout, f0p, X, px);
            IMK_LAUNCHED();
        } else if ((rc = in_conv_launch(net->in_b, b_dev + n0 * H * W * d.b_channels, IMK_IN_U8, d.b_channels, 0, d.normalize_b, X, px, stream))) return rc;
        if ((rc = conv_layer_launch(net->b3, 103, 1, X, nullptr, Y, n, H, W, stream))) return rc;
        if ((rc = conv_layer_launch(net->b1, 104, 1, Y, nullptr, X, n, H, W, stream))) return rc;
        if ((rc = maxpool_launch(X, pb, n, H, W, f0p, stream))) return rc;
        {
            IMK_PROFILE("concat", -1, stream);
            concat_kernel<<<grid_for((int64_t)n * h2 * w2 * (2 * f0p / 8)), 256, 0, stream>>>(pa, f0p, pb, f0p, X, (int64_t)n * h2 * w2);
            IMK_LAUNCHED();
        }
        // trunk: X holds the block input
        int h = h2, w = w2;
        __half *cur = X, *t1 = Y, *t2 = Z;
        for (int i = 0; i < 5; ++i) {
            if ((rc = conv_layer_launch(net->c3[i], 110 + 2 * i, 1, cur, nullptr, t1, n, h, w, stream))) return rc;
            if ((rc = conv_layer_launch(net->c1[i], 111 + 2 * i, 1, t1, nullptr, t2, n, h, w, stream))) return rc;
            if ((rc = maxpool_launch(t2, cur, n, h, w, net->c1[i].cout_p, stream))) return rc;
            h /= 2; w /= 2;
        }
        {
            const int C = net->width[4], cp = net->c1[4].cout_p;
            IMK_PROFILE("gap_dense", -1, stream);
            gap_dense_kernel<<<(unsigned)n, 256, (size_t)C * sizeof(float), stream>>>(cur, h * w, cp, C, net->dense_w[0], net->dense_b[0], net->dense_w[1],
                                                                                   net->dense_b[1], net->n_out, out0_dev + n0 * net->n_out,
                                                                                   out1_dev ? out1_dev + n0 * net->n_out : nullptr);
            IMK_LAUNCHED();
        }
    }
    return IMK_OK;
}

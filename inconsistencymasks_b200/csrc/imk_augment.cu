// IM+ / IM++ augmentation step on the device: what the reference's
// augment_image_and_mask(s) (functions.py:2725-2828) does to ONE image and its masks between the
// pseudo-label step and the augmented training set, for a whole batch that is already in HBM.
//
//   cv2.flip(.., 0) / cv2.flip(.., 1) / cv2.rotate(.., 90cw | 180 | 90ccw)     functions.py:2795-2817   (image + masks)
//   cv2.convertScaleAbs(image, alpha, beta)                                      functions.py:2822-2823   (image)
//   cv2.GaussianBlur(image, (k, k), 0), k in {3, 5, 7}                           functions.py:1496-1502   (image)
//   clip(image + randint(-max_noise, max_noise), 0, 255)                         functions.py:1463-1477   (image)
//
// RNG contract: WHICH operations run and with which parameters is decided on the host (the reference draws them from
// `random` / `np.random`, unseeded); the pixels are computed here.  Everything except the noise is deterministic and
// bit-exact against OpenCV 4.x for the same parameters:
//   * convertScaleAbs on 8-bit data is  saturate_cast<uchar>(|fma(float(x), float(alpha), float(beta))|)  with
//     round-half-to-even (checked here against cv2 over all 256 inputs x 20000 random (alpha, beta): the fused
//     multiply-add matches everywhere, a separate multiply + add does not),
//   * GaussianBlur(sigma = 0, ksize <= 7) on 8-bit data is OpenCV's fixed-point path with the small_gaussian_tab
//     kernels (64,128,64) / (16,64,96,64,16) / (8,28,56,72,56,28,8) over 256 per axis, BORDER_REFLECT_101, one rounding
//     at the end: (sum + 2^15) >> 16,
//   * the noise is a counter-based generator keyed by (seed, image, element): same seed -> same pixels, any launch shape;
//     uniform integers in [-max_noise, max_noise) like np.random.randint -- distributional parity only, as SURVEY 8f-1 says.
#include "imk_common.cuh"

namespace imk {

// source coordinates of output pixel (y, x): undo the rotation, then the horizontal, then the vertical flip
// (the reference applies flip 0, flip 1, rotate in that order).  H, W: size of the INPUT image.
__device__ __forceinline__ void aug_source(const imk_aug_params &p, int H, int W, int y, int x, int &ys, int &xs) {
    int yr = y, xr = x;
    switch (p.rot) {
        case 1: yr = H - 1 - x; xr = y; break;                   // ROTATE_90_CLOCKWISE: out[y][x] = in[H-1-x][y]
        case 2: yr = H - 1 - y; xr = W - 1 - x; break;           // ROTATE_180
        case 3: yr = x; xr = W - 1 - y; break;                   // ROTATE_90_COUNTERCLOCKWISE: out[y][x] = in[x][W-1-y]
        default: break;
    }
    if (p.flip_h) xr = W - 1 - xr;
    if (p.flip_v) yr = H - 1 - yr;
    ys = yr; xs = xr;
}

__device__ __forceinline__ uint8_t aug_scale_abs(uint8_t v, float alpha, float beta) {
    const float f = fabsf(__fmaf_rn((float)v, alpha, beta));
    const int r = __float2int_rn(f);                             // cvRound: half to even
    return (uint8_t)min(max(r, 0), 255);
}

// splitmix64 finaliser: one 64-bit hash per (seed, element)
__device__ __forceinline__ uint32_t aug_hash(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

// geometry (image + masks) and brightness (image): one thread per output pixel
__global__ void __launch_bounds__(256)
aug_geo_kernel(const uint8_t *__restrict__ img, const uint8_t *__restrict__ masks, const imk_aug_params *__restrict__ params,
               int64_t N, int H, int W, int c, int planes, uint8_t *__restrict__ img_out, uint8_t *__restrict__ masks_out) {
    const int64_t HW = (int64_t)H * W, total = N * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / HW;
        const int r = (int)(i - n * HW);
        const imk_aug_params p = params[n];
        const int Wo = (p.rot & 1) ? H : W;                       // square when rotated by 90 degrees (checked on the host)
        const int y = r / Wo, x = r - y * Wo;
        int ys, xs;
        aug_source(p, H, W, y, x, ys, xs);
        const int64_t src = n * HW + (int64_t)ys * W + xs;
        if (img_out) {
            for (int ch = 0; ch < c; ++ch) {
                uint8_t v = img[src * c + ch];
                if (p.scale_on) v = aug_scale_abs(v, p.alpha, p.beta);
                img_out[i * c + ch] = v;
            }
        }
        if (masks_out)
            for (int pl = 0; pl < planes; ++pl) masks_out[(int64_t)pl * total + i] = masks[(int64_t)pl * total + src];
    }
}

// Gaussian blur (fixed point, separable, reflect-101) + noise on the geometry kernel's output.
// A CTA owns a 32 x 32 tile of one image: haloed tile -> shared memory, horizontal pass -> shared memory (16-bit),
// vertical pass + rounding + noise -> global.  Images without blur take the noise-only path of the same kernel.
constexpr int kAugT = 32, kAugR = 3;                             // tile edge, largest radius (k = 7)

__device__ __forceinline__ int reflect101(int v, int n) {
    if (v < 0) v = -v;
    if (v >= n) v = 2 * n - 2 - v;
    return min(max(v, 0), n - 1);                                 // images narrower than the radius: clamp (cv2 needs n > r)
}

__global__ void __launch_bounds__(256)
aug_blur_noise_kernel(const uint8_t *__restrict__ src, const imk_aug_params *__restrict__ params, int H, int W, int c,
                      uint8_t *__restrict__ dst) {
    extern __shared__ uint8_t sm[];
    const int n = blockIdx.z;
    const imk_aug_params p = params[n];
    const int Ho = (p.rot & 1) ? W : H, Wo = (p.rot & 1) ? H : W;
    const int x0 = blockIdx.x * kAugT, y0 = blockIdx.y * kAugT;
    if (x0 >= Wo || y0 >= Ho) return;
    const int64_t img_off = (int64_t)n * H * W * c;
    const int k = p.blur_k, r = k >> 1;
    const int tw = kAugT + 2 * kAugR;                             // haloed tile edge
    uint8_t *tile = sm;                                           // [tw][tw][c]
    uint16_t *hp = reinterpret_cast<uint16_t *>(sm + ((tw * tw * c + 15) / 16) * 16);   // [tw][kAugT][c]
    const int tid = threadIdx.x;
    if (k > 0) {
        for (int i = tid; i < tw * tw; i += blockDim.x) {
            const int ty = i / tw, tx = i - ty * tw;
            const int yy = reflect101(y0 - kAugR + ty, Ho), xx = reflect101(x0 - kAugR + tx, Wo);
            for (int ch = 0; ch < c; ++ch) tile[i * c + ch] = src[img_off + ((int64_t)yy * Wo + xx) * c + ch];
        }
        __syncthreads();
        const int kw3[3] = {64, 128, 64}, kw5[5] = {16, 64, 96, 64, 16}, kw7[7] = {8, 28, 56, 72, 56, 28, 8};
        const int *kw = k == 3 ? kw3 : (k == 5 ? kw5 : kw7);
        for (int i = tid; i < tw * kAugT; i += blockDim.x) {      // horizontal: rows of the haloed tile, interior columns
            const int ty = i / kAugT, tx = i - ty * kAugT;
            for (int ch = 0; ch < c; ++ch) {
                int acc = 0;
                for (int j = 0; j < k; ++j) acc += kw[j] * tile[(ty * tw + tx + kAugR - r + j) * c + ch];
                hp[i * c + ch] = (uint16_t)acc;                   // <= 255 * 256
            }
        }
        __syncthreads();
        for (int i = tid; i < kAugT * kAugT; i += blockDim.x) {
            const int ty = i / kAugT, tx = i - ty * kAugT;
            const int y = y0 + ty, x = x0 + tx;
            if (y >= Ho || x >= Wo) continue;
            for (int ch = 0; ch < c; ++ch) {
                uint32_t acc = 0;
                for (int j = 0; j < k; ++j) acc += (uint32_t)kw[j] * hp[((ty + kAugR - r + j) * kAugT + tx) * c + ch];
                int v = (int)((acc + 32768u) >> 16);
                const int64_t e = ((int64_t)y * Wo + x) * c + ch;
                if (p.noise_max > 0) {
                    const uint32_t h = aug_hash(p.seed, (uint64_t)e);
                    v += (int)(((uint64_t)h * (uint32_t)(2 * p.noise_max)) >> 32) - p.noise_max;
                }
                dst[img_off + e] = (uint8_t)min(max(v, 0), 255);
            }
        }
    } else {
        for (int i = tid; i < kAugT * kAugT; i += blockDim.x) {
            const int ty = i / kAugT, tx = i - ty * kAugT;
            const int y = y0 + ty, x = x0 + tx;
            if (y >= Ho || x >= Wo) continue;
            for (int ch = 0; ch < c; ++ch) {
                const int64_t e = ((int64_t)y * Wo + x) * c + ch;
                int v = src[img_off + e];
                if (p.noise_max > 0) {
                    const uint32_t h = aug_hash(p.seed, (uint64_t)e);
                    v += (int)(((uint64_t)h * (uint32_t)(2 * p.noise_max)) >> 32) - p.noise_max;
                }
                dst[img_off + e] = (uint8_t)min(max(v, 0), 255);
            }
        }
    }
}

}  // namespace imk

using namespace imk;

extern "C" int imk_augment_u8(const uint8_t *img_dev, const uint8_t *masks_dev, int64_t N, int H, int W, int c, int mask_planes,
                              const imk_aug_params *params_host, uint8_t *img_out_dev, uint8_t *masks_out_dev,
                              uint8_t *scratch_dev, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IMK_REQUIRE(params_host && N >= 0 && H > 0 && W > 0 && c >= 1 && c <= 4 && mask_planes >= 0, "imk_augment_u8: bad arguments");
    IMK_REQUIRE((img_dev != nullptr) == (img_out_dev != nullptr) && (img_dev || masks_dev), "imk_augment_u8: image in/out must come together");
    IMK_REQUIRE(mask_planes == 0 || (masks_dev && masks_out_dev), "imk_augment_u8: NULL masks");
    if (N == 0) return IMK_OK;
    if (!imk_device_available()) { set_error("imk_augment_u8: no CUDA device (there is no CPU fallback)"); return IMK_ECUDA; }
    bool second = false;
    for (int64_t n = 0; n < N; ++n) {
        const imk_aug_params &p = params_host[n];
        IMK_REQUIRE(p.rot >= 0 && p.rot <= 3, "imk_augment_u8: image %lld: rot=%d outside 0..3", (long long)n, p.rot);
        IMK_REQUIRE(!(p.rot & 1) || H == W, "imk_augment_u8: a 90-degree rotation needs square images (%dx%d)", H, W);
        IMK_REQUIRE(p.blur_k == 0 || p.blur_k == 3 || p.blur_k == 5 || p.blur_k == 7, "imk_augment_u8: image %lld: blur_k=%d (0, 3, 5, 7)", (long long)n, p.blur_k);
        IMK_REQUIRE(p.noise_max >= 0 && p.noise_max <= 255, "imk_augment_u8: image %lld: noise_max=%d outside 0..255", (long long)n, p.noise_max);
        second = second || p.blur_k > 0 || p.noise_max > 0;
    }
    IMK_REQUIRE(!second || !img_dev || scratch_dev, "imk_augment_u8: blur / noise need a scratch buffer of N*H*W*c bytes");
    static thread_local imk_aug_params *d_params = nullptr;
    static thread_local int64_t d_cap = 0;
    if (N > d_cap) {
        if (d_params) cudaFree(d_params);
        d_params = nullptr; d_cap = 0;
        if (cudaMalloc(&d_params, sizeof(imk_aug_params) * (size_t)N) != cudaSuccess) { cudaGetLastError(); set_error("imk_augment_u8: cudaMalloc failed"); return IMK_ENOMEM; }
        d_cap = N;
    }
    IMK_CUDA(cudaMemcpyAsync(d_params, params_host, sizeof(imk_aug_params) * (size_t)N, cudaMemcpyHostToDevice, stream));
    const int64_t total = N * (int64_t)H * W;
    const bool two_pass = second && img_dev;
    {
        int64_t b = (total + 255) / 256;
        const int grid = (int)std::min<int64_t>(b, (int64_t)num_sms() * 16);
        IMK_PROFILE("aug_geo", -1, stream);
        aug_geo_kernel<<<grid, 256, 0, stream>>>(img_dev, masks_dev, d_params, N, H, W, c, mask_planes,
                                                 img_dev ? (two_pass ? scratch_dev : img_out_dev) : nullptr, mask_planes ? masks_out_dev : nullptr);
        IMK_LAUNCHED();
    }
    if (two_pass) {
        const int tw = kAugT + 2 * kAugR;
        const size_t smem = (size_t)((tw * tw * c + 15) / 16) * 16 + (size_t)tw * kAugT * c * 2;
        const int E = std::max(H, W);
        dim3 grid((E + kAugT - 1) / kAugT, (E + kAugT - 1) / kAugT, (unsigned)N);
        IMK_REQUIRE(N <= 65535, "imk_augment_u8: at most 65535 images per call");
        IMK_PROFILE("aug_blur_noise", -1, stream);
        aug_blur_noise_kernel<<<grid, 256, smem, stream>>>(scratch_dev, d_params, H, W, c, img_out_dev);
        IMK_LAUNCHED();
    }
    // the parameter upload is asynchronous from pageable memory only in name; still, keep the buffer alive per thread
    return IMK_OK;
}

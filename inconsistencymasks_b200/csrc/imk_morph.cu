// Optional morphology and stand-alone blanking of rows a7-a9.  Off in the reference's
// config.ini (ERODE_KERNEL = DILATE_KERNEL = 0) but exercised by the EvalNet data
// makers (functions.py:3599, 3627-3635), so it is provided on the device as well.
//
//   cv2.erode / cv2.dilate(mask, np.ones((k,k)), iterations=1)   functions.py:2858-2864
//   dilate_mask (3x3 per-class dilate, larger ids win)             functions.py:3075-3100
//   image[im > 0] = 0 ; label[im > 0] = 0                           functions.py:2867-2874
//
// cv2 anchors a k x k kernel at k/2 and pads with a value that never wins (erosion:
// 255, dilation: 0), so image borders do not erode.
#include "imk_common.cuh"

namespace imk {

template <bool kErode>
__global__ void __launch_bounds__(256)
morph_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int64_t N, int H, int W, int k) {
    const int64_t total = N * H * W;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int a = k / 2;                                    // anchor: window covers [-a, k-1-a]
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const uint8_t *plane = src + (i / ((int64_t)H * W)) * H * W;
        int acc = kErode ? 255 : 0;
        for (int dy = -a; dy < k - a; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            for (int dx = -a; dx < k - a; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                const int v = plane[(int64_t)yy * W + xx];
                acc = kErode ? min(acc, v) : max(acc, v);
            }
        }
        dst[i] = (uint8_t)acc;
    }
}

__global__ void __launch_bounds__(256)
blank_kernel(const uint8_t *__restrict__ im, int64_t total_px, uint8_t *__restrict__ img, int c,
             uint8_t *__restrict__ labels, int n_labels) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_px; i += stride) {
        if (im[i] == 0) continue;
        if (img) for (int ch = 0; ch < c; ++ch) img[i * c + ch] = 0;
        if (labels) for (int l = 0; l < n_labels; ++l) labels[(int64_t)l * total_px + i] = 0;
    }
}

static int grid_1d(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace imk

using namespace imk;

static int morph(const uint8_t *src, uint8_t *dst, int64_t N, int H, int W, int k, bool erode, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    IMK_REQUIRE(src && dst && src != dst, "imk_%s_u8: src/dst NULL or aliased", erode ? "erode" : "dilate");
    IMK_REQUIRE(N >= 0 && H > 0 && W > 0, "imk_%s_u8: bad shape", erode ? "erode" : "dilate");
    if (N == 0) return IMK_OK;
    const int64_t total = N * H * W;
    if (k <= 0) {   // the reference skips the call (functions.py:2858)
        IMK_CUDA(cudaMemcpyAsync(dst, src, total, cudaMemcpyDeviceToDevice, stream));
        return IMK_OK;
    }
    if (erode) morph_kernel<true><<<grid_1d(total), 256, 0, stream>>>(src, dst, N, H, W, k);
    else       morph_kernel<false><<<grid_1d(total), 256, 0, stream>>>(src, dst, N, H, W, k);
    IMK_LAUNCHED();
    return IMK_OK;
}

extern "C" int imk_erode_u8(const uint8_t *src, uint8_t *dst, int64_t N, int H, int W, int k, void *stream) {
    return morph(src, dst, N, H, W, k, true, stream);
}
extern "C" int imk_dilate_u8(const uint8_t *src, uint8_t *dst, int64_t N, int H, int W, int k, void *stream) {
    return morph(src, dst, N, H, W, k, false, stream);
}

extern "C" int imk_blank(const uint8_t *im_dev, int64_t N, int H, int W,
                         uint8_t *img_dev, int c, uint8_t *labels_dev, int n_labels, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    IMK_REQUIRE(im_dev, "imk_blank: im is NULL");
    IMK_REQUIRE(N >= 0 && H > 0 && W > 0 && c >= 0 && n_labels >= 0, "imk_blank: bad shape");
    if (N == 0 || (!img_dev && !labels_dev)) return IMK_OK;
    const int64_t total = N * H * W;
    blank_kernel<<<grid_1d(total), 256, 0, stream>>>(im_dev, total, img_dev, c, labels_dev, n_labels);
    IMK_LAUNCHED();
    return IMK_OK;
}

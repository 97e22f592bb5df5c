// Device-side reductions of the benchmark_* helpers and the bit-packed result layout.
//
//   get_IoU_binary            functions.py:1767-1787   |gt & pred| / (|gt | pred| + 1e-7)          (non-zero = set)
//   dice_score_numpy_binary   functions.py:1837-1861   masks binarised at >= 128
//   get_IoU_multi_unique      functions.py:1790-1815   per class present in gt: |gt==i & pred==i| / |gt==i | pred==i|
//   pixel_accuracy            functions.py:1819-1834   |pred == gt| / pixels
//
// The kernels return exact integer counts per image; the host forms the quotients with the reference's own
// expressions (NumPy scalar types, round(x, 4)), so the reported numbers are those of the reference bit for bit.
// One CTA row per image: 128-bit loads, warp reductions, one atomic per (warp, counter).
#include "imk_common.cuh"

namespace imk {

__device__ __forceinline__ uint32_t nz_bytes(uint32_t w) {       // 0x01 in every non-zero byte
    const uint32_t t = (w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ((t | w) >> 7) & 0x01010101u;
}
__device__ __forceinline__ uint32_t hi_bytes(uint32_t w) { return (w >> 7) & 0x01010101u; }   // byte >= 128

// out[n][5] = { |gt!=0 & pred!=0|, |gt!=0 | pred!=0|, |gt>=128 & pred>=128|, |gt>=128|, |pred>=128| }
__global__ void __launch_bounds__(256)
seg_counts_binary_kernel(const uint8_t *__restrict__ pred, const uint8_t *__restrict__ gt, int64_t hw,
                         unsigned long long *__restrict__ out) {
    const int64_t n = blockIdx.y;
    const uint8_t *p = pred + n * hw, *g = gt + n * hw;
    uint32_t c[5] = {0, 0, 0, 0, 0};
    const int64_t nv = hw / 16;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
        const uint4 a = ldg_stream(p + v * 16), b = ldg_stream(g + v * 16);
        const uint32_t pw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t pn = nz_bytes(pw[j]), gn = nz_bytes(gw[j]), ph = hi_bytes(pw[j]), gh = hi_bytes(gw[j]);
            c[0] += __popc(pn & gn); c[1] += __popc(pn | gn); c[2] += __popc(ph & gh); c[3] += __popc(gh); c[4] += __popc(ph);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)                      // tail of an image whose size is not a multiple of 16
        for (int64_t i = nv * 16; i < hw; ++i) {
            const int pv = p[i], gv = g[i];
            c[0] += (pv && gv); c[1] += (pv || gv); c[2] += (pv >= 128 && gv >= 128); c[3] += gv >= 128; c[4] += pv >= 128;
        }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        uint32_t v = c[j];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&out[n * 5 + j], (unsigned long long)v);
    }
}

// hist[n][3][256] = per class value: pixels of gt, pixels of pred, pixels where both equal it
__global__ void __launch_bounds__(256)
seg_counts_multiclass_kernel(const uint8_t *__restrict__ pred, const uint8_t *__restrict__ gt, int64_t hw,
                             unsigned long long *__restrict__ hist) {
    __shared__ uint32_t h[3 * 256];
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const int64_t n = blockIdx.y;
    const uint8_t *p = pred + n * hw, *g = gt + n * hw;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
        const int pv = p[i], gv = g[i];
        atomicAdd(&h[gv], 1u);
        atomicAdd(&h[256 + pv], 1u);
        if (pv == gv) atomicAdd(&h[512 + gv], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[n * 768 + i], (unsigned long long)h[i]);
}

// 0/255 (or 0/non-zero) byte planes -> bit planes, 8 pixels per byte, pixel i of a row of 8 in bit i (LSB first)
__global__ void __launch_bounds__(256)
pack_bits_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int64_t n_bytes_out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_bytes_out; i += (int64_t)gridDim.x * blockDim.x) {
        const uint2 v = *reinterpret_cast<const uint2 *>(src + i * 8);
        const uint32_t a = nz_bytes(v.x), b = nz_bytes(v.y);      // 0x01 per set byte (bits 0, 8, 16, 24)
        // x * (1 + 2^7 + 2^14 + 2^21) moves byte j's flag to bit 21 + j (no two terms share a bit: no carries)
        const uint32_t rl = ((a * 0x00204081u) >> 21) & 0xFu, rh = ((b * 0x00204081u) >> 21) & 0xFu;
        dst[i] = (uint8_t)(rl | (rh << 4));
    }
}

static int grid_for(int64_t items, int per_sm = 8) {
    int64_t b = (items + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * per_sm;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace imk

using namespace imk;

extern "C" int imk_seg_counts_binary(const uint8_t *pred_dev, const uint8_t *gt_dev, int64_t N, int64_t hw, int64_t *counts_dev, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IMK_REQUIRE(pred_dev && gt_dev && counts_dev && N >= 0 && hw > 0, "imk_seg_counts_binary: bad arguments");
    IMK_REQUIRE(N <= 65535, "imk_seg_counts_binary: at most 65535 images per call");
    if (N == 0) return IMK_OK;
    if (!imk_device_available()) { set_error("imk_seg_counts_binary: no CUDA device (there is no CPU fallback)"); return IMK_ECUDA; }
    IMK_CUDA(cudaMemsetAsync(counts_dev, 0, sizeof(int64_t) * 5 * (size_t)N, stream));
    // 128-bit loads need 16-byte aligned planes: hw % 16 == 0 and aligned bases (every reference shape), else a byte loop
    IMK_REQUIRE(hw % 16 == 0 && (uintptr_t)pred_dev % 16 == 0 && (uintptr_t)gt_dev % 16 == 0,
                "imk_seg_counts_binary: planes must be 16-byte aligned with H*W a multiple of 16 (got H*W = %lld)", (long long)hw);
    const int bx = (int)std::max<int64_t>(1, std::min<int64_t>((hw / 16 + 255) / 256, std::max<int64_t>(1, (int64_t)num_sms() * 8 / N)));
    dim3 grid(bx, (unsigned)N);
    IMK_PROFILE("seg_counts_binary", -1, stream);
    seg_counts_binary_kernel<<<grid, 256, 0, stream>>>(pred_dev, gt_dev, hw, reinterpret_cast<unsigned long long *>(counts_dev));
    IMK_LAUNCHED();
    return IMK_OK;
}

extern "C" int imk_seg_counts_multiclass(const uint8_t *pred_dev, const uint8_t *gt_dev, int64_t N, int64_t hw, int64_t *hist_dev, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IMK_REQUIRE(pred_dev && gt_dev && hist_dev && N >= 0 && hw > 0, "imk_seg_counts_multiclass: bad arguments");
    IMK_REQUIRE(N <= 65535, "imk_seg_counts_multiclass: at most 65535 images per call");
    if (N == 0) return IMK_OK;
    if (!imk_device_available()) { set_error("imk_seg_counts_multiclass: no CUDA device (there is no CPU fallback)"); return IMK_ECUDA; }
    IMK_CUDA(cudaMemsetAsync(hist_dev, 0, sizeof(int64_t) * 768 * (size_t)N, stream));
    const int bx = (int)std::max<int64_t>(1, std::min<int64_t>((hw + 4095) / 4096, std::max<int64_t>(1, (int64_t)num_sms() * 8 / N)));
    dim3 grid(bx, (unsigned)N);
    IMK_PROFILE("seg_counts_multiclass", -1, stream);
    seg_counts_multiclass_kernel<<<grid, 256, 0, stream>>>(pred_dev, gt_dev, hw, reinterpret_cast<unsigned long long *>(hist_dev));
    IMK_LAUNCHED();
    return IMK_OK;
}

extern "C" int imk_pack_bits(const uint8_t *planes_dev, int64_t n_bytes, uint8_t *bits_dev, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    IMK_REQUIRE(planes_dev && bits_dev && n_bytes >= 0 && n_bytes % 8 == 0, "imk_pack_bits: n_bytes must be a multiple of 8");
    IMK_REQUIRE((uintptr_t)planes_dev % 8 == 0, "imk_pack_bits: the planes must be 8-byte aligned");
    if (n_bytes == 0) return IMK_OK;
    if (!imk_device_available()) { set_error("imk_pack_bits: no CUDA device (there is no CPU fallback)"); return IMK_ECUDA; }
    IMK_PROFILE("pack_bits", -1, stream);
    pack_bits_kernel<<<grid_for(n_bytes / 8), 256, 0, stream>>>(planes_dev, bits_dev, n_bytes / 8);
    IMK_LAUNCHED();
    return IMK_OK;
}

// Standalone Inconsistency-Mask kernels: materialised fp32 probabilities in,
// pseudo-label / IM / blanked image / per-image sizes out.  HBM-bound: every byte is
// touched once, all global accesses are 128-bit and warp-contiguous, the per-image
// statistics are warp-reduced before a single RED per (warp, image).
//
// Reference semantics:
//   binary / HeLa : functions.py:3140-3202 + 3104-3120 + 2867-2874 / 2968-2974
//   multiclass    : functions.py:3206-3238 + 3123-3137 + 3054-3061
#include "imk_im.cuh"

namespace imk {

struct ProbPtrs {
    const float *p[IMK_MAX_MODELS];
};

// =============================================================================
//  binary / HeLa, vector path.  Requires H*W % 16 == 0 and 16-byte aligned bases.
//  A warp owns a chunk of 512 consecutive pixels of the flattened [N*H*W] batch.
//  Loads: per model 4K fully coalesced LDG.128 per lane (512 B contiguous per warp
//  instruction).  The per-element vote counts (one byte each) are transposed through
//  a 512*K-byte shared-memory slab so that each lane ends up with the K*16 count
//  bytes of ITS 16 consecutive pixels and can issue 128-bit stores for every output.
// =============================================================================
constexpr int kBinWarps = 8;     // 256 threads
constexpr int kChunkPx = 512;    // pixels per warp iteration

template <int K>
__global__ void __launch_bounds__(kBinWarps * 32)
im_binary_vec_kernel(ProbPtrs probs, int M, int64_t total_px, int64_t HW, float thr, int strict,
                     const uint8_t *__restrict__ img, int c, int block_in, int block_out,
                     uint8_t *__restrict__ img_out, uint8_t *__restrict__ labels, uint8_t *__restrict__ im_out,
                     int64_t *__restrict__ im_size, int64_t *__restrict__ pred_size, int64_t N) {
    __shared__ __align__(16) uint32_t slab[kBinWarps][kChunkPx * K / 4];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t n_chunks = (total_px + kChunkPx - 1) / kChunkPx;
    const int64_t warp_stride = (int64_t)gridDim.x * kBinWarps;
    const bool strict_b = strict != 0;

    for (int64_t chunk = (int64_t)blockIdx.x * kBinWarps + warp; chunk < n_chunks; chunk += warp_stride) {
        const int64_t px0 = chunk * kChunkPx;
        const int64_t rem_f = (total_px - px0) * K;           // floats left from the chunk start
        uint32_t cnt[4 * K];
#pragma unroll
        for (int j = 0; j < 4 * K; ++j) cnt[j] = 0;

        for (int m = 0; m < M; ++m) {
            const float *base = probs.p[m] + px0 * K;
            uint4 v[4 * K];
#pragma unroll
            for (int j = 0; j < 4 * K; ++j) {
                const int f = (j * 32 + lane) * 4;
                v[j] = (f < rem_f) ? ldg_stream(base + f) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int j = 0; j < 4 * K; ++j) {
                cnt[j] += decide(__uint_as_float(v[j].x), thr, strict_b)
                        | (decide(__uint_as_float(v[j].y), thr, strict_b) << 8)
                        | (decide(__uint_as_float(v[j].z), thr, strict_b) << 16)
                        | (decide(__uint_as_float(v[j].w), thr, strict_b) << 24);
            }
        }
        // transpose: element e (= px*K + head) of the chunk lives in byte e of the slab
#pragma unroll
        for (int j = 0; j < 4 * K; ++j) slab[warp][j * 32 + lane] = cnt[j];
        __syncwarp();
        uint32_t votes[4 * K];                                // K*16 bytes: [16 px][K heads]
#pragma unroll
        for (int q = 0; q < K; ++q) {
            const uint4 t = *reinterpret_cast<const uint4 *>(&slab[warp][lane * 4 * K + 4 * q]);
            votes[4 * q + 0] = t.x; votes[4 * q + 1] = t.y; votes[4 * q + 2] = t.z; votes[4 * q + 3] = t.w;
        }
        __syncwarp();

        const int64_t px = px0 + 16 * lane;
        const bool live = px < total_px;
        uint32_t lab_bits[K];
        uint32_t im_bits = 0, im_cnt = 0;
        uint32_t pred_cnt[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { lab_bits[k] = 0; pred_cnt[k] = 0; }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int e = i * K + k;
                const uint32_t s = (votes[e >> 2] >> (8 * (e & 3))) & 0xFFu;
                const uint32_t all = (s == (uint32_t)M);
                const uint32_t mixed = (s != 0u) & (s != (uint32_t)M);
                lab_bits[k] |= all << i;
                im_bits |= mixed << i;
                im_cnt += mixed;
                pred_cnt[k] += all;
            }
        }
        if (!live) { im_cnt = 0; }
        // per-image statistics (counted before blanking, functions.py:3114-3115)
        const bool uniform = (px0 / HW) == ((px0 + kChunkPx - 1) / HW) && (px0 + kChunkPx <= total_px);
        const int64_t n = live ? px / HW : -1;
        const int64_t n_u = uniform ? px0 / HW : n;
        warp_add_stat(im_size, n_u, im_cnt, uniform);
        if (pred_size) {
#pragma unroll
            for (int k = 0; k < K; ++k) warp_add_stat(pred_size + (int64_t)k * N, n_u, live ? pred_cnt[k] : 0u, uniform);
        }
        if (live) {
            auto expand = [](uint32_t bits16, int q) {       // 4 pixels -> 4 bytes of 0x00 / 0xFF
                uint32_t w = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) if ((bits16 >> (4 * q + e)) & 1u) w |= 0xFFu << (8 * e);
                return w;
            };
#pragma unroll
            for (int k = 0; k < K; ++k) {
                // HeLa: the raw position head (k == 2) is never blanked -- the reference blanks the
                // circle image drawn from it on the host instead (functions.py:2953-2974)
                const uint32_t b = (block_out && k < 2) ? (lab_bits[k] & ~im_bits) : lab_bits[k];
                stg_stream(labels + (int64_t)k * total_px + px,
                           make_uint4(expand(b, 0), expand(b, 1), expand(b, 2), expand(b, 3)));
            }
            stg_stream(im_out + px, make_uint4(expand(im_bits, 0), expand(im_bits, 1), expand(im_bits, 2), expand(im_bits, 3)));
            if (img_out) blank_image16_any(img, img_out, c, px, im_bits, block_in != 0);
        }
    }
}

// =============================================================================
//  binary / HeLa, generic path: any shape, any alignment, one pixel per thread.
// =============================================================================
__global__ void __launch_bounds__(256)
im_binary_generic_kernel(ProbPtrs probs, int M, int64_t total_px, int64_t HW, int K, float thr, int strict,
                         const uint8_t *__restrict__ img, int c, int block_in, int block_out,
                         uint8_t *__restrict__ img_out, uint8_t *__restrict__ labels, uint8_t *__restrict__ im_out,
                         int64_t *__restrict__ im_size, int64_t *__restrict__ pred_size, int64_t N) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounded = (total_px + 31) / 32 * 32;        // keep warps converged for the shuffles
    for (int64_t px = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; px < rounded; px += stride) {
        const bool live = px < total_px;
        uint32_t im_any = 0, im_cnt = 0;
        uint32_t lab[3] = {0, 0, 0};
        if (live) {
            for (int k = 0; k < K; ++k) {
                uint32_t s = 0;
                for (int m = 0; m < M; ++m) s += decide(probs.p[m][px * K + k], thr, strict != 0);
                lab[k] = (s == (uint32_t)M);
                const uint32_t mixed = (s != 0u) & (s != (uint32_t)M);
                im_any |= mixed;
                im_cnt += mixed;
            }
        }
        const int64_t n = live ? px / HW : -1;
        warp_add_stat(im_size, n, im_cnt, false);
        if (pred_size)
            for (int k = 0; k < K; ++k) warp_add_stat(pred_size + (int64_t)k * N, n, lab[k], false);
        if (live) {
            for (int k = 0; k < K; ++k)
                labels[(int64_t)k * total_px + px] = (lab[k] && !(block_out && k < 2 && im_any)) ? 255 : 0;
            im_out[px] = im_any ? 255 : 0;
            if (img_out)
                for (int ch = 0; ch < c; ++ch)
                    img_out[px * c + ch] = (block_in && im_any) ? 0 : img[px * c + ch];
        }
    }
}

// =============================================================================
//  multiclass, TMA path.  Persistent CTAs; each tile of P pixels of every model's
//  [px][K] fp32 slab is brought into shared memory with one 1-D bulk copy
//  (cp.async.bulk, SASS UBLKCP) per model, double-buffered on two mbarriers so the
//  copy of tile i+1 overlaps the argmax of tile i.  One thread per pixel scans its K
//  values from shared memory (pitch K words: conflict-free for odd K such as 9 / 35),
//  label / IM bytes are regrouped through shared memory and leave as 128-bit stores.
// =============================================================================
constexpr int kMcThreads = 256;

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kMcThreads)
im_multiclass_tma_kernel(ProbPtrs probs, int M, int64_t total_px, int64_t HW, int K, int P,
                         const uint8_t *__restrict__ img, int c, int block_in, int block_out,
                         uint8_t *__restrict__ img_out, uint8_t *__restrict__ label_out, uint8_t *__restrict__ im_out,
                         int64_t *__restrict__ im_size, unsigned long long *__restrict__ presence) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // layout: [2][M][P*K] float | label_s[P] | im_s[P] | full[2]
    float *stage_base = reinterpret_cast<float *>(smem_raw);
    const size_t stage_floats = (size_t)M * P * K;
    uint8_t *label_s = smem_raw + 2 * stage_floats * sizeof(float);
    uint8_t *im_s = label_s + P;
    uint64_t *full = reinterpret_cast<uint64_t *>(im_s + P);

    const int tid = threadIdx.x;
    const int64_t n_tiles = (total_px + P - 1) / P;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int64_t tile, int stage) {
        const int64_t p0 = tile * P;
        const int64_t pc = min((int64_t)P, total_px - p0);
        const uint32_t bytes = (uint32_t)(pc * K * sizeof(float));
        mbar_expect_tx(&full[stage], bytes * M);
        for (int m = 0; m < M; ++m)
            bulk_g2s(stage_base + stage * stage_floats + (size_t)m * P * K, probs.p[m] + p0 * K, bytes, &full[stage]);
    };

    int64_t tile = blockIdx.x;
    if (tid == 0 && tile < n_tiles) issue(tile, 0);
    for (int it = 0; tile < n_tiles; ++it, tile += gridDim.x) {
        const int stage = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        const int64_t next = tile + gridDim.x;
        if (tid == 0 && next < n_tiles) issue(next, stage ^ 1);   // stage^1 was drained before the last barrier
        mbar_wait(&full[stage], parity);

        const int64_t p0 = tile * P;
        const int pc = (int)min((int64_t)P, total_px - p0);
        const float *st = stage_base + stage * stage_floats;
        const bool uniform = (p0 / HW) == ((p0 + pc - 1) / HW);
        const int64_t n_tile = p0 / HW;
        for (int pb = 0; pb < pc; pb += kMcThreads) {           // pc, kMcThreads multiples of 16/32: warps stay converged
            const int p = pb + tid;
            const bool live = p < pc;
            uint32_t disagree = 0;
            int a0 = 0;
            if (live) {
                for (int m = 0; m < M; ++m) {
                    const float *row = st + ((size_t)m * P + p) * K;
                    float best = row[0];
                    int arg = 0;
                    for (int k = 1; k < K; ++k) argmax_step(row[k], k, best, arg);
                    if (m == 0) a0 = arg; else disagree |= (arg != a0);
                }
                label_s[p] = disagree ? 0 : (uint8_t)a0;
                im_s[p] = disagree ? 255 : 0;
            }
            const int64_t n = live ? (p0 + p) / HW : -1;
            warp_add_stat(im_size, uniform ? n_tile : n, live ? disagree : 0u, uniform);
        }
        if (presence) {
            // class-set presence per (image, model): second pass keeps the hot loop lean
            for (int m = 0; m < M; ++m) {
                for (int pb = 0; pb < pc; pb += kMcThreads) {
                    const int p = pb + tid;
                    const bool live = p < pc;
                    unsigned long long bit = 0;
                    if (live) {
                        const float *row = st + ((size_t)m * P + p) * K;
                        float best = row[0];
                        int arg = 0;
                        for (int k = 1; k < K; ++k) argmax_step(row[k], k, best, arg);
                        bit = 1ull << arg;
                    }
                    const int64_t n = live ? (p0 + p) / HW : -1;
                    const int64_t nn = uniform ? n_tile : n;
                    warp_or_stat(presence, nn < 0 ? -1 : nn * M + m, bit, uniform);
                }
            }
        }
        __syncthreads();
        // 128-bit epilogue: 16 pixels per thread
        for (int v = tid; v * 16 < pc; v += kMcThreads) {
            const uint4 lab = *reinterpret_cast<const uint4 *>(label_s + 16 * v);
            const uint4 imv = *reinterpret_cast<const uint4 *>(im_s + 16 * v);
            const int64_t px = p0 + 16 * v;
            stg_stream(label_out + px, lab);      // already 0 where the IM is set: block_out is a no-op here
            stg_stream(im_out + px, imv);
            if (img_out) {
                const uint32_t w[4] = {imv.x, imv.y, imv.z, imv.w};
                uint32_t bits = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) bits |= ((w[i >> 2] >> (8 * (i & 3))) & 1u) << i;
                blank_image16_any(img, img_out, c, px, bits, block_in != 0);
            }
        }
        __syncthreads();
    }
}

// multiclass, generic path (any shape / K / alignment): one pixel per thread.
__global__ void __launch_bounds__(256)
im_multiclass_generic_kernel(ProbPtrs probs, int M, int64_t total_px, int64_t HW, int K,
                             const uint8_t *__restrict__ img, int c, int block_in, int block_out,
                             uint8_t *__restrict__ img_out, uint8_t *__restrict__ label_out, uint8_t *__restrict__ im_out,
                             int64_t *__restrict__ im_size, unsigned long long *__restrict__ presence) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounded = (total_px + 31) / 32 * 32;
    for (int64_t px = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; px < rounded; px += stride) {
        const bool live = px < total_px;
        const int64_t n = live ? px / HW : -1;
        uint32_t disagree = 0;
        int a0 = 0;
        for (int m = 0; m < M; ++m) {
            int arg = 0;
            if (live) {
                const float *row = probs.p[m] + px * K;
                float best = row[0];
                for (int k = 1; k < K; ++k) argmax_step(row[k], k, best, arg);
                if (m == 0) a0 = arg; else disagree |= (arg != a0);
            }
            if (presence) warp_or_stat(presence, n < 0 ? -1 : n * M + m, live ? (1ull << (arg & 63)) : 0ull, false);
        }
        warp_add_stat(im_size, n, live ? disagree : 0u, false);
        if (live) {
            label_out[px] = disagree ? 0 : (uint8_t)a0;
            im_out[px] = disagree ? 255 : 0;
            if (img_out)
                for (int ch = 0; ch < c; ++ch)
                    img_out[px * c + ch] = (block_in && disagree) ? 0 : img[px * c + ch];
        }
    }
}

__global__ void lists_equal_kernel(const unsigned long long *__restrict__ presence, int M, int64_t N,
                                   uint8_t *__restrict__ lists_equal) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    bool eq = true;
    for (int m = 1; m < M; ++m) eq &= presence[n * M + m] == presence[n * M];
    lists_equal[n] = eq ? 1 : 0;
}

// =============================================================================
//  pred_masks_to_im_binary / _multiclass on integer masks (functions.py:3104-3137)
// =============================================================================
__global__ void __launch_bounds__(256)
masks_to_im_kernel(const int64_t *__restrict__ masks, int M, int64_t P, int multiclass,
                   uint8_t *__restrict__ label, uint8_t *__restrict__ im, int64_t *__restrict__ sizes) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounded = (P + 31) / 32 * 32;
    for (int64_t px = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; px < rounded; px += stride) {
        const bool live = px < P;
        uint32_t im_px = 0, pred_px = 0;
        if (live) {
            if (multiclass) {
                const int64_t first = masks[px];
                bool agree = true;
                for (int m = 1; m < M; ++m) agree &= masks[(int64_t)m * P + px] == first;
                label[px] = agree ? (uint8_t)first : 0;       // astype(np.uint8) wraps
                im[px] = agree ? 0 : 255;
                im_px = !agree;
            } else {
                int64_t s = 0;
                for (int m = 0; m < M; ++m) s += masks[(int64_t)m * P + px];
                pred_px = (s == M);
                im_px = (s != 0) && (s != M);
                label[px] = pred_px ? 255 : 0;
                im[px] = im_px ? 255 : 0;
            }
        }
        warp_add_stat(sizes, live ? 0 : -1, im_px, false);
        if (!multiclass) warp_add_stat(sizes, live ? 1 : -1, pred_px, false);
    }
}

// =============================================================================
//  host side
// =============================================================================
static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int grid_for(int64_t work_items, int per_block, int blocks_per_sm) {
    int64_t need = (work_items + per_block - 1) / per_block;
    int64_t cap = (int64_t)kNumSMs * blocks_per_sm;
    if (need < 1) need = 1;
    if (need >= cap) return (int)cap;
    return (int)need;
}

}  // namespace imk

using namespace imk;

extern "C" int imk_im_binary(const float *const *probs_dev, int M, int64_t N, int H, int W, int K,
                             float thr, int strict_gt,
                             const uint8_t *img_dev, int c, int block_in, int block_out,
                             uint8_t *img_out_dev, uint8_t *labels_dev, uint8_t *im_dev,
                             int64_t *im_size_dev, int64_t *pred_size_dev, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    IMK_REQUIRE(probs_dev && labels_dev && im_dev && im_size_dev, "imk_im_binary: NULL probs/labels/im/im_size");
    IMK_REQUIRE(M >= 1 && M <= IMK_MAX_MODELS, "imk_im_binary: M=%d outside 1..%d", M, IMK_MAX_MODELS);
    IMK_REQUIRE(K == 1 || K == 3, "imk_im_binary: K=%d (1 = ISIC head, 3 = HeLa heads)", K);
    IMK_REQUIRE(N >= 0 && H > 0 && W > 0, "imk_im_binary: bad shape N=%lld H=%d W=%d", (long long)N, H, W);
    IMK_REQUIRE(!img_out_dev || (img_dev && c >= 1 && c <= 4), "imk_im_binary: img_out needs img and 1 <= c <= 4");
    if (N == 0) return IMK_OK;
    ProbPtrs pp{};
    bool vec = ((int64_t)H * W) % 16 == 0 && aligned16(labels_dev) && aligned16(im_dev) &&
               (!img_out_dev || (aligned16(img_dev) && aligned16(img_out_dev)));
    for (int m = 0; m < M; ++m) {
        IMK_REQUIRE(probs_dev[m], "imk_im_binary: probs[%d] is NULL", m);
        pp.p[m] = probs_dev[m];
        vec = vec && aligned16(probs_dev[m]);
    }
    const int64_t HW = (int64_t)H * W, total = N * HW;
    IMK_CUDA(cudaMemsetAsync(im_size_dev, 0, sizeof(int64_t) * N, stream));
    if (pred_size_dev) IMK_CUDA(cudaMemsetAsync(pred_size_dev, 0, sizeof(int64_t) * N * K, stream));
    IMK_PROFILE(vec ? "im_binary_vec" : "im_binary_generic", -1, stream);
    if (vec) {
        const int grid = grid_for((total + kChunkPx - 1) / kChunkPx, kBinWarps, 8);
        if (K == 1)
            im_binary_vec_kernel<1><<<grid, kBinWarps * 32, 0, stream>>>(pp, M, total, HW, thr, strict_gt, img_dev, c, block_in,
                                                                          block_out, img_out_dev, labels_dev, im_dev,
                                                                          im_size_dev, pred_size_dev, N);
        else
            im_binary_vec_kernel<3><<<grid, kBinWarps * 32, 0, stream>>>(pp, M, total, HW, thr, strict_gt, img_dev, c, block_in,
                                                                          block_out, img_out_dev, labels_dev, im_dev,
                                                                          im_size_dev, pred_size_dev, N);
    } else {
        const int grid = grid_for(total, 256, 8);
        im_binary_generic_kernel<<<grid, 256, 0, stream>>>(pp, M, total, HW, K, thr, strict_gt, img_dev, c, block_in, block_out,
                                                           img_out_dev, labels_dev, im_dev, im_size_dev, pred_size_dev, N);
    }
    IMK_LAUNCHED();
    return IMK_OK;
}

namespace imk {
// scratch for the class-set presence masks of lists_equal (grown on demand, per thread)
static thread_local unsigned long long *g_presence = nullptr;
static thread_local size_t g_presence_cap = 0;
}  // namespace imk

extern "C" int imk_im_multiclass(const float *const *probs_dev, int M, int64_t N, int H, int W, int K,
                                 const uint8_t *img_dev, int c, int block_in, int block_out,
                                 uint8_t *img_out_dev, uint8_t *label_dev, uint8_t *im_dev,
                                 int64_t *im_size_dev, uint8_t *lists_equal_dev, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    IMK_REQUIRE(probs_dev && label_dev && im_dev && im_size_dev, "imk_im_multiclass: NULL probs/label/im/im_size");
    IMK_REQUIRE(M >= 1 && M <= IMK_MAX_MODELS, "imk_im_multiclass: M=%d outside 1..%d", M, IMK_MAX_MODELS);
    IMK_REQUIRE(K >= 1 && K <= IMK_MAX_CLASSES, "imk_im_multiclass: K=%d outside 1..%d", K, IMK_MAX_CLASSES);
    IMK_REQUIRE(N >= 0 && H > 0 && W > 0, "imk_im_multiclass: bad shape N=%lld H=%d W=%d", (long long)N, H, W);
    IMK_REQUIRE(!img_out_dev || (img_dev && c >= 1 && c <= 4), "imk_im_multiclass: img_out needs img and 1 <= c <= 4");
    IMK_REQUIRE(!lists_equal_dev || K <= 64, "imk_im_multiclass: lists_equal needs K <= 64 (K=%d)", K);
    if (N == 0) return IMK_OK;
    ProbPtrs pp{};
    bool vec = ((int64_t)H * W) % 16 == 0 && aligned16(label_dev) && aligned16(im_dev) &&
               (!img_out_dev || (aligned16(img_dev) && aligned16(img_out_dev)));
    for (int m = 0; m < M; ++m) {
        IMK_REQUIRE(probs_dev[m], "imk_im_multiclass: probs[%d] is NULL", m);
        pp.p[m] = probs_dev[m];
        vec = vec && aligned16(probs_dev[m]);
    }
    const int64_t HW = (int64_t)H * W, total = N * HW;
    IMK_CUDA(cudaMemsetAsync(im_size_dev, 0, sizeof(int64_t) * N, stream));
    unsigned long long *presence = nullptr;
    if (lists_equal_dev) {
        const size_t need = sizeof(unsigned long long) * (size_t)N * M;
        if (need > g_presence_cap) {
            if (g_presence) cudaFree(g_presence);
            g_presence = nullptr; g_presence_cap = 0;
            if (cudaMalloc(&g_presence, need) != cudaSuccess) { set_error("imk_im_multiclass: cudaMalloc(%zu) failed", need); return IMK_ENOMEM; }
            g_presence_cap = need;
        }
        presence = g_presence;
        IMK_CUDA(cudaMemsetAsync(presence, 0, need, stream));
    }
    // tile size: the largest multiple of 128 pixels whose M slabs fit ~48 KB per stage
    int P = 0;
    if (vec) {
        const size_t per_px = (size_t)M * K * sizeof(float);
        P = (int)((48 * 1024) / per_px) / 128 * 128;
        if (P > 1024) P = 1024;
        if (P < 128) P = ((size_t)2 * 128 * per_px + 2 * 128 + 64 <= 200 * 1024) ? 128 : 0;
    }
    {
    IMK_PROFILE((vec && P > 0) ? "im_multiclass_tma" : "im_multiclass_generic", -1, stream);
    if (vec && P > 0) {
        const size_t smem = (size_t)2 * M * P * K * sizeof(float) + 2 * (size_t)P + 2 * sizeof(uint64_t);
        IMK_CUDA(cudaFuncSetAttribute(im_multiclass_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = (int)((220 * 1024) / (smem + 1024));
        const int grid = grid_for((total + P - 1) / P, 1, per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm));
        im_multiclass_tma_kernel<<<grid, kMcThreads, smem, stream>>>(pp, M, total, HW, K, P, img_dev, c, block_in, block_out,
                                                                     img_out_dev, label_dev, im_dev, im_size_dev, presence);
    } else {
        const int grid = grid_for(total, 256, 8);
        im_multiclass_generic_kernel<<<grid, 256, 0, stream>>>(pp, M, total, HW, K, img_dev, c, block_in, block_out,
                                                               img_out_dev, label_dev, im_dev, im_size_dev, presence);
    }
    IMK_LAUNCHED();
    }
    if (lists_equal_dev) {
        lists_equal_kernel<<<(int)((N + 255) / 256), 256, 0, stream>>>(presence, M, N, lists_equal_dev);
        IMK_LAUNCHED();
    }
    return IMK_OK;
}

extern "C" int imk_masks_to_im_binary(const int64_t *masks_dev, int M, int64_t P,
                                      uint8_t *label_dev, uint8_t *im_dev, int64_t *sizes_dev, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    IMK_REQUIRE(masks_dev && label_dev && im_dev && sizes_dev, "imk_masks_to_im_binary: NULL argument");
    IMK_REQUIRE(M >= 1 && P >= 0, "imk_masks_to_im_binary: M=%d P=%lld", M, (long long)P);
    IMK_CUDA(cudaMemsetAsync(sizes_dev, 0, 2 * sizeof(int64_t), stream));
    if (P == 0) return IMK_OK;
    masks_to_im_kernel<<<grid_for(P, 256, 8), 256, 0, stream>>>(masks_dev, M, P, 0, label_dev, im_dev, sizes_dev);
    IMK_LAUNCHED();
    return IMK_OK;
}

extern "C" int imk_masks_to_im_multiclass(const int64_t *masks_dev, int M, int64_t P,
                                          uint8_t *label_dev, uint8_t *im_dev, int64_t *sizes_dev, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    IMK_REQUIRE(masks_dev && label_dev && im_dev && sizes_dev, "imk_masks_to_im_multiclass: NULL argument");
    IMK_REQUIRE(M >= 1 && P >= 0, "imk_masks_to_im_multiclass: M=%d P=%lld", M, (long long)P);
    IMK_CUDA(cudaMemsetAsync(sizes_dev, 0, 2 * sizeof(int64_t), stream));
    if (P == 0) return IMK_OK;
    masks_to_im_kernel<<<grid_for(P, 256, 8), 256, 0, stream>>>(masks_dev, M, P, 1, label_dev, im_dev, sizes_dev);
    IMK_LAUNCHED();
    return IMK_OK;
}

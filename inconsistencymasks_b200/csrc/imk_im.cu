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

// [px][3 heads] interleaved 0/1 bytes (12 words = 16 px) -> 3 head planes of 4 words each
__device__ __forceinline__ void deinterleave3(const uint32_t (&w)[12], uint32_t (&plane)[3][4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
        plane[0][q] = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);   // bytes 0, 3, 6, 9
        plane[1][q] = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);   // bytes 1, 4, 7, 10
        plane[2][q] = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);   // bytes 2, 5, 8, 11
    }
}

template <int K, bool kStrict>
__global__ void __launch_bounds__(kBinWarps * 32, 3)
im_binary_vec_kernel(ProbPtrs probs, int M, int64_t total_px, int64_t HW, float thr,
                     const uint8_t *__restrict__ img, int c, int block_in, int block_out,
                     uint8_t *__restrict__ img_out, uint8_t *__restrict__ labels, uint8_t *__restrict__ im_out,
                     int64_t *__restrict__ im_size, int64_t *__restrict__ pred_size, int64_t N) {
    __shared__ __align__(16) uint32_t slab[kBinWarps][kChunkPx * K / 4];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t n_chunks = (total_px + kChunkPx - 1) / kChunkPx;
    const int64_t warp_stride = (int64_t)gridDim.x * kBinWarps;
    const uint32_t m_rep = (uint32_t)M * 0x01010101u;

    for (int64_t chunk = (int64_t)blockIdx.x * kBinWarps + warp; chunk < n_chunks; chunk += warp_stride) {
        const int64_t px0 = chunk * kChunkPx;
        const bool full = px0 + kChunkPx <= total_px;
        const int64_t rem_f = (total_px - px0) * K;           // floats left from the chunk start
        uint32_t cnt[4 * K];
#pragma unroll
        for (int j = 0; j < 4 * K; ++j) cnt[j] = 0;

        for (int m = 0; m < M; ++m) {
            const uint4 *base = reinterpret_cast<const uint4 *>(probs.p[m] + px0 * K) + lane;
            uint4 v[4 * K];
            if (full) {
#pragma unroll
                for (int j = 0; j < 4 * K; ++j) v[j] = __ldcs(base + j * 32);       // 4K independent 128-bit loads in flight
            } else {
#pragma unroll
                for (int j = 0; j < 4 * K; ++j)
                    v[j] = ((int64_t)(j * 32 + lane) * 4 < rem_f) ? __ldcs(base + j * 32) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int j = 0; j < 4 * K; ++j) {
                cnt[j] += decide(__uint_as_float(v[j].x), thr, kStrict)
                        | (decide(__uint_as_float(v[j].y), thr, kStrict) << 8)
                        | (decide(__uint_as_float(v[j].z), thr, kStrict) << 16)
                        | (decide(__uint_as_float(v[j].w), thr, kStrict) << 24);
            }
        }
        // transpose: element e (= px*K + head) of the chunk lives in byte e of the slab
#pragma unroll
        for (int j = 0; j < 4 * K; ++j) slab[warp][j * 32 + lane] = cnt[j];
        __syncwarp();
        uint32_t all01[4 * K], mix01[4 * K];                  // per element: every model fired / models disagree
#pragma unroll
        for (int q = 0; q < K; ++q) {
            const uint4 t = *reinterpret_cast<const uint4 *>(&slab[warp][lane * 4 * K + 4 * q]);
            const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t nz = bytes_nonzero01(w[e]);            // S != 0
                const uint32_t ne = bytes_nonzero01(w[e] ^ m_rep);    // S != M
                all01[4 * q + e] = ne ^ 0x01010101u;
                mix01[4 * q + e] = nz & ne;
            }
        }
        __syncwarp();

        const int64_t px = px0 + 16 * lane;
        const bool live = px < total_px;
        uint32_t lab01[K][4], im01[4];
        if constexpr (K == 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { lab01[0][q] = all01[q]; im01[q] = mix01[q]; }
        } else {
            uint32_t mixp[3][4];
            deinterleave3(reinterpret_cast<const uint32_t (&)[12]>(all01), reinterpret_cast<uint32_t (&)[3][4]>(lab01));
            deinterleave3(reinterpret_cast<const uint32_t (&)[12]>(mix01), mixp);
#pragma unroll
            for (int q = 0; q < 4; ++q) im01[q] = mixp[0][q] | mixp[1][q] | mixp[2][q];    // combined IM = max over heads
        }
        // per-image statistics, counted before blanking (functions.py:3114-3115, 3200)
        uint32_t im_cnt = 0;
#pragma unroll
        for (int j = 0; j < 4 * K; ++j) im_cnt += __popc(mix01[j]);
        const bool small = total_px < 0x7fffffffLL;              // 32-bit image index arithmetic where possible
        const int64_t n_lo = small ? (int64_t)((uint32_t)px0 / (uint32_t)HW) : px0 / HW;
        const bool uniform = full && (px0 + kChunkPx <= (n_lo + 1) * HW);
        const int64_t n_u = uniform ? n_lo : (live ? (small ? (int64_t)((uint32_t)px / (uint32_t)HW) : px / HW) : -1);
        warp_add_stat(im_size, n_u, live ? im_cnt : 0u, uniform);
        if (pred_size) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                uint32_t pc = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) pc += __popc(lab01[k][q]);
                warp_add_stat(pred_size + (int64_t)k * N, n_u, live ? pc : 0u, uniform);
            }
        }
        if (live) {
            uint32_t imw[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) imw[q] = bytes01_to_ff(im01[q]);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                // HeLa: the raw position head (k == 2) is never blanked -- the reference blanks the
                // circle image drawn from it on the host instead (functions.py:2953-2974)
                uint32_t o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    o[q] = bytes01_to_ff(lab01[k][q]);
                    if (block_out && k < 2) o[q] &= ~imw[q];
                }
                stg_stream(labels + (int64_t)k * total_px + px, make_uint4(o[0], o[1], o[2], o[3]));
            }
            stg_stream(im_out + px, make_uint4(imw[0], imw[1], imw[2], imw[3]));
            if (img_out) blank_image16_any(img, img_out, c, px, imw, block_in != 0);
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
//  multiclass, TMA path.  Persistent CTAs of 8 consumer warps + 1 producer warp.
//  The unit of transfer is ONE model's [P = 256 px][K] fp32 slab of a tile, brought into a
//  ring of shared-memory slots with one 1-D bulk copy each (cp.async.bulk, SASS UBLKCP)
//  signalled on a per-slot `full` mbarrier; consumers release a slot through its `empty`
//  mbarrier (one arrive per warp), so several slabs (up to ~170 KB per SM) are always in
//  flight while the argmax of earlier slabs runs.  One thread per pixel scans its K values
//  from shared memory (pitch K words: conflict-free for odd K such as 9 / 35) and carries
//  (argmax of model 0, disagreement) in registers across the models.  The uint8 epilogue is
//  warp-local: a ballot gives the 32-pixel IM mask, lanes 0-1 store label / IM as 128-bit
//  vectors, lanes 0..2c-1 blank the image, 16 bytes each.
// =============================================================================
constexpr int kMcConsumers = 256;
constexpr int kMcThreads = kMcConsumers + 32;
constexpr int kMcTile = 256;              // pixels per tile (one per consumer thread)
constexpr int kMcMaxSlots = 8;

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
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

template <int C, int PPT>                 // image channels at compile time (1, 3, 4; 0 = run-time); pixels per consumer thread
__global__ void __launch_bounds__(kMcThreads, 3)
im_multiclass_tma_kernel(ProbPtrs probs, int M, int64_t total_px, int64_t HW, int K, int n_slots,
                         const uint8_t *__restrict__ img, int c, int block_in, int block_out,
                         uint8_t *__restrict__ img_out, uint8_t *__restrict__ label_out, uint8_t *__restrict__ im_out,
                         int64_t *__restrict__ im_size, unsigned long long *__restrict__ presence) {
    constexpr int TILE = kMcTile * PPT;   // pixels per tile; a warp owns 32 * PPT consecutive pixels of it
    constexpr int WPX = 32 * PPT;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // layout: [n_slots][TILE*K] float | lab_s[8 warps][WPX] | full[n_slots] | empty[n_slots]
    float *slots = reinterpret_cast<float *>(smem_raw);
    const size_t slot_floats = (size_t)TILE * K;
    uint8_t *lab_s = smem_raw + (size_t)n_slots * slot_floats * sizeof(float);
    uint64_t *full = reinterpret_cast<uint64_t *>(lab_s + 8 * WPX);
    uint64_t *empty = full + kMcMaxSlots;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = (total_px + TILE - 1) / TILE;
    if (tid == 0) {
        for (int s = 0; s < n_slots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kMcConsumers / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kMcConsumers / 32) {
        // ---------------- producer: one elected lane streams (tile, model) slabs ----------------
        if (lane == 0) {
            int slot = 0, round = 0;                             // ring position (no divisions on the hot path)
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int64_t p0 = tile * TILE;
                const int64_t pc = min((int64_t)TILE, total_px - p0);
                const uint32_t bytes = (uint32_t)(pc * K * sizeof(float));
                for (int m = 0; m < M; ++m) {
                    if (round > 0) mbar_wait(&empty[slot], (uint32_t)((round - 1) & 1));
                    mbar_expect_tx(&full[slot], bytes);
                    bulk_g2s(slots + slot * slot_floats, probs.p[m] + p0 * K, bytes, &full[slot]);
                    if (++slot == n_slots) { slot = 0; ++round; }
                }
            }
        }
        return;
    }

    // ---------------- consumers: lane l of warp w owns pixels w*WPX + 32*j + l (j < PPT) of the tile ----------------
    int slot = 0;
    uint32_t phase = 0;
    const bool small = total_px < 0x7fffffffLL;                  // 32-bit image index arithmetic (64-bit division is a subroutine)
    const int cc = C > 0 ? C : c;
    constexpr int GROUPS = WPX / 16;                             // 16-pixel groups per warp (one 128-bit label / IM vector each)
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t p0 = tile * TILE;
        const int pc = (int)min((int64_t)TILE, total_px - p0);
        const int64_t n_lo = small ? (int64_t)((uint32_t)p0 / (uint32_t)HW) : p0 / HW;
        const bool uniform = (p0 + pc <= (n_lo + 1) * HW) && pc == TILE;   // whole tile inside one image: one RED per warp
        const int64_t wpx = p0 + warp * WPX;                     // first pixel of this warp
        int tp[PPT];                                             // pixel index inside the tile
        bool live[PPT];
        int64_t n_px[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            tp[j] = warp * WPX + 32 * j + lane;
            live[j] = tp[j] < pc;
            n_px[j] = live[j] ? (uniform ? n_lo : (small ? (int64_t)((uint32_t)(p0 + tp[j]) / (uint32_t)HW) : (p0 + tp[j]) / HW)) : -1;
        }
        // the image vector this lane will blank in the epilogue: requested before the slabs are scanned, so that
        // its DRAM latency is covered by the argmax work instead of stalling the warp at the end of every tile
        const int ig = lane / cc, iv = lane - ig * cc;           // 16-pixel group, 16-byte vector inside it (lanes < GROUPS * cc)
        const bool img_lane = img_out && lane < GROUPS * cc && wpx + 16 * ig < p0 + pc;
        uint4 pix = make_uint4(0, 0, 0, 0);
        if (img_lane) pix = ldg_stream(img + (wpx + 16 * ig) * cc + 16 * iv);
        uint32_t disagree[PPT];
        int a0[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) { disagree[j] = 0; a0[j] = 0; }
        for (int m = 0; m < M; ++m) {
            mbar_wait(&full[slot], phase);
            int arg[PPT];
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
                arg[j] = 0;
                if (live[j]) {
                    arg[j] = argmax_row(slots + slot * slot_floats + (size_t)tp[j] * K, K);
                    if (m == 0) a0[j] = arg[j]; else disagree[j] |= (arg[j] != a0[j]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);            // this warp is done with the slab
            if (++slot == n_slots) { slot = 0; phase ^= 1u; }
            if (presence) {
                if (uniform) {
                    unsigned long long bits = 0;
#pragma unroll
                    for (int j = 0; j < PPT; ++j) bits |= 1ull << (arg[j] & 63);
                    warp_or_stat(presence, n_lo * M + m, bits, true);
                } else {
#pragma unroll
                    for (int j = 0; j < PPT; ++j)
                        warp_or_stat(presence, n_px[j] < 0 ? -1 : n_px[j] * M + m, live[j] ? (1ull << (arg[j] & 63)) : 0ull, false);
                }
            }
        }
        if (uniform) {
            uint32_t d = 0;
#pragma unroll
            for (int j = 0; j < PPT; ++j) d += disagree[j];
            warp_add_stat(im_size, n_lo, d, true);
        } else {
#pragma unroll
            for (int j = 0; j < PPT; ++j) warp_add_stat(im_size, n_px[j], live[j] ? disagree[j] : 0u, false);
        }

        // warp-local 128-bit epilogue over the warp's WPX consecutive pixels
        uint32_t im_mask[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            im_mask[j] = __ballot_sync(0xffffffffu, live[j] && disagree[j]);
            lab_s[warp * WPX + 32 * j + lane] = (live[j] && !disagree[j]) ? (uint8_t)a0[j] : 0;
        }
        __syncwarp();
        auto group_bits = [&](int g) -> uint32_t {               // IM bits of 16-pixel group g of the warp
            uint32_t word = im_mask[0];
#pragma unroll
            for (int j = 1; j < PPT; ++j) if ((g >> 1) == j) word = im_mask[j];
            return (word >> (16 * (g & 1))) & 0xFFFFu;
        };
        if (lane < GROUPS && wpx + 16 * lane < p0 + pc) {
            const int64_t px = wpx + 16 * lane;
            const uint32_t bits = group_bits(lane);
            uint32_t imw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) imw[j] = bytes01_to_ff(bits4_to_bytes01(bits >> (4 * j)));
            stg_stream(label_out + px, *reinterpret_cast<const uint4 *>(lab_s + warp * WPX + 16 * lane));   // 0 where the IM is set
            stg_stream(im_out + px, make_uint4(imw[0], imw[1], imw[2], imw[3]));
        }
        if (img_lane) {
            const uint32_t bits = block_in ? group_bits(ig) : 0u;
            uint32_t imw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) imw[j] = bytes01_to_ff(bits4_to_bytes01(bits >> (4 * j)));
            stg_stream(img_out + (wpx + 16 * ig) * cc + 16 * iv, blank_vec_sel<C>(pix, imw, cc, iv));
        }
        __syncwarp();
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
                arg = argmax_row(probs.p[m] + px * K, K);
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
        const int grid = grid_for((total + kChunkPx - 1) / kChunkPx, kBinWarps, 3);    // 3 resident CTAs per SM (register-bound)
#define IMK_LAUNCH_BIN(KK, SS)                                                                                         \
        im_binary_vec_kernel<KK, SS><<<grid, kBinWarps * 32, 0, stream>>>(pp, M, total, HW, thr, img_dev, c, block_in, block_out, \
                                                                          img_out_dev, labels_dev, im_dev, im_size_dev,       \
                                                                          pred_size_dev, N)
        if (K == 1) { if (strict_gt) IMK_LAUNCH_BIN(1, true); else IMK_LAUNCH_BIN(1, false); }
        else        { if (strict_gt) IMK_LAUNCH_BIN(3, true); else IMK_LAUNCH_BIN(3, false); }
#undef IMK_LAUNCH_BIN
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
    // two pixels per consumer thread for narrow rows (the per-tile hand-off -- barrier waits, ballots, statistics -- is
    // then paid once per 512 pixels); ring depth / residency: prefer 3 CTAs per SM (24 consumer warps hide the
    // shared-memory latency of the argmax scan) with at least a double buffer each; fall back to 2, then 1 CTA for wide K
    const int ppt = K <= 16 ? 2 : 1;
    const int tile_px = kMcTile * ppt;
    const size_t slab_bytes = (size_t)tile_px * K * sizeof(float);
    int n_slots = 0, per_sm = 1;
    for (int ctas = 3; ctas >= 1 && n_slots < 2; --ctas) {
        per_sm = ctas;
        n_slots = (int)(((size_t)216 * 1024 / ctas - 1024) / slab_bytes);
    }
    if (n_slots > kMcMaxSlots) n_slots = kMcMaxSlots;
    const bool tma = vec && n_slots >= 2 && (c * ppt <= 8);          // image lanes: 2 * ppt groups of c vectors
    {
    IMK_PROFILE(tma ? "im_multiclass_tma" : "im_multiclass_generic", -1, stream);
    if (tma) {
        const size_t smem = (size_t)n_slots * slab_bytes + 8 * 32 * ppt + 2 * kMcMaxSlots * sizeof(uint64_t);
        const int grid = grid_for((total + tile_px - 1) / tile_px, 1, per_sm);
#define IMK_LAUNCH_MC(CC, PP)                                                                                                      \
        do {                                                                                                                       \
            IMK_CUDA(cudaFuncSetAttribute(im_multiclass_tma_kernel<CC, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            im_multiclass_tma_kernel<CC, PP><<<grid, kMcThreads, smem, stream>>>(pp, M, total, HW, K, n_slots, img_dev, c, block_in, \
                                                                                 block_out, img_out_dev, label_dev, im_dev, im_size_dev, presence); \
        } while (0)
        if (ppt == 2) { if (c == 3) IMK_LAUNCH_MC(3, 2); else if (c == 1) IMK_LAUNCH_MC(1, 2); else if (c == 4) IMK_LAUNCH_MC(4, 2); else IMK_LAUNCH_MC(0, 2); }
        else          { if (c == 3) IMK_LAUNCH_MC(3, 1); else if (c == 1) IMK_LAUNCH_MC(1, 1); else if (c == 4) IMK_LAUNCH_MC(4, 1); else IMK_LAUNCH_MC(0, 1); }
#undef IMK_LAUNCH_MC
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

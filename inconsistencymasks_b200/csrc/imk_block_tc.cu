// Block-fused tcgen05 engine: one persistent, warp-specialised kernel runs a whole U-Net
// block (unet.py:4-43) per tile without the intermediate maps ever leaving the SM:
//
//   FRONT (level 0)  uint8 image -> [x/255 -> 1x1 conv + ReLU + BN]  -> [3x3 conv + ReLU] -> [1x1 conv + ReLU + BN] -> skip (+ pooled)
//   ENC   (level>=1) pooled map  ->                                     [3x3 conv + ReLU] -> [1x1 conv + ReLU + BN] -> skip (+ pooled)
//   DEC              up2x(lo) + skip -> [1x1 conv + ReLU + BN]       -> [3x3 conv + ReLU] -> [1x1 conv + ReLU + BN] -> map
//
// i.e. a chain of up to three GEMM stages S1 (1x1 over the haloed tile), S2 (3x3), S3 (1x1).  For grayscale uint8
// images FRONT is a chain of two: the input block is a 256-entry table of finished rows (load_kind 3).
// A tile is Th x Tw output pixels of one image.  All operands live in shared memory in the
// canonical no-swizzle K-major core-matrix layout over the FLAT padded tile index
// f = row * pitch + col (pitch = Tw + 2):  addr(kc, f) = base + (kc * Pn + f) * 16, so the
// nine taps of the 3x3 stage are nine descriptor offsets into the same buffer (no im2col) and
// the output of one stage's epilogue IS the A operand of the next stage.  Accumulators live in
// TMEM (three regions R1/R2/R3, one per stage); the weights of all stages (BN scale folded in) stay
// resident in shared memory for the life of the CTA.
//
// Roles (block_tc_kernel<16, 8>: 832 threads, 1 CTA / SM, grid = min(tiles, 148), tiles strided over CTAs):
//   warps 0-15  epilogue: tcgen05.ld (two 16-column loads per wait) -> + b' (fp32) -> fp16 -> clamp on packed halves ->
//               next stage's smem operand / output staging tile (warp w owns TMEM lanes 32*(w%4).., M blocks w/4 (mod 4))
//   warp  16    TMEM allocation + tcgen05.mma issue (one elected thread, fully unrolled K steps: a small-N
//               SS MMA retires every ~39 cycles -- the A operand read, 4 KB at 128 B/clk -- so the issue
//               sequence must not cost more than that)
//   warps 17-24 loaders.  ENC / DEC: the haloed tile is ONE TMA box per 8-channel plane ({8 ch, pitch, Th+2, 1}
//               of the NHWC map lands exactly as a plane of the flat layout, zero filled outside the image); DEC
//               fetches the half-resolution tile into registers meanwhile and adds it in place (nearest-upsample-2x
//               + add, unet.py:32-33).  FRONT: uint8 pixels through a 256-entry table (x/255 split into fp16 hi + lo
//               K slots, or finished rows for grayscale).  FRONT / ENC: the loader warps also write the 2x2
//               max-pooled tile (unet.py:18) from the staging buffer between two loads.
//   warp  25    store: the finished tile leaves the staging buffer as one cp.async.bulk per image row.
// Schedule.  A1 (S1 output / S2 input; loader output in a chain of two) is double-buffered by tile parity.  The MMA
// warp issues  S1(i+1), S3(i-1), S2(i)  per iteration; the tensor pipe executes in order, so
//   * E1(i+1) and E3(i-1) run while S2(i) executes (it reads the OTHER A1 buffer),
//   * E2(i) follows S2(i) block by block (per-block tcgen05.commit), after S3(i-1) has stopped reading the single A2,
//   * the loader works one (chain of three) or two (chain of two) tiles ahead.
// All hand-offs are mbarriers (tcgen05.commit on the MMA side, complete_tx on the TMA side); every barrier completes
// exactly once per tile (per use of a loader buffer), so the wait parity is the tile (use) parity.
#include <algorithm>
#include <type_traits>
#include <math.h>
#include <stdlib.h>
#include "imk_unet.cuh"
#include "imk_im.cuh"

namespace imk {

// Two launch shapes of the same kernel (template <EW epilogue warps, LW loader warps>, + 1 MMA warp + 1 store warp):
//   <16, 8>  832 threads, 1 CTA / SM, up to 512 TMEM columns and 227 KB of shared memory
//   < 8, 4>  448 threads, 2 CTAs / SM, up to 256 TMEM columns and 112 KB each: two independent tile pipelines share
//            the SM, so one CTA's tensor-pipe work fills the other's epilogue / load hand-off bubbles
// suspend-time hint of the mbarrier waits: a waiting warp sleeps in hardware (and is woken by the phase flip) instead of
// re-issuing try_wait -- the polling instructions of 20+ waiting warps otherwise take a quarter of the issue slots
constexpr uint32_t kBtSuspendNs = 20000;
constexpr int kBtSmemMax = 227 * 1024;
constexpr int kBtSmemMax2 = 112 * 1024;
constexpr int bt_threads(int ew, int lw) { return (ew + 1 + lw + 1) * 32; }
constexpr int kBtNumBars = 12 + 3 * kBtMaxBlocks;

namespace {

// timeline probe: role 0 = epilogue warp 0, 1 = MMA warp, 2 = loader warp 0, 3 / 4 = issue of S1 / S3 block by block,
// 5 = epilogue warp 5 (CTA 0, 16 tiles from a.dbg_skip).
// Compiled in only with -DIMK_BT_TIMELINE_BUILD (IMK_BUILD_FLAGS=-DIMK_BT_TIMELINE_BUILD python -m inconsistencymasks_b200.build):
// measured (r2l), the dormant probes alone cost the block kernels 3-9 % -- they sit on the MMA warp's issue path.
#ifdef IMK_BT_TIMELINE_BUILD
#define BT_TL(role, i, ev)                                                                         \
    do {                                                                                           \
        if (a.dbg && blockIdx.x == 0 && (i) >= a.dbg_skip && (i) < a.dbg_skip + 16 && (threadIdx.x & 31) == 0) \
            a.dbg[((role) * 16 + (int)((i) - a.dbg_skip)) * 8 + ((ev) & 7)] = clock64();                 \
    } while (0)

#define BT_TLX(role, i, ev)                                                                        \
    do {                                                                                           \
        if (a.dbg && blockIdx.x == 0 && (i) >= a.dbg_skip && (i) < a.dbg_skip + 16 && (ev) < 8)    \
            a.dbg[((role) * 16 + (int)((i) - a.dbg_skip)) * 8 + (ev)] = clock64();                 \
    } while (0)
#else
#define BT_TLX(role, i, ev) do { } while (0)
#define BT_TL(role, i, ev) do { } while (0)
#endif

// ---- PTX wrappers ----------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// arrive without release semantics: the barrier only tells the MMA warp that TMEM has been read (ordered by
// tcgen05.wait::ld + tcgen05.fence); a releasing arrive would first wait for the warp's global stores to be acknowledged
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t *bar) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity), "r"(kBtSuspendNs) : "memory");
}
#ifdef IMK_BT_ACC_BUILD
__device__ __forceinline__ uint32_t bt_clk() { uint32_t c; asm volatile("mov.u32 %0, %%clock;" : "=r"(c)); return c; }
#endif
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t e;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(e));
    return e != 0;
}
// one TMA box {8 ch, box_w, box_h, 1} at (c, x, y, n) of a 4-D NHWC map -> dst (128-byte aligned shared memory)
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c, int x, int y, int n, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c), "r"(x), "r"(y), "r"(n), "r"(smem_u32(bar)) : "memory");
}

// 8-channel maps: a pixel is ONE 16-byte vector, so a tile row is contiguous in memory.  The map is then described as
// uint64 [N][H][2W] and a box row is ONE contiguous run of 16 * pitch bytes instead of `pitch` 16-byte pieces at a stride
// (the TMA unit issues one request per inner-box row: 1408 requests per plane of a 14x86 tile otherwise).
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int x, int y, int n, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(n), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tile_coords(const BtArgs &a, long long tile, int &n, int &y0, int &x0) {
    const int per_img = a.tiles_x * a.tiles_y;
    n = (int)(tile / per_img);
    const int r = (int)(tile - (long long)n * per_img);
    const int ty = r / a.tiles_x;
    y0 = ty * a.Th;
    x0 = (r - ty * a.tiles_x) * a.Tw;
}

// epilogue of 16 accumulator columns -> 8 packed half2 words.  STAGE and CH are compile-time so that every constant
// is a constant-bank operand of the FADD / FMNMX itself (BtArgs is __grid_constant__): no loads, no registers.
template <int STAGE, int CH>
__device__ __forceinline__ void epi16(const BtArgs &a, const uint32_t (&r)[16], bool keep, uint4 &lo, uint4 &hi) {
    // v = acc + b' in fp32, ONE rounding to fp16, then the clamp on packed halves: rounding is monotonic, so
    // clamp(round(v), round(lo), round(hi)) == round(clamp(v, lo, hi)) bit for bit, at half the instructions.  The
    // upper bound exists only for BN channels with a negative scale (a.has_hi: uniform per stage).
    uint32_t o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float v0 = __uint_as_float(r[2 * q]) + a.cpar[STAGE][16 * CH + 2 * q];
        const float v1 = __uint_as_float(r[2 * q + 1]) + a.cpar[STAGE][16 * CH + 2 * q + 1];
        __half2 h = __floats2half2_rn(v0, v1);
        h = __hmax2(h, *reinterpret_cast<const __half2 *>(&a.clo[STAGE][8 * CH + q]));
        if (a.has_hi[STAGE]) h = __hmin2(h, *reinterpret_cast<const __half2 *>(&a.chi[STAGE][8 * CH + q]));
        o[q] = keep ? *reinterpret_cast<const uint32_t *>(&h) : 0u;
    }
    lo = make_uint4(o[0], o[1], o[2], o[3]);
    hi = make_uint4(o[4], o[5], o[6], o[7]);
}

// Epilogue work is done in PAIRS of 16-column units: both tcgen05.ld are in flight before the one wait, so the TMEM
// read latency is exposed once per 32 columns instead of once per 16.  A unit is (M block, chunk of 16 channels):
//   PLANES: dst is an operand buffer, chunk c lands at plane 2c / 2c+1 (stride plane_bytes); else dst is the pixel's
//   row in the output tile and chunk c lands 32 bytes further
struct EpiBlk { uint32_t taddr; bool keep, store; uint8_t *dst; };

template <int STAGE, int CH, bool PLANES>
__device__ __forceinline__ void epi_store(const BtArgs &a, const uint32_t (&rr)[16], const EpiBlk &k, size_t plane_bytes) {
    uint4 lo, hi;
    epi16<STAGE, CH>(a, rr, k.keep, lo, hi);
    if (k.store) {
        if (PLANES) {
            *reinterpret_cast<uint4 *>(k.dst + (size_t)(2 * CH) * plane_bytes) = lo;
            *reinterpret_cast<uint4 *>(k.dst + (size_t)(2 * CH + 1) * plane_bytes) = hi;
        } else {
            reinterpret_cast<uint4 *>(k.dst + 32 * CH)[0] = lo;
            reinterpret_cast<uint4 *>(k.dst + 32 * CH)[1] = hi;
        }
    }
}
template <int STAGE, int CH0, int CH1, bool PLANES>
__device__ __forceinline__ void epi_pair(const BtArgs &a, const EpiBlk &k0, const EpiBlk &k1, size_t plane_bytes) {
    uint32_t r0[16], r1[16];
    tc_ld16(k0.taddr + 16 * CH0, r0);
    tc_ld16(k1.taddr + 16 * CH1, r1);
    tc_wait_ld();
    epi_store<STAGE, CH0, PLANES>(a, r0, k0, plane_bytes);
    epi_store<STAGE, CH1, PLANES>(a, r1, k1, plane_bytes);
}
template <int STAGE, int CH, bool PLANES>
__device__ __forceinline__ void epi_single(const BtArgs &a, const EpiBlk &k, size_t plane_bytes) {
    uint32_t r0[16];
    tc_ld16(k.taddr + 16 * CH, r0);
    tc_wait_ld();
    epi_store<STAGE, CH, PLANES>(a, r0, k, plane_bytes);
}

// 8-channel stages (BtStage::n8): accumulator columns 0..7 -> one 16-byte plane entry / output row
template <int STAGE>
__device__ __forceinline__ void epi8_store(const BtArgs &a, const uint32_t (&r)[8], const EpiBlk &k) {
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float v0 = __uint_as_float(r[2 * q]) + a.cpar[STAGE][2 * q];
        const float v1 = __uint_as_float(r[2 * q + 1]) + a.cpar[STAGE][2 * q + 1];
        __half2 h = __floats2half2_rn(v0, v1);
        h = __hmax2(h, *reinterpret_cast<const __half2 *>(&a.clo[STAGE][q]));
        if (a.has_hi[STAGE]) h = __hmin2(h, *reinterpret_cast<const __half2 *>(&a.chi[STAGE][q]));
        o[q] = k.keep ? *reinterpret_cast<const uint32_t *>(&h) : 0u;
    }
    if (k.store) *reinterpret_cast<uint4 *>(k.dst) = make_uint4(o[0], o[1], o[2], o[3]);
}
template <int STAGE>
__device__ __forceinline__ void epi8_pair(const BtArgs &a, const EpiBlk &k0, const EpiBlk &k1) {
    uint32_t r0[8], r1[8];
    tc_ld8(k0.taddr, r0);
    tc_ld8(k1.taddr, r1);
    tc_wait_ld();
    epi8_store<STAGE>(a, r0, k0);
    epi8_store<STAGE>(a, r1, k1);
}
template <int STAGE>
__device__ __forceinline__ void epi8_single(const BtArgs &a, const EpiBlk &k) {
    uint32_t r0[8];
    tc_ld8(k.taddr, r0);
    tc_wait_ld();
    epi8_store<STAGE>(a, r0, k);
}

// One epilogue stage of a warp: its M blocks are g, g + G, g + 2G, ... (G groups of four warps); `blk(b)` describes block b.
// 16-channel stages pair two BLOCKS per wait, wider stages pair two CHUNKS of one block.
// `cgm` = commit-group size - 1 (power of two - 1): the MMA warp commits once per group of blocks (a tcgen05.commit costs
// tensor-pipe time), so the barrier that covers block b is the one of the group's last block
__device__ __forceinline__ int commit_idx(int b, int cgm, int nb) { return min(b | cgm, nb - 1); }

template <int STAGE, int NCH, bool PLANES, int G, typename F>
__device__ __forceinline__ void epi_stage(const BtArgs &a, int nb, int g, uint64_t *acc_full, uint32_t parity, size_t plane_bytes, int cgm, F &&blk) {
    if constexpr (NCH == 0) {                                        // 8 channels: NCH = 0
        for (int b = g; b < nb; b += 2 * G) {
            const bool two = b + G < nb;                             // warp-uniform
            mbar_wait(&acc_full[commit_idx(two ? b + G : b, cgm, nb)], parity);
            __syncwarp();
            tc_fence_after();
            const EpiBlk k0 = blk(b);
            if (two) epi8_pair<STAGE>(a, k0, blk(b + G));
            else epi8_single<STAGE>(a, k0);
        }
    } else if constexpr (NCH == 1) {
        for (int b = g; b < nb; b += 2 * G) {
            const bool two = b + G < nb;                             // warp-uniform
            mbar_wait(&acc_full[commit_idx(two ? b + G : b, cgm, nb)], parity);   // blocks complete in order
            __syncwarp();
            tc_fence_after();
            const EpiBlk k0 = blk(b);
            if (two) epi_pair<STAGE, 0, 0, PLANES>(a, k0, blk(b + G), plane_bytes);
            else epi_single<STAGE, 0, PLANES>(a, k0, plane_bytes);
        }
    } else {
        for (int b = g; b < nb; b += G) {
            mbar_wait(&acc_full[commit_idx(b, cgm, nb)], parity);
            __syncwarp();
            tc_fence_after();
            const EpiBlk k = blk(b);
            epi_pair<STAGE, 0, 1, PLANES>(a, k, k, plane_bytes);
            if constexpr (NCH == 3) epi_single<STAGE, 2, PLANES>(a, k, plane_bytes);
            if constexpr (NCH == 4) epi_pair<STAGE, 2, 3, PLANES>(a, k, k, plane_bytes);
        }
    }
}

template <int NS, typename F>
__device__ __forceinline__ void with_nch(const BtStage &st, F &&f) {
    if constexpr (NS >= 0) { f(std::integral_constant<int, NS>{}); return; }
    switch (st.n8 ? 0 : st.n >> 4) {
        case 0: f(std::integral_constant<int, 0>{}); break;
        case 1: f(std::integral_constant<int, 1>{}); break;
        case 2: f(std::integral_constant<int, 2>{}); break;
        case 3: f(std::integral_constant<int, 3>{}); break;
        default: f(std::integral_constant<int, 4>{}); break;
    }
}

// fp16x8 += fp16x8 (unet.py:33 `add`; the reference runs mixed_float16, so the add is an fp16 add)
__device__ __forceinline__ uint4 add_h8(const uint4 &v, const uint4 &u) {
    const __half2 *pa = reinterpret_cast<const __half2 *>(&v);
    const __half2 *pb = reinterpret_cast<const __half2 *>(&u);
    uint4 r4;
    __half2 *ro = reinterpret_cast<__half2 *>(&r4);
#pragma unroll
    for (int e = 0; e < 4; ++e) ro[e] = __hadd2(pa[e], pb[e]);
    return r4;
}

// fp16x8 max (MaxPooling2D on fp16 maps: exact)
__device__ __forceinline__ uint4 max_h8(const uint4 &v, const uint4 &u) {
    const __half2 *pa = reinterpret_cast<const __half2 *>(&v);
    const __half2 *pb = reinterpret_cast<const __half2 *>(&u);
    uint4 r4;
    __half2 *ro = reinterpret_cast<__half2 *>(&r4);
#pragma unroll
    for (int e = 0; e < 4; ++e) ro[e] = __hmax2(pa[e], pb[e]);
    return r4;
}

// ---- head: one pixel's K logits -> probabilities / votes / class id ---------------------------------------------
// The arithmetic after the logits is the ONE definition every path shares (pixel_activation, decide, argmax_step):
// mode 0 is what .predict returns, modes 1 / 2 are exactly `decide(p)` / `np.argmax(p)` of those probabilities.
template <int K>
__device__ __forceinline__ void head_pixel(const BtArgs &a, float (&p)[K], long long pix) {
    if (a.head_mode == 0) {
        pixel_activation<K>(p, K, a.head_act);
        float *dst = a.head_probs + pix * K;
#pragma unroll
        for (int k = 0; k < K; ++k) dst[k] = p[k];
    } else if (a.head_mode == 1) {
        uint32_t bits = 0;
        if (a.head_act == IMK_ACT_SIGMOID && a.head_dstar > 0.f) {
            // threshold of the sigmoid without its division: RN(1 / d) is monotonic in d = 1 + exp(-z), so "p >= thr"
            // (or ">") is exactly "d <= dstar" with dstar found on the host by exact fp32 division
#pragma unroll
            for (int k = 0; k < K; ++k) bits |= (__fadd_rn(1.0f, __expf(-p[k])) <= a.head_dstar ? 1u : 0u) << k;
        } else {
            pixel_activation<K>(p, K, a.head_act);
#pragma unroll
            for (int k = 0; k < K; ++k) bits |= decide(p[k], a.head_thr, a.head_strict != 0) << k;
        }
        a.head_dec[pix] = (uint8_t)bits;
    } else {
        pixel_activation<K>(p, K, a.head_act);
        int arg = 0;
        float best = p[0];
#pragma unroll
        for (int k = 1; k < K; ++k) argmax_step(p[k], k, best, arg);
        a.head_dec[pix] = (uint8_t)arg;
    }
}

// logits of one pixel from 16 finished channels (packed fp16, as they would have been stored): p[k] += sum_c x_c * w[k][c]
// in channel order with fp32 FMAs -- weights are constant-bank operands of the FFMA itself
template <int K, int CH>
__device__ __forceinline__ void head_fma16(const BtArgs &a, const uint4 &lo, const uint4 &hi, float (&p)[K]) {
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float2 x = __half22float2(*reinterpret_cast<const __half2 *>(&w[q]));
#pragma unroll
        for (int k = 0; k < K; ++k) {
            p[k] = __fmaf_rn(x.x, a.hw[k][16 * CH + 2 * q], p[k]);
            p[k] = __fmaf_rn(x.y, a.hw[k][16 * CH + 2 * q + 1], p[k]);
        }
    }
}

// E3 of the head variant for one epilogue warp: S3 accumulators -> c9 row (ReLU + BN, fp16) -> logits -> output.
// NCH = 1: two BLOCKS per TMEM wait; NCH = 2: the two 16-channel chunks of one block per wait (as epi_stage does).
template <int NCH, int K, int G>
__device__ __forceinline__ void head_stage(const BtArgs &a, uint32_t tmem, int g, int q, int lane, uint64_t *acc_full, uint32_t parity,
                                           int n, int y0, int x0) {
    const uint32_t lane_base = ((uint32_t)(q * 32)) << 16;
    const long long img0 = (long long)n * a.H * a.W;
    auto finish = [&](float (&p)[K], int b) {
        const int m = b * 128 + q * 32 + lane;
        const int ro = (int)__umulhi((unsigned)m, a.pitch_magic), co = m - ro * a.pitch;
        const int y = y0 + ro, x = x0 + co;
        if (ro < a.Th && co < a.Tw && y < a.H && x < a.W) head_pixel<K>(a, p, img0 + (long long)y * a.W + x);
    };
    auto init = [&](float (&p)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) p[k] = a.hb[k];
    };
    const int nb = a.s3.nb;
    if constexpr (NCH == 1) {
        for (int b = g; b < nb; b += 2 * G) {
            const bool two = b + G < nb;                             // warp-uniform
            mbar_wait(&acc_full[two ? b + G : b], parity);           // blocks complete in order
            __syncwarp();
            tc_fence_after();
            uint32_t r0[16], r1[16];
            tc_ld16(tmem + lane_base + (uint32_t)(a.s3.col + b * a.s3.n), r0);
            if (two) tc_ld16(tmem + lane_base + (uint32_t)(a.s3.col + (b + G) * a.s3.n), r1);
            tc_wait_ld();
            uint4 lo, hi;
            float p[K];
            epi16<2, 0>(a, r0, true, lo, hi);
            init(p);
            head_fma16<K, 0>(a, lo, hi, p);
            finish(p, b);
            if (two) {
                epi16<2, 0>(a, r1, true, lo, hi);
                init(p);
                head_fma16<K, 0>(a, lo, hi, p);
                finish(p, b + G);
            }
        }
    } else {
        for (int b = g; b < nb; b += G) {
            mbar_wait(&acc_full[b], parity);
            __syncwarp();
            tc_fence_after();
            uint32_t r0[16], r1[16];
            const uint32_t t0 = tmem + lane_base + (uint32_t)(a.s3.col + b * a.s3.n);
            tc_ld16(t0, r0);
            tc_ld16(t0 + 16, r1);
            tc_wait_ld();
            uint4 lo, hi;
            float p[K];
            init(p);
            epi16<2, 0>(a, r0, true, lo, hi);
            head_fma16<K, 0>(a, lo, hi, p);
            epi16<2, 1>(a, r1, true, lo, hi);
            head_fma16<K, 1>(a, lo, hi, p);
            finish(p, b);
        }
    }
}

}  // namespace

// kHead: the level-0 decoder + head variant (E3 differs); a separate instantiation, so that the other blocks' code --
// and its instruction-cache footprint: measured, a head path compiled into the common kernel cost every block 3-8 % --
// is exactly the plain three-stage pipeline
// KIND = load_kind (0 FRONT, 1 ENC, 2 DEC, 3 FRONT with the input block on the loaders) and the width CLASSES of the block
// (0: 8 channels (kin8 / n8), c >= 1: 16 c channels, -1: read at run time -- then all four are -1): CI = the loaded operand,
// N1 / N2 / N3 = the outputs of S1 / S2 / S3 are compile-time: an instantiation holds ONE loader, ONE epilogue shape per
// stage and ONE unrolled issue loop per stage.  Measured (ncu, r2o): the single MMA-issuing warp of the
// all-in-one kernel (19.6k instructions) spent 22 % of its time waiting for instruction fetches and only 18 % throttled by
// the tensor pipe -- the pipe was being starved by its own feeder.
template <int kBtEpiWarps, int kBtLoadWarps, bool kHead, int KIND, int CI, int N1, int N2, int N3>
__global__ void __launch_bounds__(bt_threads(kBtEpiWarps, kBtLoadWarps), kBtEpiWarps == 16 ? 1 : 2)
block_tc_kernel(const __grid_constant__ BtArgs a) {
    constexpr int kBtEpiGroups = kBtEpiWarps / 4;
    constexpr bool kHasS1 = KIND == 0 || KIND == 2;
    constexpr int kIn2 = kHasS1 ? N1 : CI, kIn3 = N2;             // width class of the operand S2 / S3 read
    constexpr int kBtThreads = bt_threads(kBtEpiWarps, kBtLoadWarps);
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *A0 = smem + a.a0_off, *A1 = smem + a.a1_off, *A2 = smem + a.a2_off, *OT = smem + a.o_off;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + a.bar_off);
    // ld_full / ld_empty / tma_full exist per loader buffer: the chain of three has ONE loader buffer (A0), the chain of
    // two loads straight into the double-buffered A1
    uint64_t *ld_full = bars, *ld_empty = bars + 2, *tma_full = bars + 4, *e1_done = bars + 6, *e2_done = bars + 7, *e3_done = bars + 8;
    uint64_t *o_free = bars + 9, *u_full = bars + 10;               // u_full[2]: raw uint8 tiles staged by TMA (FRONT)
    uint64_t *acc1_full = bars + 12, *acc2_full = acc1_full + kBtMaxBlocks, *acc3_full = acc2_full + kBtMaxBlocks;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + kBtNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n_my = (a.n_tiles - (long long)blockIdx.x + gridDim.x - 1) / gridDim.x;

    // ---- one-time setup ---------------------------------------------------------------------
    if (tid == 0) {
        for (int j = 0; j < 2; ++j) { mbar_init(&ld_full[j], kBtLoadWarps); mbar_init(&ld_empty[j], 1); mbar_init(&tma_full[j], 1); mbar_init(&u_full[j], 1); }
        mbar_init(o_free, 1 + (a.out_pool ? kBtLoadWarps : 0));      // store warp (+ the loader warps that pool the tile)
        mbar_init(e1_done, kBtEpiWarps); mbar_init(e2_done, kBtEpiWarps); mbar_init(e3_done, kBtEpiWarps);
        for (int b = 0; b < kBtMaxBlocks; ++b) { mbar_init(&acc1_full[b], 1); mbar_init(&acc2_full[b], 1); mbar_init(&acc3_full[b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kBtEpiWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // resident weights + parameters; operand buffers start zeroed (positions no loader / epilogue writes)
        const uint4 *src = reinterpret_cast<const uint4 *>(a.wpk);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = tid; i < a.w_bytes / 16; i += kBtThreads) dst[i] = src[i];
        uint4 *z = reinterpret_cast<uint4 *>(smem + a.a0_off);
        const int zn = (a.bar_off - a.a0_off) / 16;
        for (int i = tid; i < zn; i += kBtThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    if constexpr (KIND == 3) {                        // grayscale uint8 images only (fused_block_build)
        __syncthreads();                              // the table lives inside the region zeroed above
        {
            // one finished row of the input block per pixel value: [256][ld_cp] fp16
            const int KC1 = a.ld_cp >> 3;
            for (int t = tid; t < 256 * KC1; t += kBtThreads) {
                const int xv = t / KC1, kc = t - xv * KC1;
                const float xf = __fdiv_rn((float)xv, 255.0f);
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ch = kc * 8 + 2 * j;
                    float v0 = __fmaf_rn(xf, a.fw[0][ch], a.fb[ch]), v1 = __fmaf_rn(xf, a.fw[0][ch + 1], a.fb[ch + 1]);
                    v0 = fminf(fmaxf(v0, a.flo[ch]), a.fhi[ch]); v1 = fminf(fmaxf(v1, a.flo[ch + 1]), a.fhi[ch + 1]);
                    const __half2 h = __floats2half2_rn(v0, v1);
                    o[j] = *reinterpret_cast<const uint32_t *>(&h);
                }
                reinterpret_cast<uint4 *>(smem + a.lut_off)[t] = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
    }
    if constexpr (KIND == 0) {                        // x/255 as fp16 hi + lo (the same arithmetic the float path uses)
        __syncthreads();                              // the table lives inside the region zeroed above
        if (tid < 256) {
            const float xf = __fdiv_rn((float)tid, 255.0f);
            const __half h = __float2half_rn(xf);
            const __half l = __float2half_rn(xf - __half2float(h));
            reinterpret_cast<uint32_t *>(smem + a.lut_off)[tid] = (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16);
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < kBtEpiWarps) {
        // =====================================================================================
        //  epilogue warps
        // =====================================================================================
        const int q = warp & 3, g = warp >> 2;
        const uint32_t lane_base = ((uint32_t)(q * 32)) << 16;
        // ---- E1(j): S1 accumulators -> ReLU + BN, zero outside the image -> A1[j & 1] (haloed flat layout)
        auto E1 = [&](long long j) {
            int n, y0, x0;
            tile_coords(a, (long long)blockIdx.x + j * gridDim.x, n, y0, x0);
            const uint32_t par_ = (uint32_t)(j & 1);
            uint8_t *A1j = A1 + (size_t)(j & 1) * a.a1_stride;
            if (warp == 0) BT_TL(0, j, 0);
            if (warp == 5) BT_TL(5, j, 0);
            with_nch<N1>(a.s1, [&](auto nch) {
                epi_stage<0, decltype(nch)::value, true, kBtEpiGroups>(a, a.s1.nb, g, acc1_full, par_, (size_t)a.Pn1 * 16, a.cgm1, [&](int b) {
                    const int m = b * 128 + q * 32 + lane;
                    const int r = (int)__umulhi((unsigned)m, a.pitch_magic), c = m - r * a.pitch;
                    const int y = y0 - 1 + r, x = x0 - 1 + c;
                    const bool inside = r < a.Th + 2 && y >= 0 && y < a.H && x >= 0 && x < a.W;
                    return EpiBlk{tmem + lane_base + (uint32_t)(a.s1.col + b * a.s1.cs), inside, true, A1j + (size_t)m * 16};
                });
            });
            // warps without a block of their own still pace themselves on the stage (a free-running warp would
            // arrive on e1_done for FUTURE tiles and corrupt the phase counts)
            mbar_wait(&acc1_full[a.s1.nb - 1], par_);
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(e1_done);
            if (warp == 0) BT_TL(0, j, 1);
            if (warp == 5) BT_TL(5, j, 1);
        };
        // ---- E2(j): S2 accumulators -> ReLU -> A2 (flat layout, every row written)
        auto E2 = [&](long long j) {
            const uint32_t par_ = (uint32_t)(j & 1);
            if (warp == 0) BT_TL(0, j, 2);
            if (warp == 5) BT_TL(5, j, 2);
            with_nch<N2>(a.s2, [&](auto nch) {
                epi_stage<1, decltype(nch)::value, true, kBtEpiGroups>(a, a.s2.nb, g, acc2_full, par_, (size_t)a.Pn2 * 16, 0, [&](int b) {
                    const int m = b * 128 + q * 32 + lane;
                    return EpiBlk{tmem + lane_base + (uint32_t)(a.s2.col + b * a.s2.cs), true, true, A2 + (size_t)m * 16};
                });
            });
            mbar_wait(&acc2_full[a.s2.nb - 1], par_);
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(e2_done);
            if (warp == 0) BT_TL(0, j, 4);
            if (warp == 5) BT_TL(5, j, 4);
        };
        // ---- E3(j): S3 accumulators -> ReLU + BN -> output tile in shared memory ([Th][Tw][C] dense); the store
        //      warp ships it with row-wise bulk copies, so no epilogue warp ever waits on a global store
        auto E3 = [&](long long j) {
            const uint32_t par_ = (uint32_t)(j & 1);
            if (warp == 0) BT_TL(0, j, 5);
            if (warp == 5) BT_TL(5, j, 5);
            if constexpr (kHead) {
                int n, y0, x0;
                tile_coords(a, (long long)blockIdx.x + j * gridDim.x, n, y0, x0);
                auto run = [&](auto nch, auto kk) {
                    head_stage<decltype(nch)::value, decltype(kk)::value, kBtEpiGroups>(a, tmem, g, q, lane, acc3_full, par_, n, y0, x0);
                };
                using std::integral_constant;
                const int sel = (a.s3.n >> 4) * 4 + a.head_K;                 // (channel chunks, K): fused_block_build admits 1..2 x 1..3
                switch (sel) {
                    case 5: run(integral_constant<int, 1>{}, integral_constant<int, 1>{}); break;
                    case 6: run(integral_constant<int, 1>{}, integral_constant<int, 2>{}); break;
                    case 7: run(integral_constant<int, 1>{}, integral_constant<int, 3>{}); break;
                    case 9: run(integral_constant<int, 2>{}, integral_constant<int, 1>{}); break;
                    case 10: run(integral_constant<int, 2>{}, integral_constant<int, 2>{}); break;
                    default: run(integral_constant<int, 2>{}, integral_constant<int, 3>{}); break;
                }
            } else {
            if (j >= 1) mbar_wait(o_free, (uint32_t)((j - 1) & 1));          // tile j-1 has left the staging tile
            with_nch<N3>(a.s3, [&](auto nch) {
                epi_stage<2, decltype(nch)::value, false, kBtEpiGroups>(a, a.s3.nb, g, acc3_full, par_, 0, a.cgm3, [&](int b) {
                    const int m = b * 128 + q * 32 + lane;
                    const int ro = (int)__umulhi((unsigned)m, a.pitch_magic), co = m - ro * a.pitch;
                    const bool valid = ro < a.Th && co < a.Tw;
                    return EpiBlk{tmem + lane_base + (uint32_t)(a.s3.col + b * a.s3.cs), true, valid,
                                  OT + ((size_t)(ro * a.Tw + co) * a.out_c) * 2};
                });
            });
            }
            mbar_wait(&acc3_full[a.s3.nb - 1], par_);
            if constexpr (kHead) {
                // no shared-memory operand was written and the global stores need no ordering against the tensor core:
                // neither the proxy fence nor a releasing arrive (both wait for the outstanding global stores)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_relaxed(e3_done);
            } else {
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(e3_done);
            }
            if (warp == 0) BT_TL(0, j, 7);
            if (warp == 5) BT_TL(5, j, 7);
        };
        // Schedule (see the MMA warp): while S2(i) runs, E1(i+1) fills the OTHER A1 buffer and E3(i-1) drains R3; E2(i)
        // follows S2(i) block by block.
        // With a single A1 (blocks whose weights leave no room for two) E1(i+1) must not start before S2(i) has read the
        // buffer, i.e. it follows E2(i) -- same issue order on the MMA side, less overlap.
        if (n_my > 0) {
            const bool a1_double = a.a1_stride != 0;
            if constexpr (kHasS1) E1(0);
            if constexpr (kHead) {
                // same order, but the tail E3 is an iteration of the loop: a lambda with ONE call site is inlined whatever its
                // size -- an out-of-line E3 reads its captures and the __grid_constant__ block through memory (measured: 2x)
                for (long long i = 0; i <= n_my; ++i) {
                    if constexpr (kHasS1) { if (a1_double && i + 1 < n_my) E1(i + 1); }
                    if (i >= 1) E3(i - 1);
                    if (i < n_my) E2(i);
                    if constexpr (kHasS1) { if (!a1_double && i + 1 < n_my) E1(i + 1); }
                }
            } else {
                for (long long i = 0; i < n_my; ++i) {
                    if constexpr (kHasS1) { if (a1_double && i + 1 < n_my) E1(i + 1); }
                    if (i >= 1) E3(i - 1);
                    E2(i);
                    if constexpr (kHasS1) { if (!a1_double && i + 1 < n_my) E1(i + 1); }
                }
                E3(n_my - 1);
            }
        }
    } else if (warp == kBtEpiWarps) {
        // =====================================================================================
        //  MMA issue: all lanes wait on the barriers, the tcgen05 instructions of a phase are issued
        //  from inside ONE elect.sync region
        // =====================================================================================
        const uint32_t wbase = smem_u32(smem);
        const uint32_t pitch = (uint32_t)a.pitch;
        // descriptor words: lo = addr >> 4 | LBO >> 4 << 16 ; hi = SBO >> 4 | version 1 << 14
        constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
        struct StageRegs { uint32_t a_lo, a_step, b_lo, b_unit, idesc, d, n, cs, nb, ksteps, kin8; };
        auto make_stage = [&](const BtStage &st, uint32_t abase, int Pn) {
            StageRegs r;
            r.a_lo = (abase >> 4) | (st.kin8 ? 0u : (uint32_t)Pn << 16);   // LBO = Pn * 16 bytes (kin8: set per MMA, 0 for a 1x1 stage)
            r.kin8 = (uint32_t)st.kin8;
            r.a_step = 2u * (uint32_t)Pn;                             // next K step: two 8-channel planes further
            r.b_lo = ((wbase + (uint32_t)st.w_off) >> 4) | ((uint32_t)st.n << 16);   // LBO = n * 16 bytes
            r.b_unit = (uint32_t)st.n * 2u;                           // n * 32 bytes per (tap, K step)
            r.idesc = (1u << 4) | ((uint32_t)(st.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            r.d = tmem + (uint32_t)st.col;
            r.n = (uint32_t)st.n; r.cs = (uint32_t)st.cs; r.nb = (uint32_t)st.nb; r.ksteps = (uint32_t)st.ksteps;
            return r;
        };
        const uint32_t cgm1 = (uint32_t)a.cgm1, cgm3 = (uint32_t)a.cgm3;
        const StageRegs S1 = make_stage(a.s1, smem_u32(A0), a.Pn0);
        const StageRegs S2 = make_stage(a.s2, smem_u32(A1), a.Pn1);
        const StageRegs S3 = make_stage(a.s3, smem_u32(A2), a.Pn2);
        auto mma = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
            tc_mma_f16(d, ((uint64_t)kDescHi << 32) | a_lo, ((uint64_t)kDescHi << 32) | b_lo, idesc, acc);
        };
        // one M block of a 1x1 stage.  KS > 0: compile-time K steps (fully unrolled); KS == 0: runtime loop.
        auto block_1x1 = [&](auto ks_tag, const StageRegs &r, uint32_t b) {
            constexpr int KS = decltype(ks_tag)::value;
            const uint32_t d = r.d + b * r.cs;
            uint32_t al = r.a_lo + b * 128u, bl = r.b_lo;
            if constexpr (KS > 0) {
#pragma unroll
                for (int j = 0; j < KS; ++j) mma(d, al + (uint32_t)j * r.a_step, bl + (uint32_t)j * r.b_unit, r.idesc, (uint32_t)j);
            } else {
                for (uint32_t j = 0; j < r.ksteps; ++j, al += r.a_step, bl += r.b_unit) mma(d, al, bl, r.idesc, j);
            }
        };
        // one M block of the 3x3 stage: 9 taps x K steps, the taps are offsets dy * pitch + dx into the flat tile
        auto block_3x3 = [&](auto ks_tag, const StageRegs &r, uint32_t b) {
            constexpr int KS = decltype(ks_tag)::value;
            const uint32_t d = r.d + b * r.cs;
            const uint32_t arow0 = r.a_lo + b * 128u;
            if constexpr (KS < 0) {
                // single 8-channel plane: K = 16 of an MMA = the 8 channels of TWO taps, LBO = the distance between them
                mma(d, arow0 | (1u << 16), r.b_lo, r.idesc, 0u);                                              // (0,0) (0,1)
                mma(d, (arow0 + 2u) | ((pitch - 2u) << 16), r.b_lo + r.b_unit, r.idesc, 1u);                  // (0,2) (1,0)
                mma(d, (arow0 + pitch + 1u) | (1u << 16), r.b_lo + 2u * r.b_unit, r.idesc, 1u);               // (1,1) (1,2)
                mma(d, (arow0 + 2u * pitch) | (1u << 16), r.b_lo + 3u * r.b_unit, r.idesc, 1u);               // (2,0) (2,1)
                mma(d, arow0 + 2u * pitch + 2u, r.b_lo + 4u * r.b_unit, r.idesc, 1u);                         // (2,2) against zeros
            } else if constexpr (KS > 0) {
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const uint32_t arow = arow0 + (uint32_t)dy * pitch;
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                        for (int j = 0; j < KS; ++j)
                            mma(d, arow + (uint32_t)dx + (uint32_t)j * r.a_step, r.b_lo + (uint32_t)((dy * 3 + dx) * KS + j) * r.b_unit, r.idesc,
                                (uint32_t)((dy | dx | j) != 0));
                    }
                }
            } else {
                uint32_t bl = r.b_lo, acc = 0, arow = arow0;
                for (int dy = 0; dy < 3; ++dy, arow += pitch)
                    for (int dx = 0; dx < 3; ++dx) {
                        uint32_t al = arow + (uint32_t)dx;
                        for (uint32_t j = 0; j < r.ksteps; ++j, al += r.a_step, bl += r.b_unit) { mma(d, al, bl, r.idesc, acc); acc = 1; }
                    }
            }
        };
        // K steps of a stage whose operand has width class CLS (8 channels: one step; -1: run time)
        auto with_ks = [&](auto cls, uint32_t ksteps, auto &&f) {
            constexpr int CLS = decltype(cls)::value;
            if constexpr (CLS >= 0) { f(std::integral_constant<int, (CLS == 0 ? 1 : CLS)>{}); return; }
            switch (ksteps) {
                case 1: f(std::integral_constant<int, 1>{}); break;
                case 2: f(std::integral_constant<int, 2>{}); break;
                case 3: f(std::integral_constant<int, 3>{}); break;
                case 4: f(std::integral_constant<int, 4>{}); break;
                default: f(std::integral_constant<int, 0>{}); break;
            }
        };
        auto with_ks1 = [&](uint32_t ksteps, auto &&f) {                  // S1 of FRONT is the image stage: always one K step
            if constexpr (KIND == 0) f(std::integral_constant<int, 1>{}); else with_ks(std::integral_constant<int, CI>{}, ksteps, f);
        };
        auto issue_s1 = [&](long long ti) {
            if (elect_one()) {
                with_ks1(S1.ksteps, [&](auto ks) {
                    for (uint32_t b = 0; b < S1.nb; ++b) { BT_TLX(3, ti, b); block_1x1(ks, S1, b); if ((b & cgm1) == cgm1 || b + 1 == S1.nb) tc_commit(&acc1_full[b]); }
                });
                tc_commit(&ld_empty[0]);
            }
            __syncwarp();
        };
        auto issue_s3 = [&](long long ti) {
            if (elect_one()) {
                with_ks(std::integral_constant<int, kIn3>{}, S3.ksteps, [&](auto ks) {
                    for (uint32_t b = 0; b < S3.nb; ++b) { BT_TLX(4, ti, b); block_1x1(ks, S3, b); if ((b & cgm3) == cgm3 || b + 1 == S3.nb) tc_commit(&acc3_full[b]); }
                });
            }
            __syncwarp();
        };
        // Issue order per iteration:  S1(i+1)  S3(i-1)  S2(i).  The tensor pipe executes in order, so
        //   * E1(i+1) (fills A1[(i+1)&1]) and E3(i-1) run while S2(i) (reads A1[i&1]) executes,
        //   * E2(i) starts on S2(i)'s first finished block, i.e. after S3(i-1) has stopped reading the single A2,
        //   * nothing the pipe needs next waits on an epilogue that has not been running for a whole stage already.
        if (kHasS1 && n_my > 0) {
            mbar_wait(&ld_full[0], 0);
            tc_fence_after();
            issue_s1(0);
        }
        for (long long i = 0; i < n_my; ++i) {
            const uint32_t par_ = (uint32_t)(i & 1);
            BT_TL(1, i, 0);
            if (kHasS1) {
                mbar_wait(e1_done, par_);                         // A1[i&1] is complete, R1 is free
                tc_fence_after();
                BT_TL(1, i, 1);
                if (i + 1 < n_my) {
                    mbar_wait(&ld_full[0], (uint32_t)((i + 1) & 1));
                    tc_fence_after();
                    issue_s1(i + 1);
                }
            }
            BT_TL(1, i, 2);
            if (i >= 1) {                                         // S3(i-1): A2 complete and R2 drained; R3 drained by E3(i-2)
                mbar_wait(e2_done, par_ ^ 1u);
                if (i >= 2) mbar_wait(e3_done, par_);
                tc_fence_after();
                BT_TL(1, i, 3);
                issue_s3(i - 1);
            }
            BT_TL(1, i, 4);
            if (!kHasS1) {
                mbar_wait(&ld_full[i & 1], (uint32_t)((i >> 1) & 1));
                tc_fence_after();
            }
            BT_TL(1, i, 5);
            if (elect_one()) {
                StageRegs S2i = S2;
                S2i.a_lo += (uint32_t)((i & 1) * a.a1_stride) >> 4;
                if (kIn2 == 0 || (kIn2 < 0 && S2i.kin8)) {
                    for (uint32_t b = 0; b < S2i.nb; ++b) { block_3x3(std::integral_constant<int, -1>{}, S2i, b); tc_commit(&acc2_full[b]); }
                } else if constexpr (kIn2 != 0) {
                    with_ks(std::integral_constant<int, kIn2>{}, S2i.ksteps, [&](auto ks) {
                        for (uint32_t b = 0; b < S2i.nb; ++b) { block_3x3(ks, S2i, b); tc_commit(&acc2_full[b]); }
                    });
                }
                if (!kHasS1) tc_commit(&ld_empty[i & 1]);
            }
            __syncwarp();
            BT_TL(1, i, 6);
        }
        if (n_my > 0) {
            mbar_wait(e2_done, (uint32_t)((n_my - 1) & 1));
            if (n_my >= 2) mbar_wait(e3_done, (uint32_t)((n_my - 2) & 1));
            tc_fence_after();
            issue_s3(n_my - 1);
        }
    } else if (warp == kBtEpiWarps + 1 + kBtLoadWarps) {
        // =====================================================================================
        //  store warp: the finished output tile leaves shared memory as one bulk copy per image row
        // =====================================================================================
        const uint32_t row_smem = (uint32_t)(a.Tw * a.out_c * 2);
        for (long long i = 0; i < (kHead ? 0 : n_my); ++i) {          // head variant: nothing to ship, E3 wrote the output itself
            int n, y0, x0;
            tile_coords(a, (long long)blockIdx.x + i * gridDim.x, n, y0, x0);
            mbar_wait(e3_done, (uint32_t)(i & 1));
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)(min(a.Tw, a.W - x0) * a.out_c * 2);
                __half *g0 = a.out + (((long long)n * a.H + y0) * a.W + x0) * a.out_c;
                const int rows = min(a.Th, a.H - y0);
                for (int ro = 0; ro < rows; ++ro)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 :: "l"(g0 + (long long)ro * a.W * a.out_c), "r"(smem_u32(OT) + (uint32_t)ro * row_smem), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(o_free);
            }
            __syncwarp();
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    } else {
        // =====================================================================================
        //  loaders
        // =====================================================================================
#ifdef IMK_BT_ACC_BUILD
        uint32_t ld_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        unsigned long long ld_acc[5] = {0, 0, 0, 0, 0}, ld_tma = 0;
#define LD_T(k) ld_t[k] = bt_clk()
#define LD_ACC() do { ld_acc[0] += ld_t[1] - ld_t[0]; ld_acc[1] += ld_t[2] - ld_t[1]; ld_acc[2] += ld_t[3] - ld_t[2]; ld_acc[3] += ld_t[4] - ld_t[3]; ld_acc[4] += ld_t[6] - ld_t[5]; ld_t[5] = ld_t[6] = 0; } while (0)
#else
#define LD_T(k) do { } while (0)
#define LD_ACC() do { } while (0)
#endif
        const int lt = tid - (kBtEpiWarps + 1) * 32;
        constexpr int NL = kBtLoadWarps * 32;
        const int Pn = kHasS1 ? a.Pn0 : a.Pn1;
        const int npos = (a.Th + 2) * a.pitch;
        const int KC = a.ld_cp >> 3;
        const bool first = warp == kBtEpiWarps + 1;
        // MaxPooling2D 2x2 (unet.py:18) of a finished output tile straight from the staging buffer: the next level's
        // input leaves the SM together with the skip map (tile origin and extent are even, see fused_block_launch).
        // The loader warps do it between two loads; each arrives on o_free next to the store warp.
        auto pool_tile = [&](long long j) {
            int n, y0, x0;
            tile_coords(a, (long long)blockIdx.x + j * gridDim.x, n, y0, x0);
            LD_T(5);
            mbar_wait(e3_done, (uint32_t)(j & 1));
            LD_T(6);
            const int C8 = a.out_c >> 3;                                  // 16-byte vectors per pixel
            const int pw = min(a.Tw, a.W - x0) >> 1, ph = min(a.Th, a.H - y0) >> 1;
            const int Wp = a.W >> 1;
            const unsigned c8_magic = 0xFFFFFFFFu / (unsigned)C8 + 1u, pw_magic = 0xFFFFFFFFu / (unsigned)pw + 1u;
            __half *p0 = a.out_pool + (((long long)n * (a.H >> 1) + (y0 >> 1)) * Wp + (x0 >> 1)) * a.out_c;
            const int items = ph * pw * C8;
            for (int idx = lt; idx < items; idx += NL) {
                const int f = C8 == 1 ? idx : (int)__umulhi((unsigned)idx, c8_magic), cv = idx - f * C8;    // magic of 1 wraps to 0
                const int py = pw == 1 ? f : (int)__umulhi((unsigned)f, pw_magic), px = f - py * pw;
                const uint8_t *src = OT + ((size_t)((2 * py) * a.Tw + 2 * px) * a.out_c + cv * 8) * 2;
                const uint4 v0 = *reinterpret_cast<const uint4 *>(src);
                const uint4 v1 = *reinterpret_cast<const uint4 *>(src + (size_t)a.out_c * 2);
                const uint4 v2 = *reinterpret_cast<const uint4 *>(src + (size_t)a.Tw * a.out_c * 2);
                const uint4 v3 = *reinterpret_cast<const uint4 *>(src + (size_t)(a.Tw + 1) * a.out_c * 2);
                *reinterpret_cast<uint4 *>(p0 + (py * Wp + px) * a.out_c + cv * 8) = max_h8(max_h8(v0, v1), max_h8(v2, v3));
            }
            __syncwarp();
            // relaxed: o_free only says that the staging tile has been READ (the loads above have returned -- their values fed
            // the stores); a releasing arrive would first wait for the pooled global stores to be acknowledged, ~1.5k cycles
            // on the loader warps per tile (measured with the phase timers of the ACC build)
            if (lane == 0) mbar_arrive_relaxed(o_free);
        };
        const long long pool_lag = kHasS1 ? 3 : 2;                      // the tile whose E3 runs while this load is in flight
        // uint8 images (FRONT): the raw haloed tile of image bytes is staged by TMA into one of two shared-memory buffers,
        // TWO tiles ahead of its use (box rows of (Tw + 2) * c bytes in `u8_panels` panels of <= 256 bytes, zero filled
        // outside the image), so the loader warps only transform shared memory -> shared memory: no global-load latency
        // on the per-tile chain (the loader was the most loaded role of the FRONT blocks, one or two exposed latencies per tile).
        const bool u8_tma = (KIND == 0 || KIND == 3) && a.u8_tma;
        auto u8_issue = [&](long long ti) {                              // one thread
            int n_, y0_, x0_;
            tile_coords(a, (long long)blockIdx.x + ti * gridDim.x, n_, y0_, x0_);
            const int b = (int)(ti & 1);
            mbar_expect_tx(&u_full[b], (uint32_t)(a.u8_panels * (a.Th + 2) * a.u8_pw));
            const uint32_t dst = smem_u32(smem + a.u8_off) + (uint32_t)(b * a.u8_bstride);
            for (int p = 0; p < a.u8_panels; ++p)
                tma_load_3d(dst + (uint32_t)(p * a.u8_pstride), &a.tm_in, (((x0_ - 1) * a.in_c) & ~15) + p * a.u8_pw, y0_ - 1, n_, &u_full[b]);   // 16-byte aligned box start
        };
        if constexpr (KIND == 0 || KIND == 3) {
            if (u8_tma && first && n_my > 0) {
                if (elect_one()) { u8_issue(0); if (n_my > 1) u8_issue(1); }
                __syncwarp();
            }
        }
        for (long long i = 0; i < n_my; ++i) {
            int n, y0, x0;
            tile_coords(a, (long long)blockIdx.x + i * gridDim.x, n, y0, x0);
            // chain of three: one buffer (A0), free once S1(i-1) has read it; chain of two: A1[i & 1], free once S2(i-2) has
            const int lb = kHasS1 ? 0 : (int)(i & 1);
            const uint32_t lph = kHasS1 ? (uint32_t)(i & 1) : (uint32_t)((i >> 1) & 1);
            uint8_t *buf = kHasS1 ? A0 : A1 + (size_t)lb * a.a1_stride;
            if (first) BT_TL(2, i, 0);
            LD_T(0);
            if (kHasS1 ? i >= 1 : i >= 2) mbar_wait(&ld_empty[lb], lph ^ 1u);
            LD_T(1);
            if (first) BT_TL(2, i, 1);
            const uint8_t *U = smem + a.u8_off + (size_t)(i & 1) * a.u8_bstride;
            const int u8_adj = ((x0 - 1) * a.in_c) & 15;                 // the box starts at the 16-byte boundary at or below the tile's first byte
            if constexpr (KIND == 0 || KIND == 3) {
                if (u8_tma) {
                    if (first && i >= 1 && i + 1 < n_my) {              // U[(i+1)&1] was read for tile i-1: all loader warps are past it
                        if (elect_one()) {
                            const long long j = i - 1;
                            if (kHasS1) mbar_wait(&ld_full[0], (uint32_t)(j & 1));
                            else mbar_wait(&ld_full[j & 1], (uint32_t)((j >> 1) & 1));
                            u8_issue(i + 1);
                        }
                        __syncwarp();
                    }
                    mbar_wait(&u_full[i & 1], (uint32_t)((i >> 1) & 1));
                }
            }
            if constexpr (KIND == 0) {
                // image -> x/255 split into fp16 hi + lo so that the first layer keeps ~22 bits of the input and
                // of the weights: K slots [hi(c) | lo(c) | hi(c)] against [w_hi | w_hi | w_lo]
                const uint32_t *lut = reinterpret_cast<const uint32_t *>(smem + a.lut_off);
                if (u8_tma) {
                    const int s0 = a.swap_rb ? 2 : 0, s2 = a.swap_rb ? 0 : 2;
                    for (int f = lt; f < npos; f += NL) {
                        const int r = (int)__umulhi((unsigned)f, a.pitch_magic), cp_ = f - r * a.pitch;
                        const int b = cp_ * a.in_c + u8_adj;             // byte of the pixel's first channel in the (aligned-down) box row
                        auto byte_at = [&](int bb) -> uint32_t {         // a pixel may straddle two panels: locate every byte
                            const int pnl = (int)__umulhi((unsigned)bb, a.u8_pw_magic);
                            return U[(size_t)pnl * a.u8_pstride + (size_t)r * a.u8_pw + (bb - pnl * a.u8_pw)];
                        };
                        uint4 w0 = make_uint4(0, 0, 0, 0), w1 = make_uint4(0, 0, 0, 0);
                        if (a.in_c == 1) {                                // outside the image the box holds zeros: x = 0 -> a zero row
                            const uint32_t e = lut[byte_at(b)];
                            w0.x = e; w0.y = e & 0xFFFFu;
                        } else {
                            const uint32_t e0 = lut[byte_at(b + s0)], e1 = lut[byte_at(b + 1)], e2 = lut[byte_at(b + s2)];
                            const uint32_t h0 = e0 & 0xFFFFu, h1 = e1 & 0xFFFFu, h2 = e2 & 0xFFFFu;
                            w0.x = h0 | (h1 << 16);                       // hi0 hi1
                            w0.y = h2 | (e0 & 0xFFFF0000u);               // hi2 lo0
                            w0.z = (e1 >> 16) | (e2 & 0xFFFF0000u);       // lo1 lo2
                            w0.w = h0 | (h1 << 16);                       // hi0 hi1
                            w1.x = h2;                                    // hi2
                        }
                        *reinterpret_cast<uint4 *>(buf + (size_t)f * 16) = w0;
                        *reinterpret_cast<uint4 *>(buf + ((size_t)Pn + f) * 16) = w1;
                    }
                } else if (!a.in_f32 && (a.in_c == 1 || a.in_c == 3)) {
                    const uint8_t *img = reinterpret_cast<const uint8_t *>(a.in) + (long long)n * a.H * a.W * a.in_c;
                    constexpr int PB = 6;                    // positions in flight per thread
                    for (int f0 = lt; f0 < npos; f0 += PB * NL) {
                        uint32_t px[PB][3];
#pragma unroll
                        for (int k = 0; k < PB; ++k) {
                            const int f = f0 + k * NL;
                            const int r = (int)__umulhi((unsigned)f, a.pitch_magic), c = f - r * a.pitch;
                            const int y = y0 - 1 + r, x = x0 - 1 + c;
                            const bool in = f < npos && y >= 0 && y < a.H && x >= 0 && x < a.W;
                            const uint8_t *p = img + ((long long)y * a.W + x) * a.in_c;
                            if (a.in_c == 1) {
                                px[k][0] = in ? (uint32_t)__ldg(p) : 256u; px[k][1] = 256u; px[k][2] = 256u;
                            } else {
                                const int s0 = a.swap_rb ? 2 : 0, s2 = a.swap_rb ? 0 : 2;
                                px[k][0] = in ? (uint32_t)__ldg(p + s0) : 256u;
                                px[k][1] = in ? (uint32_t)__ldg(p + 1) : 256u;
                                px[k][2] = in ? (uint32_t)__ldg(p + s2) : 256u;
                            }
                        }
#pragma unroll
                        for (int k = 0; k < PB; ++k) {
                            const int f = f0 + k * NL;
                            if (f >= npos) continue;
                            uint4 w0 = make_uint4(0, 0, 0, 0), w1 = make_uint4(0, 0, 0, 0);
                            if (px[k][0] != 256u) {
                                if (a.in_c == 1) {
                                    const uint32_t e = lut[px[k][0]];            // hi | lo << 16
                                    w0.x = e; w0.y = e & 0xFFFFu;                 // slots: hi, lo, hi
                                } else {
                                    const uint32_t e0 = lut[px[k][0]], e1 = lut[px[k][1]], e2 = lut[px[k][2]];
                                    const uint32_t h0 = e0 & 0xFFFFu, h1 = e1 & 0xFFFFu, h2 = e2 & 0xFFFFu;
                                    w0.x = h0 | (h1 << 16);                       // hi0 hi1
                                    w0.y = h2 | (e0 & 0xFFFF0000u);               // hi2 lo0
                                    w0.z = (e1 >> 16) | (e2 & 0xFFFF0000u);       // lo1 lo2
                                    w0.w = h0 | (h1 << 16);                       // hi0 hi1
                                    w1.x = h2;                                    // hi2
                                }
                            }
                            *reinterpret_cast<uint4 *>(buf + (size_t)f * 16) = w0;
                            *reinterpret_cast<uint4 *>(buf + ((size_t)Pn + f) * 16) = w1;
                        }
                    }
                } else {
                    // generic path: float32 images (benchmark_hela, functions.py:1199) or 2 / 4 channels
                    const uint8_t *img = reinterpret_cast<const uint8_t *>(a.in) + (long long)n * a.H * a.W * a.in_c;
                    const float *imgf = reinterpret_cast<const float *>(a.in) + (long long)n * a.H * a.W * a.in_c;
                    for (int f = lt; f < npos; f += NL) {
                        const int r = (int)__umulhi((unsigned)f, a.pitch_magic), c = f - r * a.pitch;
                        const int y = y0 - 1 + r, x = x0 - 1 + c;
                        __align__(16) __half v[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] = __float2half_rn(0.f);
                        if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
                            const long long px = ((long long)y * a.W + x) * a.in_c;
                            for (int ch = 0; ch < a.in_c; ++ch) {
                                const int src = (a.swap_rb && a.in_c == 3) ? 2 - ch : ch;
                                const float xf = __fdiv_rn(a.in_f32 ? imgf[px + src] : (float)img[px + src], 255.0f);
                                const __half h = __float2half_rn(xf);
                                const __half l = __float2half_rn(xf - __half2float(h));
                                v[ch] = h; v[a.in_c + ch] = l; v[2 * a.in_c + ch] = h;
                            }
                        }
                        *reinterpret_cast<uint4 *>(buf + (size_t)f * 16) = reinterpret_cast<const uint4 *>(v)[0];
                        *reinterpret_cast<uint4 *>(buf + ((size_t)Pn + f) * 16) = reinterpret_cast<const uint4 *>(v)[1];
                    }
                }
            } else if constexpr (KIND == 3) {
                // input block through the table: a grayscale uint8 pixel selects the finished fp16 row of the 3x3 stage's
                // operand (zero outside the image: Conv2D 'same' pads the map the 3x3 reads, unet.py:12)
                const uint8_t *img = reinterpret_cast<const uint8_t *>(a.in) + (long long)n * a.H * a.W;
                if (u8_tma) {
                    const uint4 *lut4 = reinterpret_cast<const uint4 *>(smem + a.lut_off);
                    for (int f = lt; f < npos; f += NL) {
                        const int r = (int)__umulhi((unsigned)f, a.pitch_magic), cp_ = f - r * a.pitch;
                        const int y = y0 - 1 + r, x = x0 - 1 + cp_;
                        const bool in = y >= 0 && y < a.H && x >= 0 && x < a.W;      // pixel 0 is NOT a zero row here: test the position
                        const int bb = cp_ + u8_adj;
                        const int pnl = (int)__umulhi((unsigned)bb, a.u8_pw_magic);
                        const uint32_t v = U[(size_t)pnl * a.u8_pstride + (size_t)r * a.u8_pw + (bb - pnl * a.u8_pw)];
                        for (int kc = 0; kc < KC; ++kc)
                            *reinterpret_cast<uint4 *>(buf + ((size_t)kc * Pn + f) * 16) = in ? lut4[v * KC + kc] : make_uint4(0, 0, 0, 0);
                    }
                } else {
                    const uint4 *lut4 = reinterpret_cast<const uint4 *>(smem + a.lut_off);
                    constexpr int PB = 8;                    // positions in flight per thread
                    for (int f0 = lt; f0 < npos; f0 += PB * NL) {
                        uint32_t px[PB];
#pragma unroll
                        for (int k = 0; k < PB; ++k) {
                            const int f = f0 + k * NL;
                            const int r = (int)__umulhi((unsigned)f, a.pitch_magic), c = f - r * a.pitch;
                            const int y = y0 - 1 + r, x = x0 - 1 + c;
                            const bool in = f < npos && y >= 0 && y < a.H && x >= 0 && x < a.W;
                            px[k] = in ? (uint32_t)__ldg(img + (long long)y * a.W + x) : 256u;
                        }
#pragma unroll
                        for (int k = 0; k < PB; ++k) {
                            const int f = f0 + k * NL;
                            if (f >= npos) continue;
                            for (int kc = 0; kc < KC; ++kc)
                                *reinterpret_cast<uint4 *>(buf + ((size_t)kc * Pn + f) * 16) =
                                    px[k] != 256u ? lut4[px[k] * KC + kc] : make_uint4(0, 0, 0, 0);
                        }
                    }
                }
            } else if constexpr (KIND == 1) {
                // the haloed tile: one TMA box per 8-channel plane, zero filled outside the image
                if (first && elect_one()) {
                    mbar_expect_tx(&tma_full[lb], (uint32_t)KC * 16u * (uint32_t)npos);
                    const uint32_t dst = smem_u32(buf);
                    if (a.tm_flat8) tma_load_3d(dst, &a.tm_in, 2 * (x0 - 1), y0 - 1, n, &tma_full[lb]);
                    else for (int kc = 0; kc < KC; ++kc) tma_load_4d(dst + (uint32_t)(kc * Pn) * 16u, &a.tm_in, kc * 8, x0 - 1, y0 - 1, n, &tma_full[lb]);
                }
                __syncwarp();
                mbar_wait(&tma_full[lb], lph);
            } else {
                // up2x(lo) + skip (unet.py:32-33).  The haloed skip tile arrives like an ENC tile: one TMA box per
                // 8-channel plane straight into the operand layout (zero filled outside the image).  While it is in
                // flight every loader thread fetches its share of the half-resolution tile into registers (an item is
                // one 16-byte (row, column, plane) vector, plane fastest: coalesced); once the boxes have landed each
                // vector is added in place to the (up to) four positions it covers.
                if (first && elect_one()) {
                    mbar_expect_tx(&tma_full[0], (uint32_t)KC * 16u * (uint32_t)npos);
                    const uint32_t dst = smem_u32(buf);
                    if (a.tm_flat8) tma_load_3d(dst, &a.tm_in, 2 * (x0 - 1), y0 - 1, n, &tma_full[0]);
                    else for (int kc = 0; kc < KC; ++kc) tma_load_4d(dst + (uint32_t)(kc * Pn) * 16u, &a.tm_in, kc * 8, x0 - 1, y0 - 1, n, &tma_full[0]);
                }
                __syncwarp();
                // Two forms of the upsample-add.  <= 16 channels (KC <= 2): walk the half-resolution items and scatter each into the
                // (up to) four positions it covers -- fewest loads.  >= 32 channels: walk the (plane, position) vectors and gather.
                // Measured (r4k, us per 512 images, scatter -> gather): 8 ch 555 -> 624, 16 ch 945 -> 965, 32 ch 459 -> 405,
                // 64 ch 1744 -> 1501.
                constexpr bool kGatherOnly = CI >= 2, kScatterOnly = CI == 0 || CI == 1;
                if (kGatherOnly || (!kScatterOnly && KC >= 4)) {
                    // One item = one 16-byte (plane, position) vector of the HALOED tile, position fastest: the thread fetches the
                    // half-resolution vector under it (four neighbours share one: L1 hits, a warp reads 256 contiguous bytes)
                    // BEFORE it waits for the boxes, then adds it in place -- independent 16-byte updates at unit stride, no
                    // scatter.  (The first version walked the half-resolution items and scattered each into its four
                    // positions: 4.7k of the loader's 7.3k cycles per tile went into that read-modify-write loop, r4j.)
                    const __half *lo_n = a.in_lo + (long long)n * (a.H >> 1) * (a.W >> 1) * a.ld_cp;
                    const int Wl = a.W >> 1;
                    const int n_items = npos * KC;
                    const unsigned npos_magic = 0xFFFFFFFFu / (unsigned)npos + 1u;                  // exact for idx < 2^20
                    constexpr int PB = 6;
                    bool landed = false;
                    for (int i0 = lt; i0 < n_items; i0 += PB * NL) {
                        uint4 lv[PB];
                        int off[PB];                                              // 16-byte slot in the operand buffer, -1: nothing to add
    #pragma unroll
                        for (int k = 0; k < PB; ++k) {
                            const int idx = i0 + k * NL;
                            const int kc = KC == 1 ? 0 : (int)__umulhi((unsigned)idx, npos_magic), f = idx - kc * npos;
                            const int r = (int)__umulhi((unsigned)f, a.pitch_magic), c = f - r * a.pitch;
                            const int y = y0 - 1 + r, x = x0 - 1 + c;
                            const bool ok = idx < n_items && y >= 0 && y < a.H && x >= 0 && x < a.W;   // outside the image: zero stays zero
                            off[k] = ok ? kc * Pn + f : -1;
                            lv[k] = make_uint4(0, 0, 0, 0);
                            if (ok) lv[k] = __ldg(reinterpret_cast<const uint4 *>(lo_n + ((long long)(y >> 1) * Wl + (x >> 1)) * a.ld_cp + kc * 8));
                        }
                        if (!landed) { mbar_wait(&tma_full[0], (uint32_t)(i & 1)); landed = true; }
    #pragma unroll
                        for (int k = 0; k < PB; ++k) {
                            if (off[k] < 0) continue;
                            uint4 *q = reinterpret_cast<uint4 *>(buf + (size_t)off[k] * 16);
                            *q = add_h8(*q, lv[k]);
                        }
                    }
                    if (!landed) mbar_wait(&tma_full[0], (uint32_t)(i & 1));
                } else {
                    const __half *lo_n = a.in_lo + (long long)n * (a.H >> 1) * (a.W >> 1) * a.ld_cp;
                    const int Hl = a.H >> 1, Wl = a.W >> 1;
                    const int yl0 = (y0 - 1) >> 1, xl0 = (x0 - 1) >> 1;               // arithmetic shifts: -1 -> -1
                    const int pl = (a.Tw >> 1) + 2, rl = (a.Th >> 1) + 2;             // half-resolution columns / rows under the haloed tile
                    const int n_items = rl * pl * KC;
                    const unsigned kc_magic = 0xFFFFFFFFu / (unsigned)KC + 1u, pl_magic = 0xFFFFFFFFu / (unsigned)pl + 1u;
                    constexpr int PB = 4;
                    bool landed = false;
                    for (int i0 = lt; i0 < n_items; i0 += PB * NL) {
                        uint4 lv[PB];
                        int pos[PB];                                              // kc << 20 | (r << 10) | c of the item, -1: nothing to add
    #pragma unroll
                        for (int k = 0; k < PB; ++k) {
                            const int idx = i0 + k * NL;
                            const int f = KC == 1 ? idx : (int)__umulhi((unsigned)idx, kc_magic), kc = idx - f * KC;   // magic of 1 wraps to 0
                            const int r = (int)__umulhi((unsigned)f, pl_magic), c = f - r * pl;
                            const int yl = yl0 + r, xl = xl0 + c;
                            const bool ok = idx < n_items && yl >= 0 && yl < Hl && xl >= 0 && xl < Wl;
                            pos[k] = ok ? ((kc << 20) | (r << 10) | c) : -1;
                            lv[k] = make_uint4(0, 0, 0, 0);
                            if (ok) lv[k] = *reinterpret_cast<const uint4 *>(lo_n + ((long long)yl * Wl + xl) * a.ld_cp + kc * 8);
                        }
                        if (!landed) { mbar_wait(&tma_full[0], (uint32_t)(i & 1)); landed = true; }
    #pragma unroll
                        for (int k = 0; k < PB; ++k) {
                            if (pos[k] < 0) continue;
                            const int kc = pos[k] >> 20, r = (pos[k] >> 10) & 1023, c = pos[k] & 1023;
                            // rows / columns of the haloed tile covered by this half-resolution pixel
                            const int rr0 = 2 * (yl0 + r) - (y0 - 1), cc0 = 2 * (xl0 + c) - (x0 - 1);
                            uint8_t *base = buf + ((size_t)kc * Pn) * 16;
    #pragma unroll
                            for (int dy = 0; dy < 2; ++dy) {
                                const int rr = rr0 + dy;
                                if (rr < 0 || rr >= a.Th + 2) continue;
    #pragma unroll
                                for (int dx = 0; dx < 2; ++dx) {
                                    const int cc = cc0 + dx;
                                    if (cc < 0 || cc >= a.pitch) continue;
                                    uint4 *q = reinterpret_cast<uint4 *>(base + (size_t)(rr * a.pitch + cc) * 16);
                                    *q = add_h8(*q, lv[k]);
                                }
                            }
                        }
                    }
                    if (!landed) mbar_wait(&tma_full[0], (uint32_t)(i & 1));
                }
            }
            LD_T(2);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ld_full[lb]);
            if (first) BT_TL(2, i, 2);
            LD_T(3);
            if (a.out_pool && i >= pool_lag) pool_tile(i - pool_lag);
            LD_T(4);
            LD_ACC();
        }
        if (a.out_pool)
            for (long long j = n_my > pool_lag ? n_my - pool_lag : 0; j < n_my; ++j) pool_tile(j);
#ifdef IMK_BT_ACC_BUILD
        if (a.dbg && blockIdx.x == 0 && first && lane == 0) {
            for (int j = 0; j < 5; ++j) a.dbg[j] = (long long)ld_acc[j];
            a.dbg[6] = (long long)ld_tma;
            a.dbg[5] = n_my;
        }
#endif
    }
    // ---- teardown ----------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == kBtEpiWarps) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(a.tmem_cols) : "memory");
    }
}

// =============================================================================================
//  host side
// =============================================================================================
typedef void (*BtKernel)(const BtArgs);
// The instantiation for a launch shape (two: the 2-CTA/SM shape, run-time widths only), a load kind and the width
// classes (ci | n1 | n2 | n3, -1 where the kind has no such operand / stage).  The table holds the shapes U-Nets of
// unet.py produce -- FRONT: (c, c, c); ENC: c -> (2c, 2c); DEC: c -> (c, c, c) at level 0, (c, c, c/2) above -- for the
// widths the resident-weight design admits; anything else takes the run-time instantiation of its kind.
static BtKernel bt_kernel(bool two, int kind, int ci, int n1, int n2, int n3) {
    if (two) {
        switch (kind) {
            case 0: return block_tc_kernel<8, 4, false, 0, -1, -1, -1, -1>;
            case 1: return block_tc_kernel<8, 4, false, 1, -1, -1, -1, -1>;
            case 2: return block_tc_kernel<8, 4, false, 2, -1, -1, -1, -1>;
            default: return block_tc_kernel<8, 4, false, 3, -1, -1, -1, -1>;
        }
    }
#define IMK_BT_IS(K, A, B, C, D) if (kind == K && ci == A && n1 == B && n2 == C && n3 == D) return block_tc_kernel<16, 8, false, K, A, B, C, D>
    IMK_BT_IS(0, -1, 0, 0, 0); IMK_BT_IS(0, -1, 1, 1, 1); IMK_BT_IS(0, -1, 2, 2, 2);               // FRONT, image stage on the tensor cores
    IMK_BT_IS(3, 0, -1, 0, 0); IMK_BT_IS(3, 1, -1, 1, 1); IMK_BT_IS(3, 2, -1, 2, 2);               // FRONT, input block on the loaders
    IMK_BT_IS(1, 0, -1, 1, 1); IMK_BT_IS(1, 1, -1, 2, 2); IMK_BT_IS(1, 2, -1, 4, 4);               // ENC: c -> 2c
    IMK_BT_IS(2, 0, 0, 0, 0); IMK_BT_IS(2, 1, 1, 1, 1); IMK_BT_IS(2, 2, 2, 2, 2);                  // DEC level 0
    IMK_BT_IS(2, 1, 1, 1, 0); IMK_BT_IS(2, 2, 2, 2, 1); IMK_BT_IS(2, 4, 4, 4, 2);                  // DEC above: c -> c/2
#undef IMK_BT_IS
    switch (kind) {
        case 0: return block_tc_kernel<16, 8, false, 0, -1, -1, -1, -1>;
        case 1: return block_tc_kernel<16, 8, false, 1, -1, -1, -1, -1>;
        case 2: return block_tc_kernel<16, 8, false, 2, -1, -1, -1, -1>;
        default: return block_tc_kernel<16, 8, false, 3, -1, -1, -1, -1>;
    }
}

static void pack_umma_b(std::vector<__half> &dst, const float *hwio, int ks, int cin, int cout, int cin_p, int cout_p) {
    // [tap][kc = cin_p/8][cout_p][8] fp16, zero in the padding (the operand-B image of imk_conv_tc.cu)
    const int taps = ks * ks, KC = cin_p / 8;
    const size_t base = dst.size();
    dst.resize(base + (size_t)taps * KC * cout_p * 8, __float2half(0.f));
    for (int tap = 0; tap < taps; ++tap)
        for (int ci = 0; ci < cin; ++ci)
            for (int co = 0; co < cout; ++co)
                dst[base + (((size_t)tap * KC + ci / 8) * cout_p + co) * 8 + (ci & 7)] =
                    __float2half_rn(hwio[((size_t)tap * cin + ci) * cout + co]);
}

// operand B of a stage whose A operand is ONE 8-channel plane (BtStage::kin8): images of [2 kc][cout_p][8] fp16.
// 3x3: five images, image j = taps (2j, 2j + 1) stacked along K (zeros under the ninth tap); 1x1: one image, zeros in kc 1.
static void pack_umma_b8(std::vector<__half> &dst, const float *hwio, int ks, int cin, int cout, int cout_p) {
    const int taps = ks * ks, imgs = (taps + 1) / 2;
    const size_t base = dst.size();
    dst.resize(base + (size_t)imgs * 2 * cout_p * 8, __float2half(0.f));
    for (int tap = 0; tap < taps; ++tap)
        for (int ci = 0; ci < cin; ++ci)
            for (int co = 0; co < cout; ++co)
                dst[base + (((size_t)(tap / 2) * 2 + (tap & 1)) * cout_p + co) * 8 + ci] =
                    __float2half_rn(hwio[((size_t)tap * cin + ci) * cout + co]);
}

// first layer with the hi/lo split: K = 16 slots [w_hi(c) | w_hi(c) | w_lo(c) | 0...]
static void pack_front_b(std::vector<__half> &dst, const float *w /*[c][cout]*/, int c, int cout, int cout_p) {
    const size_t base = dst.size();
    dst.resize(base + (size_t)2 * cout_p * 8, __float2half(0.f));
    for (int ch = 0; ch < c; ++ch)
        for (int co = 0; co < cout; ++co) {
            const float v = w[(size_t)ch * cout + co];
            const __half hi = __float2half_rn(v);
            const __half lo = __float2half_rn(v - __half2float(hi));
            const int slots[3] = {ch, c + ch, 2 * c + ch};
            const __half vals[3] = {hi, hi, lo};
            for (int s = 0; s < 3; ++s)
                dst[base + ((size_t)(slots[s] / 8) * cout_p + co) * 8 + (slots[s] & 7)] = vals[s];
        }
}

// BN scale folded into the weights (see BtArgs::cpar): returns the scaled HWIO kernel and fills the stage's epilogue constants
static uint32_t pack_h2(float x, float y) {
    const __half2 h = __floats2half2_rn(x, y);
    uint32_t u;
    memcpy(&u, &h, 4);
    return u;
}
static std::vector<float> fold_stage(const ConvHost &L, BtArgs &a, int si) {
    const int taps = L.ks * L.ks;
    std::vector<float> w(L.hwio, L.hwio + (size_t)taps * L.cin * L.cout);
    const float inf = INFINITY;
    float lo[64], hi[64];
    for (int co = 0; co < 64; ++co) { a.cpar[si][co] = 0.f; lo[co] = 0.f; hi[co] = inf; }               // padding channels stay 0
    a.has_hi[si] = 0;
    for (int co = 0; co < L.cout; ++co) {
        const float sc = L.bn_scale ? L.bn_scale[co] : 1.f, sh = L.bn_shift ? L.bn_shift[co] : 0.f;
        for (size_t i = co; i < w.size(); i += L.cout) w[i] *= sc;
        a.cpar[si][co] = sc * L.bias[co] + sh;
        lo[co] = sc > 0.f ? sh : (sc < 0.f ? -inf : sh);
        hi[co] = sc > 0.f ? inf : sh;
        if (!(sc > 0.f)) a.has_hi[si] = 1;
    }
    for (int c2 = 0; c2 < 32; ++c2) { a.clo[si][c2] = pack_h2(lo[2 * c2], lo[2 * c2 + 1]); a.chi[si][c2] = pack_h2(hi[2 * c2], hi[2 * c2 + 1]); }
    return w;
}

static bool bt_disabled() {
    const char *v = getenv("IMK_BT_DISABLE");
    return v && v[0] && v[0] != '0';
}

static inline int round8(int v) { return (v + 7) / 8 * 8; }

struct BtGeom { int nb1, nb2, Pn0, Pn1, Pn2, cols; size_t bytes; };
// TMEM columns between the M blocks of a stage / columns of the whole stage (IMK_BT_NO_CS8=1: always n, for A/B runs)
static inline int bt_cs(const BtStage &st) {
    static int off = -1;
    if (off < 0) { const char *v = getenv("IMK_BT_NO_CS8"); off = (v && v[0] == '1') ? 1 : 0; }
    return (st.n8 && !off) ? 8 : st.n;
}
static inline int bt_span(const BtStage &st, int nb) { return nb == 0 ? 0 : bt_cs(st) * nb + (st.n - bt_cs(st)); }
// Raw uint8 tile staging of the FRONT blocks (see u8_issue in the kernel): box rows of (tw + 2) * c bytes, split into
// panels of <= 256 bytes (the TMA box limit) whose width is a multiple of 16 bytes AND of c, so that no pixel straddles two.
struct U8Geom { int pw, panels, pstride, bytes; };
static bool bt_u8_enabled() {
    static int off = -1;
    if (off < 0) { const char *v = getenv("IMK_BT_NO_U8TMA"); off = (v && v[0] == '1') ? 1 : 0; }
    return !off;
}
static U8Geom bt_u8_geom(const BtArgs &a, int th, int tw) {
    U8Geom u{0, 0, 0, 0};
    if (!(a.load_kind == 0 || a.load_kind == 3) || !(a.in_c == 1 || a.in_c == 3) || !bt_u8_enabled()) return u;
    const int L = 16, rb = (tw + 2) * a.in_c + 15, pw_max = 256;     // + 15: the box starts at a 16-byte boundary
    u.panels = (rb + pw_max - 1) / pw_max;
    u.pw = ((rb + u.panels - 1) / u.panels + L - 1) / L * L;
    u.pstride = ((th + 2) * u.pw + 127) / 128 * 128;
    u.bytes = 2 * u.panels * u.pstride;
    return u;
}
static inline int bt_planes(const BtStage &st) { return st.kin8 ? 1 : st.ksteps * 2; }     // 16-byte planes of the stage's A operand

// shared-memory / TMEM footprint of a candidate tile; plane strides are multiples of 8 positions so that every
// plane starts 128-byte aligned (TMA destination)
static bool bt_geom(const FusedBlock &fb, int th, int tw, int cols_max, size_t smem_max, int a1_bufs, BtGeom &g) {
    const BtArgs &a = fb.args;
    const int pitch = tw + 2;
    const int n1 = a.has_s1 ? a.s1.n : 0, n2 = a.s2.n, n3 = a.s3.n;
    (void)n3;
    g.nb1 = a.has_s1 ? ((th + 2) * pitch + 127) / 128 : 0;
    g.nb2 = (th * pitch + 127) / 128;
    if (g.nb1 > kBtMaxBlocks || g.nb2 > kBtMaxBlocks) return false;
    // TMEM columns: a stage with <= 8 real outputs (n8) spaces its M blocks 8 columns apart -- the UMMA still writes
    // N = 16 columns, but the upper eight are the NEXT block's, overwritten by it before anybody reads them (the blocks of a
    // stage are issued in order and read only after the stage's MMAs have retired) -- so twice the pixels fit a tile
    g.cols = (a.has_s1 ? bt_span(a.s1, g.nb1) : 0) + bt_span(a.s2, g.nb2) + bt_span(a.s3, g.nb2);
    (void)n1; (void)n2;
    if (g.cols > cols_max) return false;
    if (pitch > 256 || th + 2 > 256) return false;                   // TMA box limits
    g.Pn0 = round8(g.nb1 * 128);
    g.Pn1 = round8(std::max(std::max(g.nb1 * 128, g.nb2 * 128 + 2 * pitch + 2), (th + 2) * pitch));
    g.Pn2 = round8(g.nb2 * 128);
    size_t off = (size_t)fb.w_bytes + (size_t)fb.par_floats * 4;
    off = (off + 127) / 128 * 128;
    if (a.has_s1) off += (size_t)g.Pn0 * bt_planes(a.s1) * 16;
    off += a1_bufs * (((size_t)g.Pn1 * bt_planes(a.s2) * 16 + 127) / 128 * 128);
    off += (size_t)g.Pn2 * bt_planes(a.s3) * 16;
    if (!a.has_head) off += ((size_t)th * tw * a.out_c * 2 + 127) / 128 * 128;  // head variant: no output staging tile
    off += (size_t)bt_u8_geom(a, th, tw).bytes;
    if (a.load_kind == 0) off += 1024;
    if (a.load_kind == 3) off += (size_t)256 * a.ld_cp * 2;
    off += (size_t)kBtNumBars * 8 + 16;
    g.bytes = off;
    return off <= smem_max;
}

// Cheapest tile for a TMEM / shared-memory budget.  Cost = tcgen05.mma instructions per image (every small-N MMA
// costs about the same: its 4 KB A operand read) plus a fixed per-tile term for the hand-offs (measured: a tile costs a few thousand cycles of
// pipeline latency whatever its size, so fewer, larger tiles win); partial
// edge tiles are charged in full.  Tiles are even-sized so that the store warp can carry the 2x2 max-pool.
static double bt_best_tile(const FusedBlock &fb, int H, int W, int cols_max, size_t smem_max, int a1_bufs, int &bTh, int &bTw) {
    const BtArgs &a = fb.args;
    double best = -1.0;
    BtGeom g{};
    for (int th = 2; th <= 16; th += 2) {
        if (th > H + 1) break;
        for (int tw = 8; tw <= 254; tw += 2) {
            if (tw > W + 1 && tw > 8) break;
            if (!bt_geom(fb, th, tw, cols_max, smem_max, a1_bufs, g)) continue;
            const double tiles = (double)((W + tw - 1) / tw) * ((H + th - 1) / th);
            const double per_tile = (a.has_s1 ? g.nb1 * a.s1.ksteps : 0) + g.nb2 * ((a.s2.kin8 ? 5.0 : 9.0 * a.s2.ksteps) + a.s3.ksteps) + 60.0;
            const double cost = tiles * per_tile;
            if (best < 0 || cost < best - 1e-9 || (cost < best + 1e-9 && th * tw > bTh * bTw)) { best = cost; bTh = th; bTw = tw; }
        }
    }
    return best;
}

// Chooses the launch shape (1 or 2 CTAs per SM) and the tile, lays out shared memory / TMEM.  Returns false when
// the block does not fit (weights too large to stay resident, or no tile satisfies the TMEM budget).
static bool bt_plan(FusedBlock &fb, int H, int W) {
    BtArgs &a = fb.args;
    a.H = H; a.W = W;
    // candidates: {1 CTA/SM, A1 double} (the pipelined schedule), {1 CTA/SM, A1 single} when the weights leave no room
    // for a decent tile with two buffers (chain of three only: the chain of two loads into A1 and needs both),
    // {2 CTAs/SM} only when nothing else fits.  Measured (HeLa, r01): two co-resident half-size pipelines do not beat
    // one full-size pipeline -- the epilogue warps of both CTAs share the same issue slots.
    int thd = 0, twd = 0, ths = 0, tws = 0, th2 = 0, tw2 = 0;
    const double cd = bt_best_tile(fb, H, W, 512, kBtSmemMax, 2, thd, twd);
    const double cs = a.has_s1 ? bt_best_tile(fb, H, W, 512, kBtSmemMax, 1, ths, tws) : -1.0;
    const double c2 = a.has_head ? -1.0 : bt_best_tile(fb, H, W, 256, kBtSmemMax2, 2, th2, tw2);   // the head variant exists for the 1-CTA shape only
    int force = 0;
    if (const char *v = getenv("IMK_BT_CTAS"); v && v[0]) force = atoi(v);
    int mode = -1;                                                    // 0: double, 1: single, 2: two CTAs
    if (const char *v = getenv("IMK_BT_TILE"); v && v[0]) {           // tuning aid: "kind:H:th:tw[;...]" pins the tile of a block
        int k_, h_, th_, tw_;
        for (const char *p = v; p && *p; p = strchr(p, ';') ? strchr(p, ';') + 1 : nullptr)
            if (sscanf(p, "%d:%d:%d:%d", &k_, &h_, &th_, &tw_) == 4 && k_ == a.load_kind && h_ == H) {
                BtGeom gg{};
                if (bt_geom(fb, th_, tw_, 512, kBtSmemMax, 2, gg)) { thd = th_; twd = tw_; mode = 0; }
                else if (a.has_s1 && bt_geom(fb, th_, tw_, 512, kBtSmemMax, 1, gg)) { ths = th_; tws = tw_; mode = 1; }
            }
    }
    const bool pinned = mode >= 0;
    if (mode < 0 && cd > 0) mode = 0;
    if (!pinned && cs > 0 && (mode < 0 || cs * 1.25 < cd)) mode = 1;  // the overlap is worth about a quarter of the tile time
    if (mode < 0 && c2 > 0) mode = 2;
    if (!pinned && force == 2 && c2 > 0) mode = 2;
    if (mode < 0) return false;
    const bool two = mode == 2;
    const int a1_bufs = mode == 1 ? 1 : 2;
    fb.ctas_per_sm = two ? 2 : 1;
    const int bTh = mode == 0 ? thd : (mode == 1 ? ths : th2), bTw = mode == 0 ? twd : (mode == 1 ? tws : tw2);
    BtGeom g{};
    bt_geom(fb, bTh, bTw, two ? 256 : 512, two ? kBtSmemMax2 : kBtSmemMax, a1_bufs, g);
    a.Th = bTh; a.Tw = bTw; a.pitch = bTw + 2;
    a.pitch_magic = (unsigned)((0x100000000ull + a.pitch - 1) / a.pitch);
    a.tiles_x = (W + bTw - 1) / bTw; a.tiles_y = (H + bTh - 1) / bTh;
    a.s1.nb = g.nb1; a.s2.nb = a.s3.nb = g.nb2;
    a.s1.cs = bt_cs(a.s1); a.s2.cs = bt_cs(a.s2); a.s3.cs = bt_cs(a.s3);
    a.s1.col = 0; a.s2.col = a.has_s1 ? bt_span(a.s1, g.nb1) : 0; a.s3.col = a.s2.col + bt_span(a.s2, g.nb2);
    a.tmem_cols = 32;
    while (a.tmem_cols < g.cols) a.tmem_cols *= 2;
    a.Pn0 = g.Pn0; a.Pn1 = g.Pn1; a.Pn2 = g.Pn2;
    size_t off = (size_t)fb.w_bytes;
    a.par_off_b = (int)off; off += (size_t)fb.par_floats * 4; off = (off + 127) / 128 * 128;
    a.a0_off = (int)off; if (a.has_s1) off += (size_t)a.Pn0 * bt_planes(a.s1) * 16;
    a.a1_off = (int)off; a.a1_stride = (int)(((size_t)a.Pn1 * bt_planes(a.s2) * 16 + 127) / 128 * 128); off += (size_t)a1_bufs * a.a1_stride;
    if (a1_bufs == 1) a.a1_stride = 0;                               // single buffer: both parities alias
    a.a2_off = (int)off; off += (size_t)a.Pn2 * bt_planes(a.s3) * 16;
    a.o_off = (int)off; if (!a.has_head) off += ((size_t)a.Th * a.Tw * a.out_c * 2 + 127) / 128 * 128;
    {
        const U8Geom u = bt_u8_geom(a, a.Th, a.Tw);
        a.u8_off = (int)off; a.u8_pw = u.pw; a.u8_panels = u.panels; a.u8_pstride = u.pstride; a.u8_bstride = u.panels * u.pstride;
        a.u8_pw_magic = u.pw ? (unsigned)((0x100000000ull + u.pw - 1) / u.pw) : 0u;
        off += (size_t)u.bytes;
    }
    a.lut_off = (int)off; if (a.load_kind == 0) off += 1024;
    if (a.load_kind == 3) off += (size_t)256 * a.ld_cp * 2;
    a.bar_off = (int)off; off += (size_t)kBtNumBars * 8 + 16;        // everything in [a0_off, bar_off) starts zeroed
    fb.smem = off;
    return off <= (size_t)(two ? kBtSmemMax2 : kBtSmemMax);
}

// ---- TMA tensor maps (driver entry point fetched through the runtime: no -lcuda) ---------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// fp16 NHWC [n, h, w, c] with boxes {8 ch, box_w, box_h, 1}; out-of-bounds elements read as zero
int make_map(CUtensorMap *map, const void *base, int64_t n, int h, int w, int c, int box_w, int box_h) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return IMK_ECUDA; }
    const cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
    const cuuint32_t box[4] = {8, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for [%lld,%d,%d,%d] box %dx%d", (int)r, (long long)n, h, w, c, box_w, box_h); return IMK_ECUDA; }
    return IMK_OK;
}

// 8-channel fp16 NHWC map as uint64 [n][h][2w] with boxes {2 box_w, box_h, 1} (see tma_load_3d)
static int make_map_flat8(CUtensorMap *map, const void *base, int64_t n, int h, int w, int box_w, int box_h) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return IMK_ECUDA; }
    const cuuint64_t dims[3] = {(cuuint64_t)w * 2, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)w * 16, (cuuint64_t)h * w * 16};
    const cuuint32_t box[3] = {(cuuint32_t)box_w * 2, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (flat 8-channel) failed (%d) for [%lld,%d,%d] box %dx%d", (int)r, (long long)n, h, w, box_w, box_h); return IMK_ECUDA; }
    return IMK_OK;
}

static int bt_upload(FusedBlock &fb, const std::vector<__half> &w, const std::vector<float> &par, std::vector<void *> &owned) {
    void *pw = nullptr, *pp = nullptr;
    if (cudaMalloc(&pw, w.size() * sizeof(__half)) != cudaSuccess) { set_error("fused block: cudaMalloc failed"); return IMK_ENOMEM; }
    owned.push_back(pw);
    if (cudaMalloc(&pp, par.size() * sizeof(float)) != cudaSuccess) { set_error("fused block: cudaMalloc failed"); return IMK_ENOMEM; }
    owned.push_back(pp);
    IMK_CUDA(cudaMemcpy(pw, w.data(), w.size() * sizeof(__half), cudaMemcpyHostToDevice));
    IMK_CUDA(cudaMemcpy(pp, par.data(), par.size() * sizeof(float), cudaMemcpyHostToDevice));
    fb.args.wpk = reinterpret_cast<const uint8_t *>(pw);
    fb.args.par = reinterpret_cast<const float *>(pp);
    return IMK_OK;
}

// kind: 0 FRONT (L = in, conv3, conv1), 1 ENC (L = conv3, conv1), 2 DEC (L = conv1a, conv3, conv1b)
int fused_block_build(FusedBlock &fb, int kind_, const ConvHost *L, int H, int W, int in_c, std::vector<void *> &owned, int head_act, bool allow8) {
    int kind = kind_;
    const bool head = kind == 4;
    if (head) kind = 2;
    fb = FusedBlock{};
    if (bt_disabled()) return IMK_OK;
    if (const char *v = getenv("IMK_BT_KINDS"); v && v[0] && !((atoi(v) >> kind) & 1)) return IMK_OK;   // debug: bitmask of fused kinds
    BtArgs &a = fb.args;
    std::vector<__half> w;
    std::vector<float> par(4, 0.f);
    auto stage = [&](BtStage &s, const ConvHost &c, bool front) {
        const int cin_p = front ? 16 : pad_ch(c.cin), n = pad_ch(c.cout);
        s.kin8 = (allow8 && !front && c.cin <= 8) ? 1 : 0;
        s.n8 = (allow8 && c.cout <= 8) ? 1 : 0;
        s.taps = front ? 1 : c.ks * c.ks; s.ksteps = cin_p / 16; s.n = n;
        s.w_off = (int)(w.size() * sizeof(__half));
        const int si = &s == &a.s1 ? 0 : (&s == &a.s2 ? 1 : 2);
        const std::vector<float> ws = fold_stage(c, a, si);
        if (front) pack_front_b(w, ws.data(), c.cin, c.cout, n);
        else if (s.kin8) pack_umma_b8(w, ws.data(), c.ks, c.cin, c.cout, n);
        else pack_umma_b(w, ws.data(), c.ks, c.cin, c.cout, cin_p, n);
        s.par_off = 0;
    };
    a.load_kind = kind; a.in_c = in_c;
    for (int j = (kind == 3 ? 1 : 0); j < (kind == 1 ? 2 : 3); ++j)
        if (pad_ch(L[j].cout) > 64) return IMK_OK;            // BtArgs::cpar holds 64 channels per stage
    if (kind == 3) {
        // grayscale only: for c = 3 the table would be three fp32 partial-sum tables and an FMA chain per channel on the
        // loader warps, measured slower than the tensor-core input stage of kind 0 (ISIC 2.3 -> 3.0 ms per 1024 images)
        if (L[0].ks != 1 || L[1].ks != 3 || L[2].ks != 1 || in_c != 1 || pad_ch(L[0].cout) > 32) return IMK_OK;
        a.has_s1 = 0;
        stage(a.s2, L[1], false); stage(a.s3, L[2], false);
        a.ld_cp = a.s2.kin8 ? 8 : pad_ch(L[0].cout);
        const float inf = INFINITY;
        for (int ch = 0; ch < 32; ++ch) { for (int c = 0; c < 4; ++c) a.fw[c][ch] = 0.f; a.fb[ch] = 0.f; a.flo[ch] = 0.f; a.fhi[ch] = inf; }
        for (int co = 0; co < L[0].cout; ++co) {
            const float sc = L[0].bn_scale ? L[0].bn_scale[co] : 1.f, sh = L[0].bn_shift ? L[0].bn_shift[co] : 0.f;
            for (int c = 0; c < in_c; ++c) a.fw[c][co] = sc * L[0].hwio[(size_t)c * L[0].cout + co];
            a.fb[co] = sc * L[0].bias[co] + sh;
            a.flo[co] = sc > 0.f ? sh : (sc < 0.f ? -inf : sh);
            a.fhi[co] = sc > 0.f ? inf : sh;
        }
    } else if (kind == 1) {
        if (L[0].ks != 3 || L[1].ks != 1) return IMK_OK;
        a.has_s1 = 0;
        stage(a.s2, L[0], false); stage(a.s3, L[1], false);
        a.ld_cp = a.s2.kin8 ? 8 : pad_ch(L[0].cin);
    } else {
        if (L[0].ks != 1 || L[1].ks != 3 || L[2].ks != 1) return IMK_OK;
        if (kind == 0 && 3 * in_c > 16) return IMK_OK;
        a.has_s1 = 1;
        stage(a.s1, L[0], kind == 0); stage(a.s2, L[1], false); stage(a.s3, L[2], false);
        a.ld_cp = kind == 0 ? 16 : (a.s1.kin8 ? 8 : pad_ch(L[0].cin));
        if (head) {
            // the output layer (unet.py:63): 1x1, C0 -> K, no ReLU / BN, fp32 weights -- evaluated in the E3 epilogue
            const ConvHost &o = L[3];
            const int K = o.cout;
            if (o.ks != 1 || K < 1 || K > kBtHeadMaxK || pad_ch(o.cin) != a.s3.n || a.s3.n > 32 || a.s3.n8) return IMK_OK;
            a.has_head = 1; a.head_K = K; a.head_act = head_act;
            for (int k = 0; k < 3; ++k) for (int ci = 0; ci < 32; ++ci) a.hw[k][ci] = (k < K && ci < o.cin) ? o.hwio[(size_t)ci * K + k] : 0.f;
            for (int k = 0; k < 4; ++k) a.hb[k] = k < K ? o.bias[k] : 0.f;
        }
    }
    a.out_c = a.s3.n8 ? 8 : a.s3.n;
    if (w.size() * sizeof(__half) % 16) w.resize((w.size() + 7) / 8 * 8, __float2half(0.f));
    fb.w_bytes = a.w_bytes = (int)(w.size() * sizeof(__half));
    fb.par_floats = a.par_floats = (int)par.size();
    if (!bt_plan(fb, H, W)) return IMK_OK;          // does not fit: the layer-wise engine runs this block
    int rc = bt_upload(fb, w, par, owned);
    if (rc) return rc;
    fb.ok = true;
    if (const char *v = getenv("IMK_BT_VERBOSE"); v && v[0] == '1')
        fprintf(stderr, "[imk] fused block kind=%d %dx%d: %d CTA/SM, A1 x%d, tile %dx%d, M blocks %d/%d, TMEM %d cols, smem %zu B, weights %d B\n", kind, H, W,
                fb.ctas_per_sm, a.a1_stride ? 2 : 1, a.Th, a.Tw, a.s1.nb, a.s2.nb, a.tmem_cols, fb.smem, fb.w_bytes);
    return IMK_OK;
}

bool fused_block_can_pool(const FusedBlock &fb) {
    const BtArgs &a = fb.args;
    return fb.ok && a.Th % 2 == 0 && a.Tw % 2 == 0 && a.H % 2 == 0 && a.W % 2 == 0;
}

int fused_block_launch(const FusedBlock &fb, const void *in, const __half *in_lo, __half *out, __half *out_pool, int64_t n,
                       int swap_rb, int in_f32, cudaStream_t stream, const HeadOut *head) {
    BtArgs a = fb.args;
    a.in = in; a.in_lo = in_lo; a.out = out; a.swap_rb = swap_rb; a.in_f32 = in_f32;
    if ((a.has_head != 0) != (head != nullptr)) { set_error("fused block: head output %s", head ? "requested from a block without a head" : "missing"); return IMK_EINVAL; }
    if (head) {
        a.head_mode = head->mode; a.head_thr = head->thr; a.head_dstar = head->dstar; a.head_strict = head->strict;
        a.head_probs = head->probs; a.head_dec = head->dec;
        if ((head->mode == 0 && !head->probs) || (head->mode != 0 && !head->dec)) { set_error("fused block: NULL head output"); return IMK_EINVAL; }
    }
    {   // commit groups of the 1x1 stages (IMK_BT_CG1 / IMK_BT_CG3 = 1, 2, 4, 8 blocks per tcgen05.commit)
        static int cg1 = -1, cg3 = -1;
        if (cg1 < 0) {
            auto rd = [](const char *name, int dflt) { const char *v = getenv(name); const int x = v && v[0] ? atoi(v) : dflt; return (x == 2 || x == 4 || x == 8) ? x - 1 : 0; };
            cg3 = rd("IMK_BT_CG3", 1);
            cg1 = rd("IMK_BT_CG1", 1);
        }
        a.cgm1 = cg1; a.cgm3 = a.has_head ? 0 : cg3;
    }
    a.out_pool = out_pool;
    if (out_pool && !fused_block_can_pool(fb)) { set_error("fused block: the tile cannot carry the 2x2 max-pool"); return IMK_EINVAL; }
    a.n_tiles = (long long)n * a.tiles_x * a.tiles_y;
    if (a.n_tiles <= 0) return IMK_OK;
    a.u8_tma = 0;
    if ((a.load_kind == 0 || a.load_kind == 3) && a.u8_panels > 0 && !in_f32 && (a.W * a.in_c) % 16 == 0) {
        // the image as uint8 [n][H][W*c]; boxes {panel width, Th + 2, 1}, zero filled outside the image
        EncodeTiledFn fn = encode_tiled_fn();
        if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return IMK_ECUDA; }
        const cuuint64_t dims[3] = {(cuuint64_t)a.W * a.in_c, (cuuint64_t)a.H, (cuuint64_t)n};
        const cuuint64_t strides[2] = {(cuuint64_t)a.W * a.in_c, (cuuint64_t)a.H * a.W * a.in_c};
        const cuuint32_t box[3] = {(cuuint32_t)a.u8_pw, (cuuint32_t)(a.Th + 2), 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = fn(&a.tm_in, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(in), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (uint8 image) failed (%d) for [%lld,%d,%d] box %dx%d", (int)r, (long long)n, a.H, a.W * a.in_c, a.u8_pw, a.Th + 2); return IMK_ECUDA; }
        a.u8_tma = 1;
    }
    if (a.load_kind == 1 || a.load_kind == 2) {
        a.tm_flat8 = (a.ld_cp == 8 && a.pitch <= 128) ? 1 : 0;
        if (const char *v = getenv("IMK_BT_NO_FLAT8"); v && v[0] == '1') a.tm_flat8 = 0;
        int rc = a.tm_flat8 ? make_map_flat8(&a.tm_in, in, n, a.H, a.W, a.pitch, a.Th + 2)
                            : make_map(&a.tm_in, in, n, a.H, a.W, a.ld_cp, a.pitch, a.Th + 2);
        if (rc) return rc;
    }
    // width classes of the block (see bt_kernel)
    int ci = -1, n1 = -1, n2 = -1, n3 = -1;
    {
        auto out_cls = [](const BtStage &st) { return st.n8 ? 0 : st.n >> 4; };
        auto in_cls = [](const BtStage &st) { return st.kin8 ? 0 : st.ksteps; };
        n2 = out_cls(a.s2); n3 = out_cls(a.s3);
        if (a.has_s1) { n1 = out_cls(a.s1); ci = a.load_kind == 0 ? -1 : in_cls(a.s1); }
        else ci = in_cls(a.s2);
        if (a.s2.n % 16 || a.s3.n % 16 || (a.has_s1 && a.s1.n % 16)) ci = -9;          // no such instantiation
        if (const char *v = getenv("IMK_BT_GENERIC"); v && v[0] == '1') ci = -9;
    }
    const BtKernel kern = a.has_head ? (BtKernel)block_tc_kernel<16, 8, true, 2, -1, -1, -1, -1>
                                     : bt_kernel(fb.ctas_per_sm == 2, a.load_kind, ci, n1, n2, n3);
    // the opt-in is per device (a process may drive several GPUs) and per instantiation: cheap, so set at every launch
    IMK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fb.ctas_per_sm == 2 ? kBtSmemMax2 : kBtSmemMax));
    const int grid = (int)std::min<long long>(a.n_tiles, (long long)num_sms() * fb.ctas_per_sm);
    static long long *dbg_dev = nullptr;
    const char *tl = getenv("IMK_BT_TIMELINE");
    a.dbg = nullptr;
    a.dbg_skip = 0;
    if (const char *v = getenv("IMK_BT_TL_SKIP"); v && v[0]) a.dbg_skip = atoi(v);
    if (tl && tl[0] == '1') {
        if (!dbg_dev) IMK_CUDA(cudaMalloc(&dbg_dev, sizeof(long long) * 6 * 16 * 8));
        IMK_CUDA(cudaMemsetAsync(dbg_dev, 0, sizeof(long long) * 6 * 16 * 8, stream));
        a.dbg = dbg_dev;
    }
    kern<<<grid, fb.ctas_per_sm == 2 ? bt_threads(8, 4) : bt_threads(16, 8), fb.smem, stream>>>(a);
    IMK_LAUNCHED();
#ifdef IMK_BT_ACC_BUILD
    if (a.dbg) {
        long long h[8];
        IMK_CUDA(cudaStreamSynchronize(stream));
        IMK_CUDA(cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost));
        const double nt = (double)(h[5] > 0 ? h[5] : 1);
        fprintf(stderr, "[imk] loader phases kind=%d %dx%d tile %dx%d, CTA 0 warp 0, %lld tiles, cycles per tile: wait ld_empty %.0f | load + transform %.0f | fence + arrive %.0f | pool (incl. wait e3 %.0f) %.0f | DEC: wait for the TMA boxes %.0f\n",
                a.load_kind, a.H, a.W, a.Th, a.Tw, h[5], h[0] / nt, h[1] / nt, h[2] / nt, h[4] / nt, h[3] / nt, h[6] / nt);
        return IMK_OK;
    }
#endif
    if (a.dbg) {
        long long h[6 * 16 * 8];
        IMK_CUDA(cudaStreamSynchronize(stream));
        IMK_CUDA(cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost));
        long long t0 = 0;
        for (long long v : h) if (v && (!t0 || v < t0)) t0 = v;
        fprintf(stderr, "[imk] timeline kind=%d %dx%d tile %dx%d (cycles since first event; rows: tile, cols: events)\n", a.load_kind, a.H, a.W, a.Th, a.Tw);
        const char *names[6] = {"epi ", "mma ", "load", "s1is", "s3is", "epi5"};
        for (int r = 0; r < 6; ++r)
            for (int i = 0; i < 8; ++i) {
                fprintf(stderr, "[imk]   %s t%d:", names[r], i);
                for (int e = 0; e < 8; ++e) fprintf(stderr, " %8lld", h[(r * 16 + i) * 8 + e] ? h[(r * 16 + i) * 8 + e] - t0 : -1);
                fprintf(stderr, "\n");
            }
    }
    return IMK_OK;
}

}  // namespace imk

// tcgen05 implicit-GEMM convolution engine for the hidden U-Net layers (unet.py:11-43).
//
// One CTA owns a strip of Th output rows of one image.  The input strip (+1 halo row/col
// for 3x3, zero outside the image) is staged ONCE in shared memory as fp16 in the
// canonical K-major "core matrix" layout of the UMMA shared-memory descriptor with no
// swizzle:  plane kc (8 channels = 16 bytes) x flat padded pixel index f, 16 bytes each
//
//        addr(kc, f) = act + (kc * Pn + f) * 16          f = row * pitch + col
//
// so that 8 consecutive flat pixels form one 8x16B core matrix (SBO = 128 B to the next
// 8 pixels, LBO = Pn*16 B to the next 8 channels).  Because the layout is flat over the
// PADDED strip, the A operand of tap (dy,dx) is the same buffer shifted by
// (dy*pitch+dx)*16 bytes: no im2col copy exists anywhere, the 9 taps are 9 descriptor
// offsets.  The GEMM M index is the flat padded position, so the 2 halo columns of each
// row produce throw-away accumulator rows (2/(W+2) of the MMA work, free: the layers are
// HBM-bound) and every 128 consecutive flat positions are one UMMA M=128 block.
//   D[128 x Cout] (fp32, TMEM)  +=  A[128 x 16] (smem)  *  B[16 x Cout] (smem)    per (tap, 16 channels)
// Weights are pre-packed on the host in exactly the operand-B image ([tap][kc][Cout][8]
// fp16) and streamed through a 3-slot ring with 1-D bulk copies (cp.async.bulk +
// mbarrier complete_tx); all M blocks of the strip keep their accumulators in TMEM
// (n_mblocks * Cout <= 512 columns) so every weight byte is read once per CTA.
// Epilogue: tcgen05.ld (32 lanes x 16 columns) -> + bias -> ReLU -> BN affine -> fp16 ->
// 2 x 128-bit stores per 16 channels.  Optional prologue: nearest-upsample-2x + add
// (unet.py:32-33) fused into the strip load.
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include "imk_unet.cuh"

namespace imk {

constexpr int kTcThreads = 160;          // warps 0-3: strip load + epilogue (TMEM lanes 32w..32w+31); warp 4: + bulk copies, MMA issue
constexpr int kWSlots = 6;
constexpr int kWStageMax = 16 * 1024;
constexpr int kMaxMBlocks = 32;          // 512 TMEM columns / 16 output channels

struct TcArgs {
    const __half *in, *in_lo;
    __half *out;
    const __half *wpk;
    const float *bias, *bn_scale, *bn_shift;
    int H, W, cin_p, cout_p, taps, halo, pitch, Th, n_mblocks, Pn, tmem_cols;
    int units_total, units_per_stage, unit_bytes, ksteps, n_stages;
    int act_bytes, stage_bytes;
    int out_cstride, out_coff;              // channels of the whole output map / first channel of this launch (N split of wide layers)
    int use_tma;                            // strip load: one TMA box per 8-channel plane (plain layers, pitch <= 256)
    alignas(64) CUtensorMap tm_in;          // {8 ch, pitch, Th + 2 halo, 1} boxes of the fp16 NHWC input
    long long *dbg;                         // optional timeline of CTA (0,0) (IMK_TC_TIMELINE=1): 8 clock64 stamps
};

// ---- PTX wrappers ----------------------------------------------------------------------
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
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
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (layout_type 0), version 1 (Blackwell):
// bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c, int x, int y, int n, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c), "r"(x), "r"(y), "r"(n), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kTcThreads)
conv_tc_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *act = smem;
    uint8_t *wring = smem + a.act_bytes;
    float *par = reinterpret_cast<float *>(wring + kWSlots * a.stage_bytes);       // bias | bn_scale | bn_shift
    uint64_t *full = reinterpret_cast<uint64_t *>(par + 3 * a.cout_p);
    uint64_t *empty = full + kWSlots;
    uint64_t *acc_full = empty + kWSlots;                   // [kMaxMBlocks]: accumulators of M block b are final
    uint64_t *act_full = acc_full + kMaxMBlocks;            // the TMA boxes of the input strip have landed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(act_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int strip = blockIdx.x;
    const int64_t n = blockIdx.y;
    const int y0 = strip * a.Th;
    const int KC = a.cin_p >> 3;
#define TC_TL(ev) do { if (a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) a.dbg[ev] = clock64(); } while (0)
    if (warp == 0) TC_TL(0);

    // Measured (r01, HeLa level 3/4): streaming the ring through a 2/4/8-CTA cluster with multicast copies changes
    // nothing -- the MMA phase is bound by the operand reads from shared memory (A and B, 8 KB per N = 128 MMA), not by
    // the L2 -> SM weight stream -- so every CTA streams its own copy.
    if (warp == 4 && lane == 0) {
        for (int s = 0; s < kWSlots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < kMaxMBlocks; ++b) mbar_init(&acc_full[b], 1);
        mbar_init(act_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (a.use_tma) {
            // the whole haloed strip: one box per 8-channel plane, landing exactly as a plane of the flat layout, zero
            // filled outside the image -- no thread computes an address
            const int rows_ = a.Th + 2 * a.halo;
            mbar_expect_tx(act_full, (uint32_t)KC * 16u * (uint32_t)(rows_ * a.pitch));
            for (int kc = 0; kc < KC; ++kc)
                tma_load_4d(smem_u32(act) + (uint32_t)(kc * a.Pn) * 16u, &a.tm_in, kc * 8, -a.halo, y0 - a.halo, (int)n, act_full);
        }
        // weights do not depend on the strip: start streaming them right away
        const int pre = a.n_stages < kWSlots ? a.n_stages : kWSlots;
        for (int s = 0; s < pre; ++s) {
            mbar_expect_tx(&full[s], (uint32_t)a.stage_bytes);
            bulk_g2s(smem_u32(wring + s * a.stage_bytes), reinterpret_cast<const uint8_t *>(a.wpk) + (size_t)s * a.stage_bytes,
                     (uint32_t)a.stage_bytes, &full[s]);
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < a.cout_p; i += kTcThreads) {
        par[i] = a.bias[i];
        par[a.cout_p + i] = a.bn_scale ? a.bn_scale[i] : 1.f;
        par[2 * a.cout_p + i] = a.bn_scale ? a.bn_shift[i] : 0.f;
    }

    // ---- stage the input strip: global NHWC fp16 -> [kc][flat padded pixel][8 ch] ------------
    {
        const int rows = a.Th + 2 * a.halo;
        const int items = rows * a.pitch * KC;                 // 16-byte items, kc fastest (coalesced global reads)
        const unsigned kc_magic = 0xFFFFFFFFu / (unsigned)KC + 1u, pitch_magic = 0xFFFFFFFFu / (unsigned)a.pitch + 1u;   // exact for i < 2^20
        const __half *in_n = a.in + n * (int64_t)a.H * a.W * a.cin_p;
        const __half *lo_n = a.in_lo ? a.in_lo + n * (int64_t)(a.H >> 1) * (a.W >> 1) * a.cin_p : nullptr;
        if (a.use_tma) {
            __syncthreads();                                   // act_full is initialised
            mbar_wait(act_full, 0);
        } else if (!lo_n) {
            // plain layers: 16-byte cp.async straight into the operand layout (zero-fill outside the
            // image), every item of the strip in flight at once, no register staging
            const uint32_t act_u32 = smem_u32(act);
            for (int i = tid; i < items; i += kTcThreads) {
                const int f = (int)__umulhi((unsigned)i, kc_magic), kc = i - f * KC;
                const int r = (int)__umulhi((unsigned)f, pitch_magic), c = f - r * a.pitch;
                const int y = y0 + r - a.halo, x = c - a.halo;
                const bool inside = y >= 0 && y < a.H && x >= 0 && x < a.W;
                const __half *src = inside ? in_n + ((int64_t)y * a.W + x) * a.cin_p + kc * 8 : in_n;
                const uint32_t dst = act_u32 + (uint32_t)(kc * a.Pn + f) * 16u;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(inside ? 16 : 0) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else {
            // decoder entry: nearest-upsample-2x + add (unet.py:32-33) fused into the load: fp32 add, one rounding
            for (int i0 = tid; i0 < items; i0 += 4 * kTcThreads) {
                uint4 v[4], u[4];
                int dst[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = i0 + q * kTcThreads;
                    v[q] = make_uint4(0, 0, 0, 0);
                    u[q] = make_uint4(0, 0, 0, 0);
                    dst[q] = -1;
                    if (i < items) {
                        const int f = (int)__umulhi((unsigned)i, kc_magic), kc = i - f * KC;
                        const int r = (int)__umulhi((unsigned)f, pitch_magic), c = f - r * a.pitch;
                        const int y = y0 + r - a.halo, x = c - a.halo;
                        dst[q] = kc * a.Pn + f;
                        if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
                            v[q] = __ldg(reinterpret_cast<const uint4 *>(in_n + ((int64_t)y * a.W + x) * a.cin_p + kc * 8));
                            u[q] = __ldg(reinterpret_cast<const uint4 *>(lo_n + ((int64_t)(y >> 1) * (a.W >> 1) + (x >> 1)) * a.cin_p + kc * 8));
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (dst[q] < 0) continue;
                    const __half2 *pa = reinterpret_cast<const __half2 *>(&v[q]);
                    const __half2 *pb = reinterpret_cast<const __half2 *>(&u[q]);
                    uint4 r4;
                    __half2 *ro = reinterpret_cast<__half2 *>(&r4);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 fa = __half22float2(pa[e]), fb = __half22float2(pb[e]);
                        ro[e] = __floats2half2_rn(__fadd_rn(fa.x, fb.x), __fadd_rn(fa.y, fb.y));
                    }
                    *reinterpret_cast<uint4 *>(act + (size_t)dst[q] * 16) = r4;
                }
            }
        }
    }
    if (warp == 0) TC_TL(1);                    // own share of the strip has landed
    // generic-proxy writes to smem -> visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 0) TC_TL(2);                    // every warp's share has landed

    if (warp == 4) {
        if (lane == 0) {
            // instruction descriptor: D fp32 (bit 4), A/B fp16 K-major, N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(a.cout_p >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t act_base = smem_u32(act);
            // K-major, no swizzle: LBO = distance between the two 8-channel core matrices of one K=16 step,
            // SBO = distance between consecutive 8-row (8 pixels / 8 couts) core matrices
            const uint32_t a_lbo = (uint32_t)a.Pn * 16u, a_sbo = 128u;
            const uint32_t b_lbo = (uint32_t)a.cout_p * 16u, b_sbo = 128u;
            // The issue sequence of one MMA must stay far below the MMA itself (64 cycles at N = 128): descriptors are
            // two 32-bit words each, the high words are constants, the low words advance by additions only -- the
            // (tap, k-step) position of a unit is carried incrementally, no division on this path.
            //   lo = (addr >> 4) | (LBO >> 4) << 16      hi = (SBO >> 4) | version 1 << 14
            (void)a_sbo; (void)b_sbo;
            constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
            const uint32_t a_lo0 = (act_base >> 4) | ((a_lbo >> 4) << 16);
            const uint32_t b_lbo16 = (b_lbo >> 4) << 16, unit16 = (uint32_t)a.unit_bytes >> 4;
            const uint32_t jstep = 2u * (uint32_t)a.Pn, pitch = (uint32_t)a.pitch, ncol = (uint32_t)a.cout_p;
            const int ksteps = a.ksteps, nmb = a.n_mblocks;
            auto mma = [&](uint32_t d, uint32_t alo, uint32_t blo, uint32_t acc) {
                tc_mma_f16(d, ((uint64_t)kDescHi << 32) | alo, ((uint64_t)kDescHi << 32) | blo, idesc, acc);
            };
            struct Pos { int j, dx; uint32_t off; uint32_t row; };        // off = 2*j*Pn + dy*pitch + dx (16-byte units), row = dy*pitch
            auto advance = [&](Pos &p) {
                if (++p.j < ksteps) { p.off += jstep; return; }
                p.j = 0;
                if (a.taps == 9 && ++p.dx == 3) { p.dx = 0; p.row += pitch; }
                p.off = p.row + (uint32_t)p.dx;
            };
            if (a.n_stages == 1) {
                // all weights resident: finish one M block at a time so its epilogue overlaps the next block's MMAs
                mbar_wait(&full[0], 0);
                tc_fence_after();
                const uint32_t b_lo0 = (smem_u32(wring) >> 4) | b_lbo16;
                for (int b = 0; b < nmb; ++b) {
                    Pos p{0, 0, 0u, 0u};
                    uint32_t blo = b_lo0;
                    const uint32_t d = tmem + (uint32_t)b * ncol, ab = a_lo0 + (uint32_t)b * 128u;
                    for (int u = 0; u < a.units_total; ++u, blo += unit16) { mma(d, ab + p.off, blo, u > 0 ? 1u : 0u); advance(p); }
                    tc_commit(&acc_full[b]);
                }
            } else {
                Pos p{0, 0, 0u, 0u};
                uint32_t acc = 0;
                int slot = 0;
                uint32_t ph = 0;
                for (int s = 0; s < a.n_stages; ++s) {
                    mbar_wait(&full[slot], ph);
                    tc_fence_after();
                    uint32_t blo = (smem_u32(wring + slot * a.stage_bytes) >> 4) | b_lbo16;
                    for (int ul = 0; ul < a.units_per_stage; ++ul, blo += unit16) {
                        const uint32_t ab = a_lo0 + p.off;
                        uint32_t d = tmem;
                        for (int b = 0; b < nmb; ++b, d += ncol) mma(d, ab + (uint32_t)b * 128u, blo, acc);
                        acc = 1;
                        advance(p);
                    }
                    tc_commit(&empty[slot]);                           // arrives when the MMAs above have read this slot
                    if (s >= 1 && s - 1 + kWSlots < a.n_stages) {      // refill the slot stage s-1 used
                        const int ps = s - 1, pslot = slot == 0 ? kWSlots - 1 : slot - 1, ns = ps + kWSlots;
                        const uint32_t pph = slot == 0 ? ph ^ 1u : ph;                // parity of stage ps on its slot
                        mbar_wait(&empty[pslot], pph);
                        mbar_expect_tx(&full[pslot], (uint32_t)a.stage_bytes);
                        bulk_g2s(smem_u32(wring + pslot * a.stage_bytes),
                                 reinterpret_cast<const uint8_t *>(a.wpk) + (size_t)ns * a.stage_bytes, (uint32_t)a.stage_bytes, &full[pslot]);
                    }
                    if (++slot == kWSlots) { slot = 0; ph ^= 1u; }
                }
                for (int b = 0; b < nmb; ++b) tc_commit(&acc_full[b]);
            }
            TC_TL(3);                           // last MMA issued
        }
        __syncwarp();
    } else {
        // ---- epilogue: TMEM -> registers -> bias, ReLU, BN -> fp16 -> global ------------------
        __half *out_n = a.out + n * (int64_t)a.H * a.W * a.out_cstride + a.out_coff;
        for (int b = 0; b < a.n_mblocks; ++b) {
            mbar_wait(&acc_full[b], 0);
            if (warp == 0 && b == 0) TC_TL(4);  // first accumulator block final
            tc_fence_after();
            const int m = b * 128 + warp * 32 + lane;
            const int ro = m / a.pitch, co = m - ro * a.pitch;
            const int y = y0 + ro;
            const bool valid = co < a.W && ro < a.Th && y < a.H;
            __half *dst = out_n + ((int64_t)y * a.W + co) * a.out_cstride;
            for (int c0 = 0; c0 < a.cout_p; c0 += 16) {
                uint32_t r[16];
                tc_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * a.cout_p + c0), r);
                tc_wait_ld();
                if (valid) {
                    __align__(16) __half o[16];
#pragma unroll
                    for (int e4 = 0; e4 < 4; ++e4) {
                        const float4 pb = *reinterpret_cast<const float4 *>(par + c0 + 4 * e4);
                        const float4 ps = *reinterpret_cast<const float4 *>(par + a.cout_p + c0 + 4 * e4);
                        const float4 ph = *reinterpret_cast<const float4 *>(par + 2 * a.cout_p + c0 + 4 * e4);
                        const float b4[4] = {pb.x, pb.y, pb.z, pb.w}, s4[4] = {ps.x, ps.y, ps.z, ps.w}, h4[4] = {ph.x, ph.y, ph.z, ph.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float v = fmaxf(__uint_as_float(r[4 * e4 + k]) + b4[k], 0.f);
                            v = __fmaf_rn(v, s4[k], h4[k]);
                            o[4 * e4 + k] = __float2half_rn(v);
                        }
                    }
                    reinterpret_cast<uint4 *>(dst + c0)[0] = reinterpret_cast<const uint4 *>(o)[0];
                    reinterpret_cast<uint4 *>(dst + c0)[1] = reinterpret_cast<const uint4 *>(o)[1];
                }
            }
        }
        tc_fence_before();
        if (warp == 0) TC_TL(5);                // epilogue done
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(a.tmem_cols) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------
static int env_flag(const char *name) {
    const char *v = getenv(name);
    return (v && v[0] && v[0] != '0') ? 1 : 0;
}

// Layers wider than the UMMA N limit (256) run as two launches over halves of the output channels (the bottleneck
// 3x3 of the alpha = 2 networks, 256 -> 512): each half has its own operand-B image, bias / BN slice and channel offset.
static int tc_parts(const ConvLayer &L) { return L.cout_p <= 256 ? 1 : 2; }

bool conv_tc_supported(const ConvLayer &L) {
    if (env_flag("IMK_TC_DISABLE")) return false;
    return (L.ks == 1 || L.ks == 3) && L.cin_p % 16 == 0 && L.cout_p % 16 == 0 && L.cout_p <= 512 && (L.cout_p / tc_parts(L)) % 16 == 0 &&
           L.cin_p <= 1024;
}

// operand-B image per N part: [tap][kc = cin_p/8][cout_p / parts][8] fp16 (zero in the padding), parts back to back
int conv_tc_pack(ConvLayer &L, const float *hwio, std::vector<void *> &owned) {
    const int taps = L.ks * L.ks, KC = L.cin_p / 8, parts = tc_parts(L), pn = L.cout_p / parts;
    std::vector<__half> w((size_t)taps * KC * L.cout_p * 8, __float2half(0.f));
    for (int tap = 0; tap < taps; ++tap)
        for (int ci = 0; ci < L.cin; ++ci)
            for (int co = 0; co < L.cout; ++co)
                w[(size_t)(co / pn) * taps * KC * pn * 8 + (((size_t)tap * KC + ci / 8) * pn + co % pn) * 8 + (ci & 7)] =
                    __float2half_rn(hwio[((size_t)tap * L.cin + ci) * L.cout + co]);
    void *p = nullptr;
    if (cudaMalloc(&p, w.size() * sizeof(__half)) != cudaSuccess) { set_error("conv_tc_pack: cudaMalloc failed"); return IMK_ENOMEM; }
    owned.push_back(p);
    IMK_CUDA(cudaMemcpy(p, w.data(), w.size() * sizeof(__half), cudaMemcpyHostToDevice));
    L.w_umma = reinterpret_cast<__half *>(p);
    return IMK_OK;
}

static int next_pow2_cols(int c) {
    int p = 32;
    while (p < c) p <<= 1;
    return p;
}

// Chooses the strip height, the weight staging and the residency (1 or 2 CTAs per SM) for a layer with a small
// cost model (cycles): a CTA loads its strip, streams the weights once while it issues units x M-blocks MMAs, and
// drains its accumulators; a ring stage whose MMAs take less than the L2 round trip stalls on the weight stream;
// two co-resident CTAs only help against wave quantisation.  Returns false if nothing fits.
static bool tc_plan(const ConvLayer &L, int h, int w, int64_t n_images, bool tma, TcArgs &a) {
    a.H = h; a.W = w; a.cin_p = L.cin_p; a.cout_p = L.cout_p;
    a.taps = L.ks * L.ks; a.halo = L.ks / 2; a.pitch = w + 2 * a.halo;
    a.ksteps = L.cin_p / 16;
    a.units_total = a.taps * a.ksteps;
    a.unit_bytes = L.cout_p * 32;
    const int KC = L.cin_p / 8;
    // measured in situ (IMK_TC_TIMELINE, r01): the tap-shifted A operand is misaligned to the 128-byte shared-memory
    // lines and B is re-read for every M block, so an MMA costs its operand reads, about 64 + 0.8 N cycles
    const double mma_cyc = 64.0 + 0.8 * L.cout_p;
    const double l2_trip = 1500.0;
    double best = -1.0;
    int force_th = 0;                                            // tuning aid: IMK_TC_TH pins the strip height of 3x3 layers
    if (const char *v = getenv("IMK_TC_TH"); v && v[0] && L.ks == 3) force_th = atoi(v);
    for (int stage_max = kWStageMax; stage_max >= 8 * 1024; stage_max >>= 1) {
        int U = 1;
        for (int d = 1; d <= a.units_total; ++d)
            if (a.units_total % d == 0 && d * a.unit_bytes <= stage_max) U = d;
        const int stage_bytes = U * a.unit_bytes, n_stages = a.units_total / U;
        const int fixed = kWSlots * stage_bytes + 3 * L.cout_p * 4 + 128;
        for (int pass = 0; pass < 2; ++pass) {
            const int max_cols = pass == 0 ? 256 : 512;
            const int max_smem = pass == 0 ? 110 * 1024 : 220 * 1024;
            const int resident = pass == 0 ? 2 : 1;
            for (int th = std::min(h, 32); th >= 1; --th) {
                if (force_th > 0 && th != force_th) continue;
                const int nmb = (th * a.pitch + 127) / 128;
                if (nmb > kMaxMBlocks || nmb * L.cout_p > max_cols) continue;
                // plane stride in positions: a multiple of 8 (128-byte aligned TMA destinations) or odd (conflict-free
                // 16-byte writes of the thread-staged loaders)
                const int pn0 = nmb * 128 + 2 * a.halo * a.pitch + 2 * a.halo;
                const int pn = tma ? (std::max(pn0, (th + 2 * a.halo) * a.pitch) + 7) / 8 * 8 : (pn0 | 1);
                const int act_bytes = (KC * pn * 16 + 127) / 128 * 128;
                if (act_bytes + fixed > max_smem) continue;
                const double ctas = (double)((h + th - 1) / th) * (double)std::max<int64_t>(n_images, 1);
                const double t_mma = a.units_total * nmb * mma_cyc;
                const double t_stream = n_stages > 1 ? n_stages * std::max(U * nmb * mma_cyc, l2_trip / (kWSlots - 1)) : t_mma;
                // CTA start-up (barriers, TMEM, first weight stage in flight: ~3 us measured on small strips) + strip load +
                // epilogue, none of them overlapped inside a CTA
                const double t_io = 6000.0 + act_bytes / 12.0 + nmb * (L.cout_p / 16) * 400.0;
                const double t_cta = t_io + std::max(t_mma, t_stream);
                // measured (IMK_TC_TH sweep, r01): two co-resident CTAs do not overlap -- every phase of this kernel is bound
                // by the SM's shared-memory bandwidth -- so a pair costs a little more than two CTAs back to back
                // (3x3 layers on maps of at least 32x32; for the 1x1 layers and the 16x16 bottleneck the pair does overlap)
                const bool no_overlap = a.taps == 9 && h * w >= 1024;
                const double wave = resident == 2 ? (no_overlap ? 2.2 * t_cta : std::max(2.0 * t_mma, t_cta)) : t_cta;
                const double waves = std::ceil(ctas / (double)(kNumSMs * resident));
                const double cost = waves * wave;
                if (best < 0 || cost < best) {
                    best = cost;
                    a.units_per_stage = U; a.stage_bytes = stage_bytes; a.n_stages = n_stages;
                    a.Th = th; a.n_mblocks = nmb; a.Pn = pn; a.act_bytes = act_bytes;
                    a.tmem_cols = next_pow2_cols(nmb * L.cout_p);
                }
            }
        }
    }
    return best >= 0;
}

bool conv_tc_fits(const ConvLayer &L, int h, int w) {
    TcArgs a{};
    if (!conv_tc_supported(L) || !L.w_umma) return false;
    ConvLayer Lp = L;
    Lp.cout_p = L.cout_p / tc_parts(L);
    return tc_plan(Lp, h, w, 64, false, a);
}

int conv_tc_launch(const ConvLayer &L, const __half *in, const __half *in_lo, __half *out, __half *pool_out,
                   int64_t n, int h, int w, cudaStream_t stream) {
    (void)pool_out;
    const int parts = tc_parts(L);
    ConvLayer Lp = L;
    Lp.cout_p = L.cout_p / parts;
    TcArgs a{};
    bool tma = !in_lo && w + 2 * (L.ks / 2) <= 256 && !env_flag("IMK_TC_NO_TMA");
    if (tma && !tc_plan(Lp, h, w, n, true, a)) tma = false;          // the aligned plane stride did not fit: thread-staged load
    if (!tma && !tc_plan(Lp, h, w, n, false, a)) { set_error("conv_tc_launch: no strip configuration fits (cin_p=%d cout_p=%d %dx%d)", L.cin_p, L.cout_p, h, w); return IMK_ESTATE; }
    a.in = in; a.in_lo = in_lo; a.out = out;
    a.out_cstride = L.cout_p;
    a.use_tma = tma ? 1 : 0;
    if (tma) {
        int rc = make_map(&a.tm_in, in, n, h, w, L.cin_p, a.pitch, a.Th + 2 * a.halo);
        if (rc) return rc;
    }
    const size_t smem = (size_t)a.act_bytes + kWSlots * a.stage_bytes + 3 * a.cout_p * 4 + (2 * kWSlots + kMaxMBlocks + 1) * 8 + 16;
    {   // the opt-in is per device (a process may drive several GPUs): once per device
        static bool attr_set[64] = {false};
        int dev = 0;
        IMK_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            IMK_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    if (env_flag("IMK_TC_VERBOSE"))
        fprintf(stderr, "[imk] conv_tc %dx%d k%d %d->%d: strip %d rows, %d M blocks, TMEM %d cols, %d stages of %d B, act %d B\n", h, w, L.ks,
                L.cin_p, L.cout_p, a.Th, a.n_mblocks, a.tmem_cols, a.n_stages, a.stage_bytes, a.act_bytes);
    dim3 grid((h + a.Th - 1) / a.Th, (unsigned)n);
    static long long *dbg_dev = nullptr;
    for (int part = 0; part < parts; ++part) {
        a.out_coff = part * Lp.cout_p;
        a.wpk = L.w_umma + (size_t)part * L.ks * L.ks * L.cin_p * Lp.cout_p;
        a.bias = L.bias + a.out_coff;
        a.bn_scale = L.has_bn ? L.bn_scale + a.out_coff : nullptr;
        a.bn_shift = L.has_bn ? L.bn_shift + a.out_coff : nullptr;
        a.dbg = nullptr;
        if (env_flag("IMK_TC_TIMELINE")) {
            if (!dbg_dev) IMK_CUDA(cudaMalloc(&dbg_dev, 8 * sizeof(long long)));
            IMK_CUDA(cudaMemsetAsync(dbg_dev, 0, 8 * sizeof(long long), stream));
            a.dbg = dbg_dev;
        }
        conv_tc_kernel<<<grid, kTcThreads, smem, stream>>>(a);
        IMK_LAUNCHED();
        if (a.dbg) {
            long long t[8];
            IMK_CUDA(cudaStreamSynchronize(stream));
            IMK_CUDA(cudaMemcpy(t, dbg_dev, sizeof(t), cudaMemcpyDeviceToHost));
            fprintf(stderr, "[imk] conv_tc timeline %dx%d k%d %d->%d (cycles): strip own %lld, all %lld, last MMA issued %lld, first block final %lld, epilogue done %lld\n",
                    h, w, L.ks, L.cin_p, Lp.cout_p, t[1] - t[0], t[2] - t[0], t[3] - t[0], t[4] - t[0], t[5] - t[0]);
        }
    }
    return IMK_OK;
}

}  // namespace imk

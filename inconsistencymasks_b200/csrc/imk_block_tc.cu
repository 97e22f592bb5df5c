// Block-fused tcgen05 engine: one persistent, warp-specialised kernel runs a whole U-Net
// block (unet.py:4-43) per tile without the intermediate maps ever leaving the SM:
//
//   FRONT (level 0)  uint8 image -> [x/255 -> 1x1 conv + ReLU + BN]  -> [3x3 conv + ReLU] -> [1x1 conv + ReLU + BN] -> skip
//   ENC   (level>=1) pooled map  ->                                     [3x3 conv + ReLU] -> [1x1 conv + ReLU + BN] -> skip
//   DEC              up2x(lo) + skip -> [1x1 conv + ReLU + BN]       -> [3x3 conv + ReLU] -> [1x1 conv + ReLU + BN] -> map
//
// i.e. a chain of up to three GEMM stages S1 (1x1 over the haloed tile), S2 (3x3), S3 (1x1).
// A tile is Th x Tw output pixels of one image.  All operands live in shared memory in the
// canonical no-swizzle K-major core-matrix layout over the FLAT padded tile index
// f = row * pitch + col (pitch = Tw + 2):  addr(kc, f) = base + (kc * Pn + f) * 16, so the
// nine taps of the 3x3 stage are nine descriptor offsets into the same buffer (no im2col) and
// the output of one stage's epilogue IS the A operand of the next stage.  Accumulators live in
// TMEM (three regions R1/R2/R3, one per stage); the weights of all stages stay resident in
// shared memory for the life of the CTA.
//
// Roles (416 threads, 1 CTA / SM, grid = min(tiles, 148), tiles strided over CTAs):
//   warps 0-7   epilogue: tcgen05.ld -> +bias, ReLU, BN -> fp16 -> next stage's smem operand / global
//               (warp w owns TMEM lanes 32*(w%4).., M blocks b = w/4 (mod 2))
//   warp  8     TMEM allocation + single-thread tcgen05.mma issue
//   warps 9-12  loaders: global -> smem operand of the first stage (uint8 image with hi/lo fp16
//               split, plain cp.async copy, or nearest-upsample-2x + add)
// The stages of consecutive tiles are software-pipelined so that the epilogue warps (the
// instruction-issue bottleneck) never wait for the tensor pipe:
//   epilogue iteration i :  E1(i)            E3(i-1)          E2(i)
//   MMA      iteration i :  S2(i)  S1(i+1)                    S3(i, block by block behind E2)
//   loader               :  tile i+1 as soon as S1(i) / S2(i) has consumed the buffer
// All hand-offs are mbarriers (tcgen05.commit on the MMA side), every barrier completes exactly
// once per tile so the wait parity is the tile parity.
#include <algorithm>
#include <stdlib.h>
#include "imk_unet.cuh"

namespace imk {

constexpr int kBtEpiWarps = 8;
constexpr int kBtLoadWarps = 4;
constexpr int kBtThreads = (kBtEpiWarps + 1 + kBtLoadWarps) * 32;
constexpr int kBtSmemMax = 227 * 1024;

namespace {

// ---- PTX wrappers ----------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
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
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, no swizzle, version 1 (see imk_conv_tc.cu)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void tile_coords(const BtArgs &a, long long tile, long long &n, int &y0, int &x0) {
    const int per_img = a.tiles_x * a.tiles_y;
    n = tile / per_img;
    const int r = (int)(tile - n * per_img);
    const int ty = r / a.tiles_x;
    y0 = ty * a.Th;
    x0 = (r - ty * a.tiles_x) * a.Tw;
}

// bias + ReLU + BN of 16 accumulator columns -> 8 packed half2 words
__device__ __forceinline__ void epi16(const uint32_t (&r)[16], const float *__restrict__ p /*bias[n]|scale[n]|shift[n] at chunk*/,
                                      int n, bool keep, uint4 &lo, uint4 &hi) {
    uint32_t o[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b = *reinterpret_cast<const float4 *>(p + 4 * q);
        const float4 s = *reinterpret_cast<const float4 *>(p + n + 4 * q);
        const float4 t = *reinterpret_cast<const float4 *>(p + 2 * n + 4 * q);
        float v0 = __fmaf_rn(fmaxf(__uint_as_float(r[4 * q + 0]) + b.x, 0.f), s.x, t.x);
        float v1 = __fmaf_rn(fmaxf(__uint_as_float(r[4 * q + 1]) + b.y, 0.f), s.y, t.y);
        float v2 = __fmaf_rn(fmaxf(__uint_as_float(r[4 * q + 2]) + b.z, 0.f), s.z, t.z);
        float v3 = __fmaf_rn(fmaxf(__uint_as_float(r[4 * q + 3]) + b.w, 0.f), s.w, t.w);
        if (!keep) { v0 = 0.f; v1 = 0.f; v2 = 0.f; v3 = 0.f; }
        const __half2 h0 = __floats2half2_rn(v0, v1), h1 = __floats2half2_rn(v2, v3);
        o[2 * q] = *reinterpret_cast<const uint32_t *>(&h0);
        o[2 * q + 1] = *reinterpret_cast<const uint32_t *>(&h1);
    }
    lo = make_uint4(o[0], o[1], o[2], o[3]);
    hi = make_uint4(o[4], o[5], o[6], o[7]);
}

}  // namespace

__global__ void __launch_bounds__(kBtThreads, 1)
block_tc_kernel(const BtArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    float *par = reinterpret_cast<float *>(smem + a.par_off_b);
    uint8_t *A0 = smem + a.a0_off, *A1 = smem + a.a1_off, *A2 = smem + a.a2_off;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + a.bar_off);
    uint64_t *ld_full = bars, *ld_empty = bars + 1, *e1_done = bars + 2, *e3_done = bars + 3;
    uint64_t *acc1_full = bars + 4, *acc2_full = acc1_full + kBtMaxBlocks, *acc3_full = acc2_full + kBtMaxBlocks;
    uint64_t *e2_done = acc3_full + kBtMaxBlocks;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(e2_done + kBtMaxBlocks);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n_my = (a.n_tiles - (long long)blockIdx.x + gridDim.x - 1) / gridDim.x;

    // ---- one-time setup ---------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(ld_full, kBtLoadWarps); mbar_init(ld_empty, 1);
        mbar_init(e1_done, kBtEpiWarps); mbar_init(e3_done, kBtEpiWarps);
        for (int b = 0; b < kBtMaxBlocks; ++b) {
            mbar_init(&acc1_full[b], 1); mbar_init(&acc2_full[b], 1); mbar_init(&acc3_full[b], 1);
            mbar_init(&e2_done[b], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kBtEpiWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // resident weights + parameters; operand buffers start zeroed (positions no loader / epilogue writes)
        const uint4 *src = reinterpret_cast<const uint4 *>(a.wpk);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = tid; i < a.w_bytes / 16; i += kBtThreads) dst[i] = src[i];
        for (int i = tid; i < a.par_floats; i += kBtThreads) par[i] = a.par[i];
        uint4 *z = reinterpret_cast<uint4 *>(smem + a.a0_off);
        const int zn = (a.bar_off - a.a0_off) / 16;
        for (int i = tid; i < zn; i += kBtThreads) z[i] = make_uint4(0, 0, 0, 0);
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
        for (long long i = 0; i <= n_my; ++i) {
            long long n = 0; int y0 = 0, x0 = 0;
            // ---- E1(i): S1 accumulators -> ReLU + BN, zero outside the image -> A1 (haloed flat layout)
            if (a.has_s1 && i < n_my) {
                tile_coords(a, (long long)blockIdx.x + i * gridDim.x, n, y0, x0);
                const uint32_t par_ = (uint32_t)(i & 1);
                for (int b = g; b < a.s1.nb; b += 2) {
                    mbar_wait(&acc1_full[b], par_);
                    __syncwarp();
                    tc_fence_after();
                    const int m = b * 128 + q * 32 + lane;
                    const int r = (int)__umulhi((unsigned)m, a.pitch_magic), c = m - r * a.pitch;
                    const int y = y0 - 1 + r, x = x0 - 1 + c;
                    const bool inside = r < a.Th + 2 && y >= 0 && y < a.H && x >= 0 && x < a.W;
                    for (int c0 = 0; c0 < a.s1.n; c0 += 16) {
                        uint32_t rr[16];
                        tc_ld16(tmem + lane_base + (uint32_t)(a.s1.col + b * a.s1.n + c0), rr);
                        tc_wait_ld();
                        uint4 lo, hi;
                        epi16(rr, par + a.s1.par_off + c0, a.s1.n, inside, lo, hi);
                        *reinterpret_cast<uint4 *>(A1 + ((size_t)(c0 >> 3) * a.Pn1 + m) * 16) = lo;
                        *reinterpret_cast<uint4 *>(A1 + ((size_t)((c0 >> 3) + 1) * a.Pn1 + m) * 16) = hi;
                    }
                }
                fence_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(e1_done);
            }
            // ---- E3(i-1): S3 accumulators -> ReLU + BN -> global
            if (i >= 1) {
                tile_coords(a, (long long)blockIdx.x + (i - 1) * gridDim.x, n, y0, x0);
                const uint32_t par_ = (uint32_t)((i - 1) & 1);
                __half *out_n = a.out + n * (long long)a.H * a.W * a.s3.n;
                for (int b = g; b < a.s3.nb; b += 2) {
                    mbar_wait(&acc3_full[b], par_);
                    __syncwarp();
                    tc_fence_after();
                    const int m = b * 128 + q * 32 + lane;
                    const int ro = (int)__umulhi((unsigned)m, a.pitch_magic), co = m - ro * a.pitch;
                    const int y = y0 + ro, x = x0 + co;
                    const bool valid = ro < a.Th && co < a.Tw && y < a.H && x < a.W;
                    __half *dst = out_n + ((long long)y * a.W + x) * a.s3.n;
                    for (int c0 = 0; c0 < a.s3.n; c0 += 16) {
                        uint32_t rr[16];
                        tc_ld16(tmem + lane_base + (uint32_t)(a.s3.col + b * a.s3.n + c0), rr);
                        tc_wait_ld();
                        uint4 lo, hi;
                        epi16(rr, par + a.s3.par_off + c0, a.s3.n, true, lo, hi);
                        if (valid) {
                            reinterpret_cast<uint4 *>(dst + c0)[0] = lo;
                            reinterpret_cast<uint4 *>(dst + c0)[1] = hi;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(e3_done);
            }
            // ---- E2(i): S2 accumulators -> ReLU -> A2 (flat layout, every row written)
            if (i < n_my) {
                const uint32_t par_ = (uint32_t)(i & 1);
                for (int b = g; b < a.s2.nb; b += 2) {
                    mbar_wait(&acc2_full[b], par_);
                    __syncwarp();
                    tc_fence_after();
                    const int m = b * 128 + q * 32 + lane;
                    for (int c0 = 0; c0 < a.s2.n; c0 += 16) {
                        uint32_t rr[16];
                        tc_ld16(tmem + lane_base + (uint32_t)(a.s2.col + b * a.s2.n + c0), rr);
                        tc_wait_ld();
                        uint4 lo, hi;
                        epi16(rr, par + a.s2.par_off + c0, a.s2.n, true, lo, hi);
                        *reinterpret_cast<uint4 *>(A2 + ((size_t)(c0 >> 3) * a.Pn2 + m) * 16) = lo;
                        *reinterpret_cast<uint4 *>(A2 + ((size_t)((c0 >> 3) + 1) * a.Pn2 + m) * 16) = hi;
                    }
                    fence_async_smem();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&e2_done[b]);
                }
            }
        }
    } else if (warp == kBtEpiWarps) {
        // =====================================================================================
        //  MMA issue (one thread)
        // =====================================================================================
        // Every lane runs the (warp-uniform) bookkeeping so that the descriptor arithmetic stays in the uniform
        // datapath; only the tcgen05 instructions themselves are issued by one elected lane.
        uint32_t elected;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(elected));
        const uint32_t wbase = smem_u32(smem);
        const uint32_t pitch = (uint32_t)a.pitch;
        // descriptor words: lo = addr >> 4 | LBO >> 4 << 16 ; hi = SBO >> 4 | version 1 << 14
        constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
        struct StageRegs { uint32_t a_lo, a_step, b_lo, b_unit, idesc, d, n, nb, ksteps; };
        auto make_stage = [&](const BtStage &st, uint32_t abase, int Pn) {
            StageRegs r;
            r.a_lo = (abase >> 4) | ((uint32_t)Pn << 16);            // LBO = Pn * 16 bytes
            r.a_step = 2u * (uint32_t)Pn;                             // next K step: two 8-channel planes further
            r.b_lo = ((wbase + (uint32_t)st.w_off) >> 4) | ((uint32_t)st.n << 16);   // LBO = n * 16 bytes
            r.b_unit = (uint32_t)st.n * 2u;                           // n * 32 bytes per (tap, K step)
            r.idesc = (1u << 4) | ((uint32_t)(st.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            r.d = tmem + (uint32_t)st.col;
            r.n = (uint32_t)st.n; r.nb = (uint32_t)st.nb; r.ksteps = (uint32_t)st.ksteps;
            return r;
        };
        const StageRegs S1 = make_stage(a.s1, smem_u32(A0), a.Pn0);
        const StageRegs S2 = make_stage(a.s2, smem_u32(A1), a.Pn1);
        const StageRegs S3 = make_stage(a.s3, smem_u32(A2), a.Pn2);
        auto mma = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
            if (elected)
                tc_mma_f16(d, ((uint64_t)kDescHi << 32) | a_lo, ((uint64_t)kDescHi << 32) | b_lo, idesc, acc);
        };
        auto commit = [&](uint64_t *bar) { if (elected) tc_commit(bar); };
        // one M block of a 1x1 stage: ksteps MMAs
        auto block_1x1 = [&](const StageRegs &r, uint32_t b) {
            const uint32_t d = r.d + b * r.n;
            uint32_t al = r.a_lo + b * 128u, bl = r.b_lo;
            for (uint32_t j = 0; j < r.ksteps; ++j, al += r.a_step, bl += r.b_unit) mma(d, al, bl, r.idesc, j);
        };
        // one M block of the 3x3 stage: 9 taps x ksteps MMAs, the taps are offsets dy * pitch + dx into the flat tile
        auto block_3x3 = [&](const StageRegs &r, uint32_t b) {
            const uint32_t d = r.d + b * r.n;
            uint32_t bl = r.b_lo, acc = 0;
            uint32_t arow = r.a_lo + b * 128u;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy, arow += pitch) {
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    uint32_t al = arow + (uint32_t)dx;
                    for (uint32_t j = 0; j < r.ksteps; ++j, al += r.a_step, bl += r.b_unit) { mma(d, al, bl, r.idesc, acc); acc = 1; }
                }
            }
        };
        if (a.has_s1 && n_my > 0) {
            mbar_wait(ld_full, 0);
            tc_fence_after();
            for (uint32_t b = 0; b < S1.nb; ++b) { block_1x1(S1, b); commit(&acc1_full[b]); }
            commit(ld_empty);
        }
        for (long long i = 0; i < n_my; ++i) {
            const uint32_t par_ = (uint32_t)(i & 1);
            // S2(i): its operand is A1 -- written by E1(i) (chain of three) or by the loaders (chain of two).
            // R2 is free: every e2_done[b] of tile i-1 was waited for below.
            mbar_wait(a.has_s1 ? e1_done : ld_full, par_);
            tc_fence_after();
            for (uint32_t b = 0; b < S2.nb; ++b) { block_3x3(S2, b); commit(&acc2_full[b]); }
            if (!a.has_s1) commit(ld_empty);
            // S1(i+1): R1 is free (e1_done(i) above)
            if (a.has_s1 && i + 1 < n_my) {
                mbar_wait(ld_full, (uint32_t)((i + 1) & 1));
                tc_fence_after();
                for (uint32_t b = 0; b < S1.nb; ++b) { block_1x1(S1, b); commit(&acc1_full[b]); }
                commit(ld_empty);
            }
            // S3(i), block by block behind E2(i); R3 is free once E3(i-1) has drained it
            if (i > 0) { mbar_wait(e3_done, (uint32_t)((i - 1) & 1)); }
            for (uint32_t b = 0; b < S3.nb; ++b) {
                mbar_wait(&e2_done[b], par_);
                tc_fence_after();
                block_1x1(S3, b);
                commit(&acc3_full[b]);
            }
        }
        __syncwarp();
    } else {
        // =====================================================================================
        //  loaders
        // =====================================================================================
        const int lt = tid - (kBtEpiWarps + 1) * 32;
        constexpr int NL = kBtLoadWarps * 32;
        uint8_t *buf = a.has_s1 ? A0 : A1;
        const int Pn = a.has_s1 ? a.Pn0 : a.Pn1;
        const int npos = (a.Th + 2) * a.pitch;
        const int KC = a.ld_cp >> 3;
        for (long long i = 0; i < n_my; ++i) {
            long long n; int y0, x0;
            tile_coords(a, (long long)blockIdx.x + i * gridDim.x, n, y0, x0);
            if (i > 0) mbar_wait(ld_empty, (uint32_t)((i - 1) & 1));
            if (a.load_kind == 0) {
                // uint8 image -> x/255 split into fp16 hi + lo so that the first layer keeps ~22 bits of the
                // input and of the weights: K slots [hi(c) | lo(c) | hi(c)] against [w_hi | w_hi | w_lo]
                const uint8_t *img = reinterpret_cast<const uint8_t *>(a.in) + n * (long long)a.H * a.W * a.in_c;
                const float *imgf = reinterpret_cast<const float *>(a.in) + n * (long long)a.H * a.W * a.in_c;
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
            } else if (a.load_kind == 1) {
                // plain copy of the haloed tile: 16-byte cp.async with zero fill outside the image
                const __half *in_n = reinterpret_cast<const __half *>(a.in) + n * (long long)a.H * a.W * a.ld_cp;
                const uint32_t dst0 = smem_u32(buf);
                const int items = npos * KC;
                for (int it = lt; it < items; it += NL) {
                    const int f = it / KC, kc = it - f * KC;
                    const int r = (int)__umulhi((unsigned)f, a.pitch_magic), c = f - r * a.pitch;
                    const int y = y0 - 1 + r, x = x0 - 1 + c;
                    const bool inside = y >= 0 && y < a.H && x >= 0 && x < a.W;
                    const __half *src = inside ? in_n + ((long long)y * a.W + x) * a.ld_cp + kc * 8 : in_n;
                    const uint32_t dst = dst0 + (uint32_t)(kc * Pn + f) * 16u;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(inside ? 16 : 0) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            } else {
                // nearest-upsample-2x(lo) + skip (unet.py:32-33): fp32 add, one rounding
                const __half *in_n = reinterpret_cast<const __half *>(a.in) + n * (long long)a.H * a.W * a.ld_cp;
                const __half *lo_n = a.in_lo + n * (long long)(a.H >> 1) * (a.W >> 1) * a.ld_cp;
                const int items = npos * KC;
                for (int i0 = lt; i0 < items; i0 += 4 * NL) {
                    uint4 v[4], u[4];
                    int dst[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int it = i0 + k * NL;
                        v[k] = make_uint4(0, 0, 0, 0); u[k] = make_uint4(0, 0, 0, 0); dst[k] = -1;
                        if (it < items) {
                            const int f = it / KC, kc = it - f * KC;
                            const int r = (int)__umulhi((unsigned)f, a.pitch_magic), c = f - r * a.pitch;
                            const int y = y0 - 1 + r, x = x0 - 1 + c;
                            dst[k] = kc * Pn + f;
                            if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
                                v[k] = __ldg(reinterpret_cast<const uint4 *>(in_n + ((long long)y * a.W + x) * a.ld_cp + kc * 8));
                                u[k] = __ldg(reinterpret_cast<const uint4 *>(lo_n + ((long long)(y >> 1) * (a.W >> 1) + (x >> 1)) * a.ld_cp + kc * 8));
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (dst[k] < 0) continue;
                        const __half2 *pa = reinterpret_cast<const __half2 *>(&v[k]);
                        const __half2 *pb = reinterpret_cast<const __half2 *>(&u[k]);
                        uint4 r4;
                        __half2 *ro = reinterpret_cast<__half2 *>(&r4);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 fa = __half22float2(pa[e]), fb = __half22float2(pb[e]);
                            ro[e] = __floats2half2_rn(__fadd_rn(fa.x, fb.x), __fadd_rn(fa.y, fb.y));
                        }
                        *reinterpret_cast<uint4 *>(buf + (size_t)dst[k] * 16) = r4;
                    }
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(ld_full);
        }
    }
    // ---- teardown ----------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == kBtEpiWarps) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
    }
}

// =============================================================================================
//  host side
// =============================================================================================
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

struct HostLayer {                      // what imk_unet_create hands over per convolution
    const float *hwio, *bias, *bn_scale, *bn_shift;   // bn_* may be null; host pointers (scale / shift already folded)
    int ks, cin, cout;
};

static void append_par(std::vector<float> &par, const HostLayer &L, int n) {
    const size_t base = par.size();
    par.resize(base + 3 * (size_t)n, 0.f);
    for (int i = 0; i < L.cout; ++i) {
        par[base + i] = L.bias[i];
        par[base + n + i] = L.bn_scale ? L.bn_scale[i] : 1.f;
        par[base + 2 * n + i] = L.bn_shift ? L.bn_shift[i] : 0.f;
    }
    for (int i = L.cout; i < n; ++i) par[base + n + i] = 1.f;      // padding channels: relu(0) * 1 + 0 = 0
}

static bool bt_disabled() {
    const char *v = getenv("IMK_BT_DISABLE");
    return v && v[0] && v[0] != '0';
}

// Chooses the tile and lays out shared memory / TMEM.  Returns false when the block does not fit
// (weights too large to stay resident, or no tile satisfies the 512-column TMEM budget).
static bool bt_plan(FusedBlock &fb, int H, int W) {
    BtArgs &a = fb.args;
    a.H = H; a.W = W;
    const int n1 = a.has_s1 ? a.s1.n : 0, n2 = a.s2.n, n3 = a.s3.n;
    double best = -1.0;
    int bTh = 0, bTw = 0;
    const int th_opts[] = {8, 6, 4, 2};
    for (int th : th_opts) {
        for (int split = 1; split <= 16; split *= 2) {
            int tw = (W + split - 1) / split;
            if (tw < 8 && split > 1) break;
            const int pitch = tw + 2;
            const int nb1 = a.has_s1 ? ((th + 2) * pitch + 127) / 128 : 0, nb2 = (th * pitch + 127) / 128;
            if (nb1 > kBtMaxBlocks || nb2 > kBtMaxBlocks) continue;
            if (nb1 * n1 + nb2 * n2 + nb2 * n3 > 512) continue;
            const int Pn0 = (nb1 * 128) | 1;
            const int Pn1 = std::max(nb1 * 128, nb2 * 128 + 2 * pitch + 2) | 1;
            const int Pn2 = (nb2 * 128) | 1;
            const size_t bytes = (size_t)fb.w_bytes + (size_t)fb.par_floats * 4 + 256 +
                                 (a.has_s1 ? (size_t)Pn0 * (a.s1.ksteps * 2) * 16 : 0) + (size_t)Pn1 * (a.s2.ksteps * 2) * 16 +
                                 (size_t)Pn2 * (a.s3.ksteps * 2) * 16 + (4 + 4 * kBtMaxBlocks) * 8 + 64;
            if (bytes > (size_t)kBtSmemMax) continue;
            const int th_eff = std::min(th, H), tw_eff = std::min(tw, W);
            // useful fraction of the haloed work, with a mild preference for larger tiles (fewer hand-offs)
            const double eff = (double)(th_eff * tw_eff) / ((th + 2.0) * pitch) + 1e-3 * th * tw / (8.0 * 256.0);
            if (eff > best) { best = eff; bTh = th; bTw = tw; }
        }
    }
    if (best < 0) return false;
    a.Th = bTh; a.Tw = bTw; a.pitch = bTw + 2;
    a.pitch_magic = (unsigned)((0x100000000ull + a.pitch - 1) / a.pitch);
    a.tiles_x = (W + bTw - 1) / bTw; a.tiles_y = (H + bTh - 1) / bTh;
    a.s1.nb = a.has_s1 ? ((bTh + 2) * a.pitch + 127) / 128 : 0;
    a.s2.nb = a.s3.nb = (bTh * a.pitch + 127) / 128;
    a.s1.col = 0; a.s2.col = a.s1.nb * n1; a.s3.col = a.s2.col + a.s2.nb * n2;
    a.Pn0 = (a.s1.nb * 128) | 1;
    a.Pn1 = std::max(a.s1.nb * 128, a.s2.nb * 128 + 2 * a.pitch + 2) | 1;
    a.Pn2 = (a.s2.nb * 128) | 1;
    size_t off = (size_t)fb.w_bytes;
    a.par_off_b = (int)off; off += (size_t)fb.par_floats * 4; off = (off + 127) / 128 * 128;
    a.a0_off = (int)off; if (a.has_s1) off += (size_t)a.Pn0 * (a.s1.ksteps * 2) * 16;
    a.a1_off = (int)off; off += (size_t)a.Pn1 * (a.s2.ksteps * 2) * 16;
    a.a2_off = (int)off; off += (size_t)a.Pn2 * (a.s3.ksteps * 2) * 16;
    off = (off + 15) / 16 * 16;
    a.bar_off = (int)off; off += (4 + 4 * kBtMaxBlocks) * 8 + 16;
    fb.smem = off;
    return off <= (size_t)kBtSmemMax;
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
int fused_block_build(FusedBlock &fb, int kind, const ConvHost *L, int H, int W, int in_c, std::vector<void *> &owned) {
    fb = FusedBlock{};
    if (bt_disabled()) return IMK_OK;
    BtArgs &a = fb.args;
    std::vector<__half> w;
    std::vector<float> par;
    auto stage = [&](BtStage &s, const ConvHost &c, bool front) {
        const int cin_p = front ? 16 : pad_ch(c.cin), n = pad_ch(c.cout);
        s.taps = front ? 1 : c.ks * c.ks; s.ksteps = cin_p / 16; s.n = n;
        s.w_off = (int)(w.size() * sizeof(__half));
        if (front) pack_front_b(w, c.hwio, c.cin, c.cout, n); else pack_umma_b(w, c.hwio, c.ks, c.cin, c.cout, cin_p, n);
        s.par_off = (int)par.size();
        HostLayer hl{c.hwio, c.bias, c.bn_scale, c.bn_shift, c.ks, c.cin, c.cout};
        append_par(par, hl, n);
    };
    a.load_kind = kind; a.in_c = in_c;
    if (kind == 1) {
        if (L[0].ks != 3 || L[1].ks != 1) return IMK_OK;
        a.has_s1 = 0;
        stage(a.s2, L[0], false); stage(a.s3, L[1], false);
        a.ld_cp = pad_ch(L[0].cin);
    } else {
        if (L[0].ks != 1 || L[1].ks != 3 || L[2].ks != 1) return IMK_OK;
        if (kind == 0 && 3 * in_c > 16) return IMK_OK;
        a.has_s1 = 1;
        stage(a.s1, L[0], kind == 0); stage(a.s2, L[1], false); stage(a.s3, L[2], false);
        a.ld_cp = kind == 0 ? 16 : pad_ch(L[0].cin);
    }
    if (w.size() * sizeof(__half) % 16) w.resize((w.size() + 7) / 8 * 8, __float2half(0.f));
    fb.w_bytes = a.w_bytes = (int)(w.size() * sizeof(__half));
    fb.par_floats = a.par_floats = (int)par.size();
    if (!bt_plan(fb, H, W)) return IMK_OK;          // does not fit: the layer-wise engine runs this block
    int rc = bt_upload(fb, w, par, owned);
    if (rc) return rc;
    fb.ok = true;
    if (const char *v = getenv("IMK_BT_VERBOSE"); v && v[0] == '1')
        fprintf(stderr, "[imk] fused block kind=%d %dx%d: tile %dx%d, M blocks %d/%d, TMEM %d cols, smem %zu B, weights %d B\n", kind, H, W,
                a.Th, a.Tw, a.s1.nb, a.s2.nb, a.s3.col + a.s3.nb * a.s3.n, fb.smem, fb.w_bytes);
    return IMK_OK;
}

int fused_block_launch(const FusedBlock &fb, const void *in, const __half *in_lo, __half *out, int64_t n, int swap_rb,
                       int in_f32, cudaStream_t stream) {
    BtArgs a = fb.args;
    a.in = in; a.in_lo = in_lo; a.out = out; a.swap_rb = swap_rb; a.in_f32 = in_f32;
    a.n_tiles = (long long)n * a.tiles_x * a.tiles_y;
    if (a.n_tiles <= 0) return IMK_OK;
    static bool attr_set = false;
    if (!attr_set) {
        IMK_CUDA(cudaFuncSetAttribute(block_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBtSmemMax));
        attr_set = true;
    }
    const int grid = (int)std::min<long long>(a.n_tiles, kNumSMs);
    block_tc_kernel<<<grid, kBtThreads, fb.smem, stream>>>(a);
    IMK_LAUNCHED();
    return IMK_OK;
}

}  // namespace imk

// U-Net forward (reference unet.py:46-67, i.e. model.predict at functions.py:3157 /
// 3184 / 3224) and the fused ensemble epilogue.
//
// Data layout in HBM: activations fp16 NHWC with channels padded to a multiple of 16
// (zero in the padding), fp32 accumulation everywhere, one rounding to fp16 per layer
// output.  The 24 convolutions run as
//   - first layer  (c -> f16, 1x1, x/255 fused)          : in_conv_kernel       (CUDA cores, HBM-bound)
//   - hidden 3x3 / 1x1 layers                             : tcgen05 implicit GEMM (imk_conv_tc.cu) or the
//                                                           shared-memory-tiled direct convolution below
//   - last layer   (f16 -> K, 1x1) + sigmoid / softmax    : out_probs_kernel (materialised fp32, .predict)
//                                                           or ensemble_im_kernel (fused with the IM, no fp32 in HBM)
// Order inside a block is Conv -> ReLU -> BatchNorm (unet.py:6-7, 12-16, 34-41): BN is an
// affine epilogue AFTER the ReLU and is not folded into the convolution.
#include <math.h>
#include <algorithm>
#include <vector>
#include "imk_im.cuh"
#include <stdlib.h>
#include "imk_unet.cuh"

namespace imk {

constexpr float kBnEps = 1e-3f;          // Keras BatchNormalization default epsilon

// =============================================================================
//  first layer: Lambda(x/255) -> Conv2D 1x1 (c -> C1) + ReLU -> BN        unet.py:4-9
//  one thread = one pixel x 8 output channels (one 128-bit store)
// =============================================================================
template <typename TIn>
__global__ void __launch_bounds__(256)
in_conv_kernel(const TIn *__restrict__ img, int c, int swap_rb, const float *__restrict__ w /*[c][cout]*/, int cout,
               const float *__restrict__ bias, const float *__restrict__ bn_scale, const float *__restrict__ bn_shift,
               __half *__restrict__ out, int cout_p, int64_t total_px, int normalize = 1) {
    const int chunks = cout_p / 8;
    const int64_t total = total_px * chunks;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t px = i / chunks;
        const int co0 = (int)(i % chunks) * 8;
        float x[4];
        for (int ch = 0; ch < c; ++ch) {
            const int src = (swap_rb && c == 3) ? 2 - ch : ch;
            x[ch] = normalize ? __fdiv_rn((float)img[px * c + src], 255.0f) : (float)img[px * c + src];   // evalnet.py:5-6: normalize is optional there
        }
        __align__(16) __half o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int co = co0 + j;
            float v = 0.f;
            if (co < cout) {
                v = bias[co];
                for (int ch = 0; ch < c; ++ch) v = __fmaf_rn(x[ch], w[ch * cout + co], v);
                v = fmaxf(v, 0.f);
                v = __fmaf_rn(v, bn_scale[co], bn_shift[co]);
            }
            o[j] = __float2half_rn(v);
        }
        *reinterpret_cast<uint4 *>(out + px * cout_p + co0) = *reinterpret_cast<const uint4 *>(o);
    }
}

// =============================================================================
//  hidden layers, direct engine: KS x KS conv (+ optional nearest-upsample-2x + add
//  prologue, unet.py:32-33) + bias + ReLU (+ BN).  16x16 output pixels x 16 output
//  channels per CTA, input halo tile and the weight slice staged in shared memory in
//  chunks of 16 input channels, fp32 accumulation in registers (2 pixels x 16 couts per
//  thread).
// =============================================================================
constexpr int kDT = 16;      // tile edge
constexpr int kDCo = 16;     // output channels per CTA
constexpr int kDCi = 16;     // input channels per shared-memory stage

template <int KS>
__global__ void __launch_bounds__(128)
conv_direct_kernel(const __half *__restrict__ in, const __half *__restrict__ in_lo,
                   const __half *__restrict__ wgt /*[KS*KS][cin_p][cout_p]*/,
                   const float *__restrict__ bias, const float *__restrict__ bn_scale, const float *__restrict__ bn_shift,
                   __half *__restrict__ out, int h, int w, int cin_p, int cout_p, int tiles_x) {
    constexpr int R = KS / 2;
    constexpr int TH = kDT + KS - 1;                    // halo tile edge
    constexpr int PITCH = kDCi / 2 + 1;                 // 32-bit words per pixel (+1: conflict-free)
    __shared__ uint32_t in_s[TH * TH * PITCH];
    __shared__ __align__(16) float w_s[KS * KS * kDCi][kDCo];

    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;                 // ty in 0..7, rows ty and ty+8
    const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int x0 = tile_x * kDT, y0 = tile_y * kDT;
    const int co0 = blockIdx.y * kDCo;
    const int64_t n = blockIdx.z;
    const __half *in_n = in + n * (int64_t)h * w * cin_p;
    const __half *lo_n = in_lo ? in_lo + n * (int64_t)(h / 2) * (w / 2) * cin_p : nullptr;

    float acc0[kDCo], acc1[kDCo];
#pragma unroll
    for (int j = 0; j < kDCo; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }

    for (int cc = 0; cc < cin_p; cc += kDCi) {
        __syncthreads();
        // ---- stage the input halo tile (16 channels) --------------------------------
        for (int i = t; i < TH * TH * 2; i += 128) {
            const int half_sel = i & 1;                 // which 8-channel half (one 128-bit load)
            const int p = i >> 1;
            const int py = p / TH, pxx = p % TH;
            const int gy = y0 + py - R, gx = x0 + pxx - R;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
                v = *reinterpret_cast<const uint4 *>(in_n + ((int64_t)gy * w + gx) * cin_p + cc + half_sel * 8);
                if (lo_n) {
                    const uint4 u = *reinterpret_cast<const uint4 *>(
                        lo_n + ((int64_t)(gy >> 1) * (w >> 1) + (gx >> 1)) * cin_p + cc + half_sel * 8);
                    const __half2 *a = reinterpret_cast<const __half2 *>(&v);
                    const __half2 *b = reinterpret_cast<const __half2 *>(&u);
                    uint4 r;
                    __half2 *ro = reinterpret_cast<__half2 *>(&r);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 fa = __half22float2(a[q]), fb = __half22float2(b[q]);
                        ro[q] = __floats2half2_rn(__fadd_rn(fa.x, fb.x), __fadd_rn(fa.y, fb.y));
                    }
                    v = r;
                }
            }
            uint32_t *dst = in_s + p * PITCH + half_sel * 4;
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
        }
        // ---- stage the weight slice [KS*KS][16 ci][16 co] as fp32 -----------------------
        for (int i = t; i < KS * KS * kDCi * 2; i += 128) {
            const int half_sel = i & 1;
            const int row = i >> 1;                     // tap*16 + ci
            const int tap = row / kDCi, ci = row % kDCi;
            const uint4 v = *reinterpret_cast<const uint4 *>(wgt + ((int64_t)tap * cin_p + cc + ci) * cout_p + co0 + half_sel * 8);
            const __half2 *hv = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 f = __half22float2(hv[q]);
                w_s[row][half_sel * 8 + 2 * q] = f.x;
                w_s[row][half_sel * 8 + 2 * q + 1] = f.y;
            }
        }
        __syncthreads();
        // ---- accumulate ---------------------------------------------------------------
#pragma unroll
        for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
            for (int dx = 0; dx < KS; ++dx) {
                const uint32_t *p0 = in_s + ((ty + dy) * TH + tx + dx) * PITCH;
                const uint32_t *p1 = in_s + ((ty + 8 + dy) * TH + tx + dx) * PITCH;
                const int tap = dy * KS + dx;
#pragma unroll
                for (int c2 = 0; c2 < kDCi / 2; ++c2) {
                    const uint32_t u0 = p0[c2], u1 = p1[c2];
                    const float2 a0 = __half22float2(*reinterpret_cast<const __half2 *>(&u0));
                    const float2 a1 = __half22float2(*reinterpret_cast<const __half2 *>(&u1));
                    const float4 *wr0 = reinterpret_cast<const float4 *>(w_s[tap * kDCi + 2 * c2]);
                    const float4 *wr1 = reinterpret_cast<const float4 *>(w_s[tap * kDCi + 2 * c2 + 1]);
#pragma unroll
                    for (int q = 0; q < kDCo / 4; ++q) {
                        const float4 wa = wr0[q], wb = wr1[q];
                        acc0[4 * q + 0] = fmaf(a0.x, wa.x, acc0[4 * q + 0]); acc1[4 * q + 0] = fmaf(a1.x, wa.x, acc1[4 * q + 0]);
                        acc0[4 * q + 1] = fmaf(a0.x, wa.y, acc0[4 * q + 1]); acc1[4 * q + 1] = fmaf(a1.x, wa.y, acc1[4 * q + 1]);
                        acc0[4 * q + 2] = fmaf(a0.x, wa.z, acc0[4 * q + 2]); acc1[4 * q + 2] = fmaf(a1.x, wa.z, acc1[4 * q + 2]);
                        acc0[4 * q + 3] = fmaf(a0.x, wa.w, acc0[4 * q + 3]); acc1[4 * q + 3] = fmaf(a1.x, wa.w, acc1[4 * q + 3]);
                        acc0[4 * q + 0] = fmaf(a0.y, wb.x, acc0[4 * q + 0]); acc1[4 * q + 0] = fmaf(a1.y, wb.x, acc1[4 * q + 0]);
                        acc0[4 * q + 1] = fmaf(a0.y, wb.y, acc0[4 * q + 1]); acc1[4 * q + 1] = fmaf(a1.y, wb.y, acc1[4 * q + 1]);
                        acc0[4 * q + 2] = fmaf(a0.y, wb.z, acc0[4 * q + 2]); acc1[4 * q + 2] = fmaf(a1.y, wb.z, acc1[4 * q + 2]);
                        acc0[4 * q + 3] = fmaf(a0.y, wb.w, acc0[4 * q + 3]); acc1[4 * q + 3] = fmaf(a1.y, wb.w, acc1[4 * q + 3]);
                    }
                }
            }
        }
    }
    // ---- epilogue: + bias, ReLU, BN affine, one rounding to fp16 -----------------------
    auto store = [&](const float *acc, int gy, int gx) {
        if (gy >= h || gx >= w) return;
        __align__(16) __half o[kDCo];
#pragma unroll
        for (int j = 0; j < kDCo; ++j) {
            float v = fmaxf(acc[j] + bias[co0 + j], 0.f);
            if (bn_scale) v = __fmaf_rn(v, bn_scale[co0 + j], bn_shift[co0 + j]);
            o[j] = __float2half_rn(v);
        }
        uint4 *dst = reinterpret_cast<uint4 *>(out + (n * (int64_t)h * w + (int64_t)gy * w + gx) * cout_p + co0);
        dst[0] = reinterpret_cast<const uint4 *>(o)[0];
        dst[1] = reinterpret_cast<const uint4 *>(o)[1];
    };
    store(acc0, y0 + ty, x0 + tx);
    store(acc1, y0 + ty + 8, x0 + tx);
}

// MaxPooling2D((2,2)) on fp16 NHWC  (unet.py:17); one thread = one output pixel x 8 channels
__global__ void __launch_bounds__(256)
maxpool_kernel(const __half *__restrict__ in, __half *__restrict__ out, int64_t n, int h, int w, int cp) {
    const int ho = h / 2, wo = w / 2, chunks = cp / 8;
    const int64_t total = n * ho * wo * chunks;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int ch = (int)(i % chunks) * 8;
        int64_t r = i / chunks;
        const int xo = (int)(r % wo); r /= wo;
        const int yo = (int)(r % ho);
        const int64_t img = r / ho;
        const __half *base = in + ((img * h + 2 * yo) * (int64_t)w + 2 * xo) * cp + ch;
        const uint4 a = *reinterpret_cast<const uint4 *>(base);
        const uint4 b = *reinterpret_cast<const uint4 *>(base + cp);
        const uint4 c = *reinterpret_cast<const uint4 *>(base + (int64_t)w * cp);
        const uint4 d = *reinterpret_cast<const uint4 *>(base + (int64_t)w * cp + cp);
        uint4 o;
        const __half2 *pa = reinterpret_cast<const __half2 *>(&a), *pb = reinterpret_cast<const __half2 *>(&b);
        const __half2 *pc = reinterpret_cast<const __half2 *>(&c), *pd = reinterpret_cast<const __half2 *>(&d);
        __half2 *po = reinterpret_cast<__half2 *>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) po[q] = __hmax2(__hmax2(pa[q], pb[q]), __hmax2(pc[q], pd[q]));
        *reinterpret_cast<uint4 *>(out + ((img * ho + yo) * (int64_t)wo + xo) * cp + ch) = o;
    }
}

// =============================================================================
//  last layer: Conv2D 1x1 (C1 -> K) + sigmoid / softmax in fp32        unet.py:63
//  The arithmetic is written with explicit round-to-nearest intrinsics so that the
//  materialised (.predict) kernel and the fused ensemble kernel produce the same bits.
// =============================================================================
struct OutParams {                      // per model, in shared memory: w[K][c1p] then b[K]
    const float *w;                     // device [K][c1p] fp32 (zero in the channel padding)
    const float *b;                     // device [K]
};

// run-time shapes: one thread = one pixel, sequential FMAs
template <int KMAX>
__device__ __forceinline__ void pixel_probs(const __half *__restrict__ x /*c1p halves of one pixel*/, int c1p,
                                            const float *__restrict__ w_s, const float *__restrict__ b_s, int K,
                                            int act, float (&p)[KMAX]) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) p[k] = (k < K) ? b_s[k] : 0.f;
    for (int c0 = 0; c0 < c1p; c0 += 8) {
        const uint4 v = *reinterpret_cast<const uint4 *>(x + c0);
        const __half2 *hv = reinterpret_cast<const __half2 *>(&v);
        float xf[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 f = __half22float2(hv[q]); xf[2 * q] = f.x; xf[2 * q + 1] = f.y; }
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                const float4 wa = *reinterpret_cast<const float4 *>(w_s + k * c1p + c0);
                const float4 wb = *reinterpret_cast<const float4 *>(w_s + k * c1p + c0 + 4);
                const float wk[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) p[k] = __fmaf_rn(xf[j], wk[j], p[k]);
            }
        }
    }
    pixel_activation<KMAX>(p, K, act);
}

// -----------------------------------------------------------------------------
//  The reference's heads (K = 1, 3, 9, 35 on 16 / 32 channels) on the tensor cores: the last layer of a warp's 32
//  pixels is a [32 x C1] x [C1 x K] product, i.e. 2 m-tiles x ceil(K/8) n-tiles of mma.sync.m16n8k16 (fp16 operands,
//  fp32 accumulation).  The activations ARE fp16; the fp32 weights enter as hi + lo halves (two MMAs into the same
//  accumulator), which keeps ~22 bits of them.  Pixels go through a per-warp shared-memory tile (ldmatrix on the way
//  in, one row per lane on the way out), so everything after the logits stays one-thread-per-pixel.  Both the
//  materialised (.predict) kernel and the fused ensemble kernel use THIS function for these shapes: identical bits.
// -----------------------------------------------------------------------------
template <int KFIX, int C1FIX>
struct HeadMma {
    static constexpr int NT = (KFIX + 7) / 8;                    // n tiles of 8 classes
    static constexpr int KS = C1FIX / 16;                        // k steps of 16 channels
    static constexpr int XV = C1FIX / 8;                         // 16-byte vectors per pixel
    static constexpr int APITCH = C1FIX * 2 + 16;                // bytes per pixel row of the A tile: 16-byte aligned, ldmatrix conflict-free
    static constexpr int CP = NT * 8 + 1;                        // floats per pixel row of the C tile (odd: conflict-free row reads)
    static constexpr int WARP_BYTES = (32 * APITCH + 32 * CP * 4 + 15) / 16 * 16;
    static constexpr int BFRAG_WORDS = NT * KS * 4 * 32;         // per model: [n tile][k step][b0_hi, b1_hi, b0_lo, b1_lo][lane]

    // operand-B fragments of one model from its fp32 [K][C1] weights (all threads of the CTA)
    static __device__ __forceinline__ void prepare(const float *__restrict__ w /*[K][C1]*/, uint32_t *__restrict__ bfrag) {
        for (int idx = threadIdx.x; idx < BFRAG_WORDS; idx += blockDim.x) {
            const int lane = idx & 31, q = (idx >> 5) & 3, js = idx >> 7, s = js % KS, j = js / KS;
            const int g = lane >> 2, t = lane & 3;
            const int col = j * 8 + g, r0 = s * 16 + 2 * t + ((q & 1) ? 8 : 0);
            float v0 = 0.f, v1 = 0.f;
            if (col < KFIX) { v0 = w[col * C1FIX + r0]; v1 = w[col * C1FIX + r0 + 1]; }
            __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            if (q >= 2) { h0 = __float2half_rn(v0 - __half2float(h0)); h1 = __float2half_rn(v1 - __half2float(h1)); }
            bfrag[idx] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        }
    }

    // logits of the warp's 32 pixels (lane = pixel; xv = the pixel's C1 halves, zeros for dead lanes): p[k] = b[k] + x . w[k]
    static __device__ __forceinline__ void logits(const uint4 (&xv)[C1FIX / 8], uint8_t *__restrict__ wsm, const uint32_t *__restrict__ bfrag,
                                                  const float *__restrict__ bias, float (&p)[KFIX]) {
        const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
        uint8_t *At = wsm;
        float *Ct = reinterpret_cast<float *>(wsm + 32 * APITCH);
#pragma unroll
        for (int i = 0; i < C1FIX / 8; ++i) *reinterpret_cast<uint4 *>(At + lane * APITCH + 16 * i) = xv[i];
        __syncwarp();
        uint32_t a[2][KS][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kb = s * 32 + (lane >> 4) * 16;
                const uint32_t addr = smem_u32(At + row * APITCH + kb);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(a[mt][s][0]), "=r"(a[mt][s][1]), "=r"(a[mt][s][2]), "=r"(a[mt][s][3]) : "r"(addr));
            }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            uint32_t b[KS][4];
#pragma unroll
            for (int s = 0; s < KS; ++s)
#pragma unroll
                for (int q = 0; q < 4; ++q) b[s][q] = bfrag[((j * KS + s) * 4 + q) * 32 + lane];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int s = 0; s < KS; ++s)
#pragma unroll
                    for (int hl = 0; hl < 2; ++hl)
                        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                                     : "r"(a[mt][s][0]), "r"(a[mt][s][1]), "r"(a[mt][s][2]), "r"(a[mt][s][3]), "r"(b[s][2 * hl]), "r"(b[s][2 * hl + 1]));
                float *c0 = Ct + (mt * 16 + g) * CP + j * 8 + 2 * t;
                c0[0] = c[0]; c0[1] = c[1]; c0[8 * CP] = c[2]; c0[8 * CP + 1] = c[3];
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < KFIX; ++k) p[k] = __fadd_rn(Ct[lane * CP + k], bias[k]);
        __syncwarp();                                            // the tiles are reused by the next model / pixel batch
    }
};

// c9 as ONE 8-channel plane (imk_unet::c8, the alpha = 0.5 networks): 8 K FMAs per pixel are cheaper than staging an MMA
// tile.  Same interface; the fp32 weights [K][8] sit right before the bias in shared memory (both kernels lay them out so).
template <int KFIX>
struct HeadMma<KFIX, 8> {
    static constexpr int NT = 0, KS = 0, XV = 1, WARP_BYTES = 0, BFRAG_WORDS = 0;
    static __device__ __forceinline__ void prepare(const float *__restrict__, uint32_t *__restrict__) {}
    static __device__ __forceinline__ void logits(const uint4 (&xv)[1], uint8_t *__restrict__, const uint32_t *__restrict__,
                                                  const float *__restrict__ bias, float (&p)[KFIX]) {
        const float *w = bias - KFIX * 8;
        const __half2 *hv = reinterpret_cast<const __half2 *>(&xv[0]);
        float xf[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 f = __half22float2(hv[q]); xf[2 * q] = f.x; xf[2 * q + 1] = f.y; }
#pragma unroll
        for (int k = 0; k < KFIX; ++k) {
            float acc = bias[k];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = __fmaf_rn(xf[j], w[k * 8 + j], acc);
            p[k] = acc;
        }
    }
};

// KFIX / C1FIX > 0: class count and padded input width known at compile time (the reference's heads: K = 1, 3, 9, 35 on
// 16 or 32 channels) -- loops unroll, the `k < K` predicates vanish; 0: run-time values, KMAX bounds the registers
template <int KMAX, int KFIX, int C1FIX>
__global__ void __launch_bounds__(256, 2)
out_probs_kernel(const __half *__restrict__ c9, int c1p_, const float *__restrict__ w, const float *__restrict__ b,
                 int K_, int act, float *__restrict__ probs, int64_t total_px) {
    const int K = KFIX > 0 ? KFIX : K_, c1p = C1FIX > 0 ? C1FIX : c1p_;
    extern __shared__ __align__(16) float osm[];
    float *w_s = osm, *b_s = osm + K * c1p;
    for (int i = threadIdx.x; i < K * c1p; i += blockDim.x) w_s[i] = w[i];
    for (int i = threadIdx.x; i < K; i += blockDim.x) b_s[i] = b[i];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if constexpr (KFIX > 0) {
        using H = HeadMma<KFIX, C1FIX>;
        uint32_t *bfrag = reinterpret_cast<uint32_t *>(osm + (K * c1p + K + 3) / 4 * 4);
        uint8_t *wsm = reinterpret_cast<uint8_t *>(bfrag + H::BFRAG_WORDS) + (threadIdx.x >> 5) * H::WARP_BYTES;
        H::prepare(w, bfrag);
        __syncthreads();
        const int64_t rounded = (total_px + 31) / 32 * 32;      // whole warps take part in the MMAs
        for (int64_t px = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; px < rounded; px += stride) {
            const bool live = px < total_px;
            uint4 xv[C1FIX / 8];
#pragma unroll
            for (int i = 0; i < C1FIX / 8; ++i) xv[i] = live ? *reinterpret_cast<const uint4 *>(c9 + px * C1FIX + 8 * i) : make_uint4(0, 0, 0, 0);
            float p[KFIX];
            H::logits(xv, wsm, bfrag, b_s, p);
            if (live) {
                pixel_activation<KFIX>(p, KFIX, act);
                float *dst = probs + px * KFIX;
#pragma unroll
                for (int k = 0; k < KFIX; ++k) dst[k] = p[k];
            }
        }
    } else {
        __syncthreads();
        for (int64_t px = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; px < total_px; px += stride) {
            float p[KMAX];
            pixel_probs<KMAX>(c9 + px * c1p, c1p, w_s, b_s, K, act, p);
            float *dst = probs + px * K;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) if (k < K) dst[k] = p[k];
        }
    }
}

// =============================================================================
//  fused ensemble epilogue: for every pixel, every model's last layer + activation +
//  threshold / argmax + ensemble agreement + IM + blanking + per-image sizes.  Reads the
//  M fp16 c9 maps and the uint8 image, writes only uint8 maps: no fp32 map touches HBM.
//  256 pixels per CTA iteration; label / IM bytes are regrouped through shared memory so
//  that all uint8 traffic is 128-bit.
// =============================================================================
struct EnsPtrs {
    const __half *c9[IMK_MAX_MODELS];
    const float *w[IMK_MAX_MODELS];
    const float *b[IMK_MAX_MODELS];
};

template <int KMAX, bool kMulticlass, int KFIX, int C1FIX>
__global__ void __launch_bounds__(256, 2)
ensemble_im_kernel(EnsPtrs ens, int M, int c1p_, int K_, int act, float thr, int strict, float dstar,
                   int64_t total_px, int64_t HW, int64_t N, int64_t plane_stride,
                   const uint8_t *__restrict__ img, int c, int block_in, int block_out,
                   uint8_t *__restrict__ img_out, uint8_t *__restrict__ labels, uint8_t *__restrict__ im_out,
                   int64_t *__restrict__ im_size, int64_t *__restrict__ pred_size,
                   unsigned long long *__restrict__ presence) {
    // labels: plane k of this launch starts at labels + k * plane_stride; pred_size plane k at pred_size + k * N
    const int K = KFIX > 0 ? KFIX : K_, c1p = C1FIX > 0 ? C1FIX : c1p_;
    extern __shared__ __align__(16) float esm[];
    const int per_model = (K * c1p + K + 3) / 4 * 4;             // 16-byte aligned rows for the 128-bit weight reads
    float *w_all = esm;
    uint8_t *lab_s = reinterpret_cast<uint8_t *>(esm + (size_t)M * per_model);   // [8 warps][32] class ids (multiclass)
    for (int m = 0; m < M; ++m) {
        for (int i = threadIdx.x; i < K * c1p; i += blockDim.x) w_all[m * per_model + i] = ens.w[m][i];
        for (int i = threadIdx.x; i < K; i += blockDim.x) w_all[m * per_model + K * c1p + i] = ens.b[m][i];
    }
    // tensor-core head (compile-time shapes): operand-B fragments of every model, one A / C tile pair per warp
    using H = HeadMma<(KFIX > 0 ? KFIX : 1), (C1FIX > 0 ? C1FIX : 16)>;
    uint32_t *bfrag = reinterpret_cast<uint32_t *>(lab_s + 256);
    uint8_t *wsm = reinterpret_cast<uint8_t *>(bfrag + (size_t)M * H::BFRAG_WORDS) + (threadIdx.x >> 5) * H::WARP_BYTES;
    if constexpr (KFIX > 0)
        for (int m = 0; m < M; ++m) H::prepare(ens.w[m], bfrag + (size_t)m * H::BFRAG_WORDS);
    __syncthreads();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = (total_px + 255) / 256;
    constexpr int NL = kMulticlass ? 1 : 3;
    // image vector of the warp-local epilogue: lanes < 2c own one 16-byte vector of the warp's 32 pixels
    const int ig = lane >= c ? 1 : 0, iv = lane - ig * c;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t p0 = tile * 256;
        const int64_t px = p0 + tid;
        const bool live = px < total_px;
        const int64_t wpx = p0 + warp * 32;                      // first pixel of this warp (total_px % 16 == 0 on this path)
        // requested before the models are evaluated: the DRAM latency hides behind the arithmetic
        const bool img_lane = img_out && lane < 2 * c && wpx + 16 * ig < total_px;
        uint4 pix = make_uint4(0, 0, 0, 0);
        if (img_lane) pix = ldg_stream(img + (wpx + 16 * ig) * c + 16 * iv);
        uint32_t im_any = 0, im_cnt = 0;
        uint32_t lab[NL];
#pragma unroll
        for (int l = 0; l < NL; ++l) lab[l] = 0;
        int a0 = 0;
        uint32_t votes[NL];
#pragma unroll
        for (int l = 0; l < NL; ++l) votes[l] = 0;
        // image index: one 32-bit division per tile (the 64-bit one is a subroutine; chunks are far below 2^31 pixels)
        const int64_t n_first = (int64_t)((uint32_t)p0 / (uint32_t)HW);
        const bool uniform = (p0 + 256 <= (n_first + 1) * HW) && (p0 + 256 <= total_px);
        const int64_t n = live ? (uniform ? n_first : (int64_t)((uint32_t)px / (uint32_t)HW)) : -1;
        const int64_t n_u = uniform ? n_first : n;
        uint4 x_next = make_uint4(0, 0, 0, 0);
        if constexpr (KFIX > 0 && H::XV == 1) { if (live) x_next = ldg_stream(ens.c9[0] + px * c1p); }
        for (int m = 0; m < M; ++m) {
            int arg = 0;
            bool fast_arg = false;                              // arg already holds the argmax of the probabilities
            float p[KMAX];
            if constexpr (KFIX > 0) {                           // every lane takes part in the MMAs (dead lanes feed zeros)
                uint4 xv[H::XV];
                if constexpr (H::XV == 1) {
                    // one 16-byte plane per model: the next model's vector is requested before this one is used
                    xv[0] = x_next;
                    x_next = (live && m + 1 < M) ? ldg_stream(ens.c9[m + 1] + px * c1p) : make_uint4(0, 0, 0, 0);
                } else {
#pragma unroll
                    for (int i = 0; i < H::XV; ++i)
                        xv[i] = live ? *reinterpret_cast<const uint4 *>(ens.c9[m] + px * c1p + 8 * i) : make_uint4(0, 0, 0, 0);
                }
                H::logits(xv, wsm, bfrag + (size_t)m * H::BFRAG_WORDS, w_all + m * per_model + K * c1p, p);
                if (live) {
                    if (kMulticlass && act == IMK_ACT_SOFTMAX) {
                        // argmax of the softmax without its K divisions: p_k = e_k / sum is monotonic in e_k, and two
                        // numerators more than 2^-22 apart (relative) cannot round to the same quotient, so the first
                        // maximum of e is the first maximum of p unless another numerator is that close to it (or a
                        // NaN is around) -- only then the exact probabilities are formed, with pixel_activation's ops
                        // First, on the LOGITS: exp is monotonic and the winner's numerator is exp(0) = 1 exactly, so a runner-up
                        // more than 1e-5 below the winner has a numerator < 1 - 9.7e-6 (with __expf's 2-ulp error) and its
                        // rounded quotient is strictly smaller: the first maximum of the logits IS np.argmax of the
                        // probabilities, and none of the K exponentials is needed.  NaN / infinite logits (the probabilities are
                        // all NaN then) and closer calls take the numerator path below.
                        float lbest = p[0], lsecond = -INFINITY, lsum = p[0];
#pragma unroll
                        for (int k = 1; k < KMAX; ++k) {
                            lsum += p[k];
                            if (p[k] > lbest) { lsecond = lbest; lbest = p[k]; arg = k; } else lsecond = fmaxf(lsecond, p[k]);
                        }
                        fast_arg = (lbest - lsecond > 1e-5f) && (lsum == lsum) && (lbest < INFINITY);
                        if (!fast_arg) {
                        arg = 0;
                        float mx = p[0];
#pragma unroll
                        for (int k = 1; k < KMAX; ++k) mx = fmaxf(mx, p[k]);
                        float sum = 0.f;
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) { p[k] = __expf(__fsub_rn(p[k], mx)); sum = __fadd_rn(sum, p[k]); }
                        float best = p[0];
#pragma unroll
                        for (int k = 1; k < KMAX; ++k) if (p[k] > best) { best = p[k]; arg = k; }
                        const float lim = best * 0.99999976f;
                        int close = 0;
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) close += (p[k] >= lim) ? 1 : 0;
                        fast_arg = close == 1 && sum == sum;
                        if (!fast_arg) {
#pragma unroll
                            for (int k = 0; k < KMAX; ++k) p[k] = __fdiv_rn(p[k], sum);
                        }
                        }
                    } else if (!kMulticlass && act == IMK_ACT_SIGMOID && dstar > 0.f) {
                        // threshold of the sigmoid without its division: RN(1 / d) is monotonic in d = 1 + exp(-z), so
                        // "p >= thr" (or ">") is exactly "d <= dstar" with dstar found on the host by exact fp32 division;
                        // the vote is stored as 1.0 / 0.0 so that decide() below sees p >= thr <=> vote
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) {
                            const float dk = __fadd_rn(1.0f, __expf(-p[k]));
                            p[k] = dk <= dstar ? 2.0f : -1.0f;       // any value above / below every threshold in (0, 1)
                        }
                    } else {
                        pixel_activation<KMAX>(p, K, act);
                    }
                }
            } else if (live) {
                pixel_probs<KMAX>(ens.c9[m] + px * c1p, c1p, w_all + m * per_model, w_all + m * per_model + K * c1p, K, act, p);
            }
            if (live) {
                if (kMulticlass) {
                    if (!fast_arg) {
                        arg = 0;
                        float best = p[0];
#pragma unroll
                        for (int k = 1; k < KMAX; ++k) if (k < K) argmax_step(p[k], k, best, arg);
                    }
                    if (m == 0) a0 = arg; else im_any |= (arg != a0);
                } else {
#pragma unroll
                    for (int k = 0; k < NL; ++k) if (k < K) votes[k] += decide(p[k], thr, strict != 0);
                }
            }
            if (kMulticlass && presence)
                warp_or_stat(presence, n_u < 0 ? -1 : n_u * M + m, live ? (1ull << (arg & 63)) : 0ull, uniform);
        }
        if (kMulticlass) {
            im_cnt = im_any;
            lab[0] = im_any ? 0 : (uint32_t)a0;
        } else {
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                if (k < K) {
                    const uint32_t all = votes[k] == (uint32_t)M;
                    const uint32_t mixed = (votes[k] != 0u) & (votes[k] != (uint32_t)M);
                    lab[k] = all;
                    im_any |= mixed;
                    im_cnt += mixed;
                }
            }
        }
        if (!live) { im_cnt = 0; im_any = 0; }
        warp_add_stat(im_size, n_u, im_cnt, uniform);
        if (!kMulticlass && pred_size) {
#pragma unroll
            for (int k = 0; k < NL; ++k)
                if (k < K) warp_add_stat(pred_size + (int64_t)k * N, n_u, live ? lab[k] : 0u, uniform);
        }
        // warp-local 128-bit epilogue over the warp's 32 consecutive pixels: ballots give the 0/255 planes, lanes 0-1
        // store 16 pixels each, lanes < 2c blank the image; no block-wide barrier on the hot path
        const uint32_t im_mask = __ballot_sync(0xffffffffu, im_any != 0);
        const bool vec_lane = lane < 2 && wpx + 16 * lane < total_px;
        auto expand16 = [](uint32_t bits) {                      // 16 bits -> 16 bytes of 0x00 / 0xFF
            return make_uint4(bytes01_to_ff(bits4_to_bytes01(bits)), bytes01_to_ff(bits4_to_bytes01(bits >> 4)),
                              bytes01_to_ff(bits4_to_bytes01(bits >> 8)), bytes01_to_ff(bits4_to_bytes01(bits >> 12)));
        };
        if (kMulticlass) {
            lab_s[warp * 32 + lane] = (uint8_t)lab[0];
            __syncwarp();
            if (vec_lane) stg_stream(labels + wpx + 16 * lane, *reinterpret_cast<const uint4 *>(lab_s + warp * 32 + 16 * lane));
            __syncwarp();
        } else {
#pragma unroll
            for (int k = 0; k < NL; ++k) {
                if (k < K) {
                    // head 2 (HeLa position) stays raw: the reference blanks the circle image drawn from it on the host
                    const uint32_t mk = __ballot_sync(0xffffffffu, lab[k] && !(block_out && k < 2 && im_any));
                    if (vec_lane) stg_stream(labels + (int64_t)k * plane_stride + wpx + 16 * lane, expand16(mk >> (16 * lane)));
                }
            }
        }
        if (vec_lane) stg_stream(im_out + wpx + 16 * lane, expand16(im_mask >> (16 * lane)));
        if (img_lane) {
            const uint4 imv = expand16(block_in ? (im_mask >> (16 * ig)) : 0u);
            const uint32_t imw[4] = {imv.x, imv.y, imv.z, imv.w};
            uint4 o;                                             // compile-time channel count: the byte selectors fold to constants
            switch (c) {
                case 1: o = blank_vec_sel<1>(pix, imw, 1, iv); break;
                case 2: o = blank_vec_sel<2>(pix, imw, 2, iv); break;
                case 3: o = blank_vec_sel<3>(pix, imw, 3, iv); break;
                default: o = blank_vec_sel<4>(pix, imw, 4, iv); break;
            }
            stg_stream(img_out + (wpx + 16 * ig) * c + 16 * iv, o);
        }
    }
}

// =============================================================================
//  ensemble epilogue on head-stage decisions: every model's level-0 decoder kernel has already run its output layer
//  and left ONE byte per pixel (bit k = head k fires, or the argmax class id).  This kernel only counts votes:
//  reads M bytes + the image per pixel, writes label(s), IM, blanked image (uint8) and the per-image sizes.
//  One thread = 16 consecutive pixels, every access 128-bit and warp-contiguous; SIMD-in-word byte arithmetic.
//  MODE 0: K = 1 (functions.py:3104-3120)   1: K = 3 HeLa (functions.py:3185-3200)   2: multiclass (functions.py:3123-3137)
// =============================================================================
struct DecPtrs { const uint8_t *d[IMK_MAX_MODELS]; };

template <int MODE>
__global__ void __launch_bounds__(256)
ensemble_votes_kernel(DecPtrs dec, int M, int64_t total_px, int64_t HW, int64_t N, int64_t plane_stride,
                      const uint8_t *__restrict__ img, int c, int block_in, int block_out,
                      uint8_t *__restrict__ img_out, uint8_t *__restrict__ labels, uint8_t *__restrict__ im_out,
                      int64_t *__restrict__ im_size, int64_t *__restrict__ pred_size, unsigned long long *__restrict__ presence) {
    constexpr int NK = MODE == 1 ? 3 : 1;
    const int64_t n_groups = total_px / 16;                          // H, W are multiples of 16: a group never straddles images
    const int64_t n_iter = (n_groups + 31) / 32 * 32;                // whole warps (the statistics are warp reductions)
    const uint32_t m_rep = (uint32_t)M * 0x01010101u;
    for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < n_iter; gi += (int64_t)gridDim.x * blockDim.x) {
        const bool live = gi < n_groups;
        const int64_t px = gi * 16;
        const int64_t n = live ? (int64_t)((uint32_t)px / (uint32_t)HW) : -1;       // chunks stay far below 2^31 pixels
        uint32_t first[4] = {0, 0, 0, 0};
        uint32_t cnt[NK][4], diff[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < NK; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) cnt[k][j] = 0;
        for (int m = 0; m < M; ++m) {
            const uint4 v = live ? ldg_stream(dec.d[m] + px) : make_uint4(0, 0, 0, 0);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            if (MODE == 2) {
                if (m == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) first[j] = w[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) diff[j] |= w[j] ^ first[j];
                }
                if (presence) {                                      // class sets for lists_equal (functions.py:3226-3234)
                    unsigned long long bits = 0ull;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int e = 0; e < 4; ++e) bits |= 1ull << ((w[j] >> (8 * e)) & 63u);
                    warp_or_stat(presence, n < 0 ? -1 : n * M + m, live ? bits : 0ull, false);
                }
            } else {
#pragma unroll
                for (int k = 0; k < NK; ++k)
#pragma unroll
                    for (int j = 0; j < 4; ++j) cnt[k][j] += (w[j] >> k) & 0x01010101u;
            }
        }
        uint32_t lab[NK][4], im01[4] = {0, 0, 0, 0};
        uint32_t im_cnt = 0, pred_cnt[NK];
        if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                im01[j] = bytes_nonzero01(diff[j]);                  // class ids and their XORs are < 0x80
                lab[0][j] = first[j] & ~bytes01_to_ff(im01[j]);      // agree ? class id : 0
                im_cnt += __popc(im01[j]);
            }
            pred_cnt[0] = 0;
        } else {
            uint32_t all01[NK][4];
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                pred_cnt[k] = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t nz = bytes_nonzero01(cnt[k][j]);                 // S != 0
                    const uint32_t ne = bytes_nonzero01(cnt[k][j] ^ m_rep);         // S != M
                    all01[k][j] = ne ^ 0x01010101u;
                    const uint32_t mix = nz & ne;
                    im01[j] |= mix;                                  // combined IM = max over heads
                    im_cnt += __popc(mix);                           // HeLa: the SUM of the three head IM sizes
                    pred_cnt[k] += __popc(all01[k][j]);
                }
            }
#pragma unroll
            for (int k = 0; k < NK; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j)      // head 2 (HeLa position) stays raw: the host blanks the circle image drawn from it
                    lab[k][j] = bytes01_to_ff((block_out && k < 2) ? (all01[k][j] & ~im01[j]) : all01[k][j]);
        }
        warp_add_stat(im_size, n, live ? im_cnt : 0u, false);
        if (MODE != 2 && pred_size) {
#pragma unroll
            for (int k = 0; k < NK; ++k) warp_add_stat(pred_size + (int64_t)k * N, n, live ? pred_cnt[k] : 0u, false);
        }
        if (live) {
            uint32_t imw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) imw[j] = bytes01_to_ff(im01[j]);
#pragma unroll
            for (int k = 0; k < NK; ++k)
                stg_stream(labels + (int64_t)k * plane_stride + px, make_uint4(lab[k][0], lab[k][1], lab[k][2], lab[k][3]));
            stg_stream(im_out + px, make_uint4(imw[0], imw[1], imw[2], imw[3]));
            if (img_out) blank_image16_any(img, img_out, c, px, imw, block_in != 0);
        }
    }
}

__global__ void lists_equal_kernel2(const unsigned long long *__restrict__ presence, int M, int64_t N,
                                    uint8_t *__restrict__ lists_equal) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    bool eq = true;
    for (int m = 1; m < M; ++m) eq &= presence[n * M + m] == presence[n * M];
    lists_equal[n] = eq ? 1 : 0;
}

// =============================================================================
//  host side: plan, packing, workspace, trunk
// =============================================================================
static int grid_1d(int64_t items, int per_block = 256, int per_sm = 8) {
    int64_t b = (items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)kNumSMs * per_sm;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename T>
static int upload(std::vector<void *> &owned, const std::vector<T> &host, T **dev) {
    void *p = nullptr;
    if (cudaMalloc(&p, host.size() * sizeof(T) + 16) != cudaSuccess) { set_error("cudaMalloc(%zu) failed", host.size() * sizeof(T)); return IMK_ENOMEM; }
    owned.push_back(p);
    IMK_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dev = reinterpret_cast<T *>(p);
    return IMK_OK;
}

struct PlanItem { bool is_conv; int ks, cin, cout; };

// creation order of the parameterised layers, unet.py:49-63
static std::vector<PlanItem> make_plan(const imk_unet_desc &d, const int f[5]) {
    std::vector<PlanItem> p;
    auto conv = [&](int ks, int ci, int co) { p.push_back({true, ks, ci, co}); };
    auto bn = [&](int ch) { p.push_back({false, 0, ch, ch}); };
    conv(1, d.in_channels, f[0]); bn(f[0]);                                   // input_block
    int cin = f[0];
    for (int l = 0; l < 4; ++l) { conv(d.ks, cin, f[l]); conv(1, f[l], f[l]); bn(f[l]); cin = f[l]; }   // encoder_block x4
    conv(d.ks, cin, f[4]); conv(1, f[4], f[3]); bn(f[3]);                     // bottleneck_block
    cin = f[3];
    const int c1s[4] = {f[3], f[2], f[1], f[0]}, c2s[4] = {f[2], f[1], f[0], f[0]};
    for (int l = 0; l < 4; ++l) {                                             // decoder_block x4
        conv(1, cin, c1s[l]); bn(c1s[l]); conv(d.ks, c1s[l], c1s[l]); conv(1, c1s[l], c2s[l]); bn(c2s[l]);
        cin = c2s[l];
    }
    conv(1, cin, d.num_outputmasks);                                          // 'out'
    return p;
}

static int64_t g_max_chunk = 0;
static int64_t default_chunk() {
    int64_t v = 512;                         // measured (HeLa): 64 -> 35.5k, 128 -> 39.6k, 256 -> 42.5k, 512 -> 43.7k, 1024 -> 44.7k img/s
    if (const char *e = getenv("IMK_CHUNK"); e && e[0]) v = atoll(e);
    return v < 1 ? 1 : (v > 1024 ? 1024 : v);
}
int64_t max_chunk() {
    if (!g_max_chunk) g_max_chunk = default_chunk();
    return g_max_chunk;
}

}  // namespace imk
extern "C" int64_t imk_max_chunk(void) { return imk::max_chunk(); }
extern "C" int imk_set_max_chunk(int64_t n) {
    IMK_REQUIRE(n >= 0 && n <= 1024, "imk_set_max_chunk: %lld outside 0..1024", (long long)n);
    imk::g_max_chunk = n ? n : imk::default_chunk();
    return IMK_OK;
}
namespace imk {

int unet_reserve(imk_unet *net, int64_t n) {
    if (n <= net->cap_n) return IMK_OK;
    if (net->ws) { cudaFree(net->ws); net->ws = nullptr; net->cap_n = 0; net->dec = nullptr; }
    size_t total = 0;
    size_t off[5][3];
    for (int l = 0; l < 5; ++l) {
        const int h = net->desc.height >> l, w = net->desc.width >> l;
        const int chp = pad_ch(l < 4 ? net->widths[l] : net->widths[4]);
        // level 4 buffer `b` holds the 256a-wide bottleneck map, `a` the 128a-wide ones
        for (int j = 0; j < 3; ++j) {
            size_t bytes = (size_t)n * h * w * chp * sizeof(__half);
            bytes = (bytes + 255) / 256 * 256;
            off[l][j] = total;
            total += bytes;
        }
        net->lvl[l].h = h; net->lvl[l].w = w; net->lvl[l].ch_p = chp;
    }
    const size_t dec_off = total;
    total += ((size_t)n * net->desc.height * net->desc.width + 255) / 256 * 256;
    // IMK_WS_LIMIT_MB: refuse workspaces above this size (deployments that share a GPU; also how the tests reach this path)
    if (const char *v = getenv("IMK_WS_LIMIT_MB"); v && v[0] && total > (size_t)atoll(v) * (1u << 20)) {
        set_error("unet workspace: %zu bytes for %lld images exceed IMK_WS_LIMIT_MB=%s", total, (long long)n, v);
        return IMK_ENOMEM;
    }
    if (cudaMalloc(&net->ws, total) != cudaSuccess) { cudaGetLastError(); set_error("unet workspace: cudaMalloc(%zu) failed", total); return IMK_ENOMEM; }
    net->dec = reinterpret_cast<uint8_t *>((char *)net->ws + dec_off);
    net->ws_bytes = total;
    for (int l = 0; l < 5; ++l) {
        net->lvl[l].skip = reinterpret_cast<__half *>((char *)net->ws + off[l][0]);
        net->lvl[l].a = reinterpret_cast<__half *>((char *)net->ws + off[l][1]);
        net->lvl[l].b = reinterpret_cast<__half *>((char *)net->ws + off[l][2]);
    }
    net->cap_n = n;
    return IMK_OK;
}

static int launch_conv(imk_unet *net, int layer, const __half *in, const __half *in_lo, __half *out,
                       int64_t n, int h, int w, cudaStream_t stream) {
    return conv_layer_launch(net->conv[layer], layer, net->engine, in, in_lo, out, n, h, w, stream);
}

int conv_layer_launch(const ConvLayer &L, int layer, int engine, const __half *in, const __half *in_lo, __half *out,
                      int64_t n, int h, int w, cudaStream_t stream) {
    if (engine >= 1 && L.w_umma && conv_tc_fits(L, h, w)) {
        IMK_PROFILE(L.ks == 3 ? "conv_tc3" : "conv_tc1", layer, stream);
        return conv_tc_launch(L, in, in_lo, out, nullptr, n, h, w, stream);
    }
    IMK_PROFILE(L.ks == 3 ? "conv_direct3" : "conv_direct1", layer, stream);
    const int tiles_x = (w + kDT - 1) / kDT, tiles_y = (h + kDT - 1) / kDT;
    dim3 grid(tiles_x * tiles_y, L.cout_p / kDCo, (unsigned)n);
    if (L.ks == 3)
        conv_direct_kernel<3><<<grid, 128, 0, stream>>>(in, in_lo, L.w_direct, L.bias, L.has_bn ? L.bn_scale : nullptr,
                                                        L.bn_shift, out, h, w, L.cin_p, L.cout_p, tiles_x);
    else
        conv_direct_kernel<1><<<grid, 128, 0, stream>>>(in, in_lo, L.w_direct, L.bias, L.has_bn ? L.bn_scale : nullptr,
                                                        L.bn_shift, out, h, w, L.cin_p, L.cout_p, tiles_x);
    IMK_LAUNCHED();
    return IMK_OK;
}

int unet_mark_used(imk_unet *net, cudaStream_t stream) {
    if (!net->last_use) IMK_CUDA(cudaEventCreateWithFlags(&net->last_use, cudaEventDisableTiming));
    IMK_CUDA(cudaEventRecord(net->last_use, stream));
    return IMK_OK;
}

int unet_trunk(imk_unet *net, const void *images, int in_dtype, int swap_rb, int64_t n, cudaStream_t stream, const HeadOut *head) {
    int rc = unet_reserve(net, n);
    if (rc) return rc;
    // a model has ONE workspace: whatever still reads it on another stream (the previous call's epilogue) goes first
    if (net->last_use) IMK_CUDA(cudaStreamWaitEvent(stream, net->last_use, 0));
    const imk_unet_desc &d = net->desc;
    const ConvLayer *L = net->conv.data();
    Level *lv = net->lvl;
    const int64_t px0 = n * d.height * d.width;
    const bool fused = net->engine == 2;
    auto pool = [&](int l) -> const __half * {                     // MaxPooling2D of lvl[l].skip -> next level's input
        const int chp = (fused && net->c8 && net->widths[l] <= 8) ? 8 : lv[l].ch_p;   // 8-channel maps keep one plane (BtStage::n8)
        const int64_t items = n * (lv[l].h / 2) * (lv[l].w / 2) * (chp / 8);
        __half *pooled = (l < 3) ? lv[l + 1].b : lv[4].a;
        IMK_PROFILE("maxpool", -1, stream);
        maxpool_kernel<<<grid_1d(items), 256, 0, stream>>>(lv[l].skip, pooled, n, lv[l].h, lv[l].w, chp);
        ++launch_counter();
        return pooled;
    };
    int li = 1;
    const __half *x = nullptr;
    // a fused block whose tile is even-sized also writes the 2x2 max-pooled map (the next level's input)
    auto pooled_of = [&](int l) -> __half * { return (l < 3) ? lv[l + 1].b : lv[4].a; };
    const bool front_u8 = fused && in_dtype == IMK_IN_U8 && net->fb_front_u8.ok;
    if (front_u8 || (fused && net->fb_enc[0].ok)) {
        // input block + encoder block 1 in one kernel: image -> lvl0.skip (+ pooled)
        const FusedBlock &fb = front_u8 ? net->fb_front_u8 : net->fb_enc[0];
        const bool pf = fused_block_can_pool(fb);
        {
            IMK_PROFILE("block_front", 0, stream);
            if ((rc = fused_block_launch(fb, images, nullptr, lv[0].skip, pf ? pooled_of(0) : nullptr, n, swap_rb,
                                         in_dtype == IMK_IN_F32, stream))) return rc;
        }
        li = 3;
        x = pf ? pooled_of(0) : pool(0);
    } else {
        {
            const ConvLayer &c0 = L[0];
            const int grid = grid_1d(px0 * (c0.cout_p / 8));
            IMK_PROFILE("in_conv", 0, stream);
            if (in_dtype == IMK_IN_U8)
                in_conv_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t *)images, d.in_channels, swap_rb, c0.w_f32, c0.cout,
                                                                   c0.bias, c0.bn_scale, c0.bn_shift, lv[0].b, c0.cout_p, px0);
            else
                in_conv_kernel<float><<<grid, 256, 0, stream>>>((const float *)images, d.in_channels, swap_rb, c0.w_f32, c0.cout,
                                                                 c0.bias, c0.bn_scale, c0.bn_shift, lv[0].b, c0.cout_p, px0);
            IMK_LAUNCHED();
        }
        if ((rc = launch_conv(net, li++, lv[0].b, nullptr, lv[0].a, n, lv[0].h, lv[0].w, stream))) return rc;
        if ((rc = launch_conv(net, li++, lv[0].a, nullptr, lv[0].skip, n, lv[0].h, lv[0].w, stream))) return rc;
        x = pool(0);
    }
    IMK_CUDA(cudaGetLastError());
    // encoder blocks 2..4: conv3 -> a ; conv1+BN -> skip ; maxpool -> next level's b
    for (int l = 1; l < 4; ++l) {
        if (fused && net->fb_enc[l].ok) {
            const bool pf = fused_block_can_pool(net->fb_enc[l]);
            {
                IMK_PROFILE("block_enc", li, stream);
                if ((rc = fused_block_launch(net->fb_enc[l], x, nullptr, lv[l].skip, pf ? pooled_of(l) : nullptr, n, 0, 0, stream))) return rc;
            }
            li += 2;
            x = pf ? pooled_of(l) : pool(l);
        } else {
            if ((rc = launch_conv(net, li++, x, nullptr, lv[l].a, n, lv[l].h, lv[l].w, stream))) return rc;
            if ((rc = launch_conv(net, li++, lv[l].a, nullptr, lv[l].skip, n, lv[l].h, lv[l].w, stream))) return rc;
            x = pool(l);
        }
        IMK_CUDA(cudaGetLastError());
    }
    // bottleneck: conv3 (128a -> 256a) -> lvl4.b ; conv1+BN (256a -> 128a) -> lvl4.skip
    if (fused && net->fb_enc[4].ok) {
        IMK_PROFILE("block_enc", li, stream);
        if ((rc = fused_block_launch(net->fb_enc[4], x, nullptr, lv[4].skip, nullptr, n, 0, 0, stream))) return rc;
        li += 2;
    } else {
        if ((rc = launch_conv(net, li++, x, nullptr, lv[4].b, n, lv[4].h, lv[4].w, stream))) return rc;
        if ((rc = launch_conv(net, li++, lv[4].b, nullptr, lv[4].skip, n, lv[4].h, lv[4].w, stream))) return rc;
    }
    x = lv[4].skip;
    // decoder: (up(x) + skip) conv1+BN -> a ; conv3 -> b ; conv1+BN -> a
    for (int l = 3; l >= 0; --l) {
        if (l == 0 && head && unet_has_head(net)) {
            IMK_PROFILE("block_head", li, stream);
            if ((rc = fused_block_launch(net->fb_head, lv[l].skip, x, nullptr, nullptr, n, 0, 0, stream, head))) return rc;
            li += 3;
        } else if (fused && net->fb_dec[l].ok) {
            IMK_PROFILE("block_dec", li, stream);
            if ((rc = fused_block_launch(net->fb_dec[l], lv[l].skip, x, lv[l].a, nullptr, n, 0, 0, stream))) return rc;
            li += 3;
        } else {
            if ((rc = launch_conv(net, li++, lv[l].skip, x, lv[l].a, n, lv[l].h, lv[l].w, stream))) return rc;
            if ((rc = launch_conv(net, li++, lv[l].a, nullptr, lv[l].b, n, lv[l].h, lv[l].w, stream))) return rc;
            if ((rc = launch_conv(net, li++, lv[l].b, nullptr, lv[l].a, n, lv[l].h, lv[l].w, stream))) return rc;
        }
        x = lv[l].a;
    }
    return IMK_OK;       // c9 == lv[0].a
}

// f(KMAX, KFIX, C1FIX): the reference's heads get compile-time shapes, anything else the run-time path
template <typename F>
static int dispatch_head(int K, int c1p, F &&f) {
    using std::integral_constant;
#define IMK_HEAD(KK, CC) if (K == KK && c1p == CC) return f(integral_constant<int, KK>{}, integral_constant<int, KK>{}, integral_constant<int, CC>{})
    IMK_HEAD(1, 8); IMK_HEAD(3, 8); IMK_HEAD(1, 16); IMK_HEAD(1, 32); IMK_HEAD(3, 16); IMK_HEAD(3, 32); IMK_HEAD(9, 16); IMK_HEAD(9, 32); IMK_HEAD(35, 16); IMK_HEAD(35, 32);
#undef IMK_HEAD
    if (K <= 4) return f(integral_constant<int, 4>{}, integral_constant<int, 0>{}, integral_constant<int, 0>{});
    if (K <= 16) return f(integral_constant<int, 16>{}, integral_constant<int, 0>{}, integral_constant<int, 0>{});
    return f(integral_constant<int, 64>{}, integral_constant<int, 0>{}, integral_constant<int, 0>{});
}

static int launch_out_probs(imk_unet *net, int64_t n, float *probs, cudaStream_t stream) {
    const imk_unet_desc &d = net->desc;
    const ConvLayer &Lo = net->conv.back();
    const int64_t px = n * d.height * d.width;
    const int K = d.num_outputmasks, c1p = unet_c9_channels(net);
    const float *w_out = unet_out_weights(net);
    return dispatch_head(K, c1p, [&](auto kmax, auto kfix, auto c1fix) -> int {
        constexpr int KM = decltype(kmax)::value, KF = decltype(kfix)::value, CF = decltype(c1fix)::value;
        size_t smem = (size_t)((K * c1p + K + 3) / 4 * 4) * sizeof(float);
        if constexpr (KF > 0) smem += (size_t)HeadMma<KF, CF>::BFRAG_WORDS * 4 + 8 * (size_t)HeadMma<KF, CF>::WARP_BYTES;
        IMK_CUDA(cudaFuncSetAttribute(out_probs_kernel<KM, KF, CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        IMK_PROFILE("out_probs", 23, stream);
        out_probs_kernel<KM, KF, CF><<<grid_1d(px, 256, 4), 256, smem, stream>>>(net->lvl[0].a, c1p, w_out, Lo.bias, K, d.act_out, probs, px);
        IMK_LAUNCHED();
        return IMK_OK;
    });
}

// ---- building blocks shared with the EvalNet forward (imk_evalnet.cu) ----------------------------------------------
int conv_layer_pack(ConvLayer &L, int ks, int cin, int cout, const float *k /*HWIO*/, const float *b, const float *const bn[4],
                    bool first, std::vector<void *> &owned) {
    int rc;
    L = ConvLayer{};
    L.ks = ks; L.cin = cin; L.cout = cout; L.cin_p = pad_ch(cin); L.cout_p = pad_ch(cout);
    const int64_t wsz = (int64_t)ks * ks * cin * cout;
    if (first) {
        std::vector<float> wf(k, k + wsz);                              // [1][1][c][cout] == [c][cout]
        if ((rc = upload(owned, wf, &L.w_f32))) return rc;
    } else {
        std::vector<__half> wh((size_t)ks * ks * L.cin_p * L.cout_p, __float2half(0.f));
        for (int tap = 0; tap < ks * ks; ++tap)
            for (int ci = 0; ci < cin; ++ci)
                for (int co = 0; co < cout; ++co)
                    wh[((size_t)tap * L.cin_p + ci) * L.cout_p + co] = __float2half_rn(k[((size_t)tap * cin + ci) * cout + co]);
        if ((rc = upload(owned, wh, &L.w_direct))) return rc;
        if (conv_tc_supported(L) && (rc = conv_tc_pack(L, k, owned))) return rc;
    }
    std::vector<float> bias(L.cout_p, 0.f);
    for (int co = 0; co < cout; ++co) bias[co] = b[co];
    if ((rc = upload(owned, bias, &L.bias))) return rc;
    if (bn) {
        std::vector<float> sc(L.cout_p, 0.f), sh(L.cout_p, 0.f);
        for (int ch = 0; ch < cout; ++ch) {
            sc[ch] = bn[0][ch] / sqrtf(bn[3][ch] + kBnEps);
            sh[ch] = bn[1][ch] - bn[2][ch] * sc[ch];
        }
        if ((rc = upload(owned, sc, &L.bn_scale))) return rc;
        if ((rc = upload(owned, sh, &L.bn_shift))) return rc;
        L.has_bn = true;
    }
    return IMK_OK;
}

int in_conv_launch(const ConvLayer &L, const void *img, int in_dtype, int c, int swap_rb, int normalize, __half *out, int64_t px,
                   cudaStream_t stream) {
    const int grid = grid_1d(px * (L.cout_p / 8));
    IMK_PROFILE("in_conv", 0, stream);
    if (in_dtype == IMK_IN_U8)
        in_conv_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t *)img, c, swap_rb, L.w_f32, L.cout, L.bias, L.bn_scale, L.bn_shift,
                                                           out, L.cout_p, px, normalize);
    else
        in_conv_kernel<float><<<grid, 256, 0, stream>>>((const float *)img, c, swap_rb, L.w_f32, L.cout, L.bias, L.bn_scale, L.bn_shift,
                                                         out, L.cout_p, px, normalize);
    IMK_LAUNCHED();
    return IMK_OK;
}

int maxpool_launch(const __half *in, __half *out, int64_t n, int h, int w, int cp, cudaStream_t stream) {
    const int64_t items = n * (h / 2) * (w / 2) * (cp / 8);
    IMK_PROFILE("maxpool", -1, stream);
    maxpool_kernel<<<grid_1d(items), 256, 0, stream>>>(in, out, n, h, w, cp);
    IMK_LAUNCHED();
    return IMK_OK;
}

}  // namespace imk

using namespace imk;

extern "C" int imk_unet_create(const imk_unet_desc *desc, const float *const *weights_host,
                               const int64_t *weight_sizes, int n_weights, imk_unet_t **out) {
    IMK_REQUIRE(desc && weights_host && weight_sizes && out, "imk_unet_create: NULL argument");
    const imk_unet_desc &d = *desc;
    IMK_REQUIRE(d.height > 0 && d.width > 0 && d.height % 16 == 0 && d.width % 16 == 0,
                "imk_unet_create: height/width must be positive multiples of 16 (4 poolings), got %dx%d", d.height, d.width);
    IMK_REQUIRE(d.in_channels >= 1 && d.in_channels <= 4, "imk_unet_create: in_channels=%d outside 1..4", d.in_channels);
    IMK_REQUIRE(d.num_outputmasks >= 1 && d.num_outputmasks <= 64, "imk_unet_create: num_outputmasks=%d outside 1..64", d.num_outputmasks);
    IMK_REQUIRE(d.ks == 1 || d.ks == 3, "imk_unet_create: ks=%d (1 or 3)", d.ks);
    IMK_REQUIRE(d.act_out == IMK_ACT_SIGMOID || d.act_out == IMK_ACT_SOFTMAX, "imk_unet_create: act_out=%d", d.act_out);
    if (!imk_device_available()) { set_error("imk_unet_create: no CUDA device (there is no CPU fallback)"); return IMK_ECUDA; }
    int f[5];
    const int base[5] = {16, 32, 64, 128, 256};
    for (int i = 0; i < 5; ++i) {
        f[i] = (int)(base[i] * (double)d.alpha);      // int(k * alpha), unet.py:49-61
        IMK_REQUIRE(f[i] >= 1, "imk_unet_create: alpha=%g gives an empty layer", (double)d.alpha);
    }
    IMK_REQUIRE(pad_ch(f[0]) <= 64, "imk_unet_create: alpha=%g too wide for the output kernels (int(16*alpha) <= 64)", (double)d.alpha);
    const std::vector<PlanItem> plan = make_plan(d, f);
    int expect = 0;
    for (const PlanItem &it : plan) expect += it.is_conv ? 2 : 4;
    IMK_REQUIRE(n_weights == expect, "imk_unet_create: expected %d weight arrays (Keras get_weights order), got %d", expect, n_weights);

    imk_unet *net = new imk_unet();
    net->desc = d;
    for (int i = 0; i < 5; ++i) net->widths[i] = f[i];
    int wi = 0, rc = IMK_OK;
    auto fail = [&](int code) { imk_unet_destroy(net); return code; };
    std::vector<ConvHost> host;                                   // per conv, for the block-fused packs
    std::vector<std::vector<float>> host_bn;                      // keeps the folded BN vectors alive
    for (size_t pi = 0; pi < plan.size(); ++pi) {
        const PlanItem &it = plan[pi];
        if (it.is_conv) {
            ConvLayer L;
            L.ks = it.ks; L.cin = it.cin; L.cout = it.cout;
            L.cin_p = pad_ch(it.cin); L.cout_p = pad_ch(it.cout);
            const int64_t wsz = (int64_t)it.ks * it.ks * it.cin * it.cout;
            if (weight_sizes[wi] != wsz || weight_sizes[wi + 1] != it.cout) {
                set_error("imk_unet_create: weight %d: expected kernel %dx%dx%dx%d (+bias %d), got sizes %lld, %lld", wi, it.ks, it.ks,
                          it.cin, it.cout, it.cout, (long long)weight_sizes[wi], (long long)weight_sizes[wi + 1]);
                return fail(IMK_EINVAL);
            }
            const float *k = weights_host[wi], *b = weights_host[wi + 1];
            wi += 2;
            net->n_params += wsz + it.cout;
            const bool first = (pi == 0), last = (pi + 1 == plan.size());
            if (first) {
                std::vector<float> wf(k, k + wsz);                              // [1][1][c][cout] == [c][cout]
                if ((rc = upload(net->owned, wf, &L.w_f32))) return fail(rc);
            } else if (last) {
                std::vector<float> wf((size_t)it.cout * L.cin_p, 0.f);          // [K][c1p]
                for (int ci = 0; ci < it.cin; ++ci)
                    for (int co = 0; co < it.cout; ++co) wf[(size_t)co * L.cin_p + ci] = k[(size_t)ci * it.cout + co];
                if ((rc = upload(net->owned, wf, &L.w_f32))) return fail(rc);
                if (it.cin <= 8) {                                              // c9 as ONE 8-channel plane (imk_unet::c8): [K][8]
                    std::vector<float> w8((size_t)it.cout * 8, 0.f);
                    for (int ci = 0; ci < it.cin; ++ci)
                        for (int co = 0; co < it.cout; ++co) w8[(size_t)co * 8 + ci] = k[(size_t)ci * it.cout + co];
                    if ((rc = upload(net->owned, w8, &net->w_out8))) return fail(rc);
                }
            } else {
                std::vector<__half> wh((size_t)it.ks * it.ks * L.cin_p * L.cout_p, __float2half(0.f));
                for (int tap = 0; tap < it.ks * it.ks; ++tap)
                    for (int ci = 0; ci < it.cin; ++ci)
                        for (int co = 0; co < it.cout; ++co)
                            wh[((size_t)tap * L.cin_p + ci) * L.cout_p + co] =
                                __float2half_rn(k[((size_t)tap * it.cin + ci) * it.cout + co]);
                if ((rc = upload(net->owned, wh, &L.w_direct))) return fail(rc);
                if (conv_tc_supported(L) && (rc = conv_tc_pack(L, k, net->owned))) return fail(rc);
            }
            std::vector<float> bias(last ? it.cout : L.cout_p, 0.f);
            for (int co = 0; co < it.cout; ++co) bias[co] = b[co];
            if ((rc = upload(net->owned, bias, &L.bias))) return fail(rc);
            net->conv.push_back(L);
            host.push_back(ConvHost{k, b, nullptr, nullptr, it.ks, it.cin, it.cout});
        } else {
            ConvLayer &L = net->conv.back();
            for (int j = 0; j < 4; ++j)
                if (weight_sizes[wi + j] != it.cin) {
                    set_error("imk_unet_create: weight %d: BatchNormalization vector of %d expected, got %lld", wi + j, it.cin,
                              (long long)weight_sizes[wi + j]);
                    return fail(IMK_EINVAL);
                }
            const float *g = weights_host[wi], *be = weights_host[wi + 1], *mu = weights_host[wi + 2], *var = weights_host[wi + 3];
            wi += 4;
            net->n_params += 4 * (int64_t)it.cin;
            std::vector<float> sc(L.cout_p, 0.f), sh(L.cout_p, 0.f);
            for (int ch = 0; ch < it.cin; ++ch) {
                sc[ch] = g[ch] / sqrtf(var[ch] + kBnEps);
                sh[ch] = be[ch] - mu[ch] * sc[ch];
            }
            if ((rc = upload(net->owned, sc, &L.bn_scale))) return fail(rc);
            if ((rc = upload(net->owned, sh, &L.bn_shift))) return fail(rc);
            L.has_bn = true;
            host_bn.push_back(sc); host_bn.push_back(sh);
        }
    }
    {   // block-fused packs (imk_block_tc.cu); a block that does not fit keeps ok == false and runs layer-wise
        size_t bi = 0;
        for (size_t ci = 0; ci < host.size(); ++ci)
            if (net->conv[ci].has_bn) { host[ci].bn_scale = host_bn[bi].data(); host[ci].bn_shift = host_bn[bi + 1].data(); bi += 2; }
        if (d.ks == 3) {
            auto build_all = [&](bool allow8) -> int {
                int r;
                if ((r = fused_block_build(net->fb_enc[0], 0, &host[0], d.height, d.width, d.in_channels, net->owned, 0, allow8))) return r;
                if (!getenv("IMK_BT_NO_FRONT_U8") && (r = fused_block_build(net->fb_front_u8, 3, &host[0], d.height, d.width, d.in_channels, net->owned, 0, allow8))) return r;
                for (int l = 1; l < 5; ++l)
                    if ((r = fused_block_build(net->fb_enc[l], 1, &host[1 + 2 * l], d.height >> l, d.width >> l, 0, net->owned, 0, allow8))) return r;
                for (int l = 0; l < 4; ++l)
                    if ((r = fused_block_build(net->fb_dec[l], 2, &host[11 + 3 * (3 - l)], d.height >> l, d.width >> l, 0, net->owned, 0, allow8))) return r;
                return IMK_OK;
            };
            // Networks with int(16 * alpha) <= 8 (the reference's alpha = 0.5 ISIC models): maps of <= 8 channels live as ONE
            // 16-byte plane per pixel in HBM and in the operand buffers (BtStage::kin8 / n8).  Every producer and consumer
            // of such a map must be a fused block, so the mode needs all nine blocks (and their pooled outputs); otherwise
            // everything is built with the 16-channel padding the layer-wise engines share.
            bool want8 = f[0] <= 8 && net->w_out8;
            if (const char *v = getenv("IMK_BT_NO_C8"); v && v[0] == '1') want8 = false;
            if ((rc = build_all(want8))) return fail(rc);
            if (want8) {
                // enc[l] reads widths[l-1] channels and writes widths[l] (+ pooled); dec[l] reads widths[l], writes widths[max(l-1, 0)]
                bool all = net->fb_enc[0].ok && fused_block_can_pool(net->fb_enc[0]);
                for (int l = 1; l < 5; ++l)
                    if (f[l - 1] <= 8) all = all && net->fb_enc[l].ok && (l == 4 || f[l] > 8 || fused_block_can_pool(net->fb_enc[l]));
                for (int l = 0; l < 4; ++l)
                    if (f[l > 0 ? l - 1 : 0] <= 8) all = all && net->fb_dec[l].ok;
                if (net->fb_front_u8.ok) all = all && fused_block_can_pool(net->fb_front_u8);
                if (all) net->c8 = true;
                else if ((rc = build_all(false))) return fail(rc);
            }
            // Head-in-epilogue variant of the level-0 decoder (K <= 3): opt-in.  Measured (r2m, B200, 512 images): the block
            // kernel is bound by the instruction issue of its epilogue warps, so the K * C0 FMAs + activation per pixel cost
            // more there (HeLa 1107 -> 1950 us, ISIC 1150 -> 1750 us per model) than the separate ensemble_im kernel they
            // replace (417 / 264 us per model); an MMA head stage (S4 + a fourth epilogue pass, r2c-r2h) measured the same.
            if (const char *v = getenv("IMK_BT_HEAD"); v && v[0] == '1' && !net->c8)
                if ((rc = fused_block_build(net->fb_head, 4, &host[20], d.height, d.width, 0, net->owned, d.act_out))) return fail(rc);
        }
    }
    *out = net;
    return IMK_OK;
}

extern "C" void imk_unet_destroy(imk_unet_t *net) {
    if (!net) return;
    for (void *p : net->owned) cudaFree(p);
    if (net->ws) cudaFree(net->ws);
    if (net->last_use) cudaEventDestroy(net->last_use);
    if (net->stage_in) cudaFree(net->stage_in);
    if (net->stage_probs) cudaFree(net->stage_probs);
    delete net;
}

extern "C" int imk_unet_param_count(const imk_unet_t *net, int64_t *count) {
    IMK_REQUIRE(net && count, "imk_unet_param_count: NULL argument");
    *count = net->n_params;
    return IMK_OK;
}

extern "C" int imk_unet_set_engine(imk_unet_t *net, int engine) {
    IMK_REQUIRE(net && engine >= 0 && engine <= 2, "imk_unet_set_engine: engine must be 0 (direct), 1 (tcgen05 layer-wise) or 2 (tcgen05 block-fused)");
    net->engine = engine;
    return IMK_OK;
}

extern "C" int imk_unet_set_swap_rb(imk_unet_t *net, int swap_rb) {
    IMK_REQUIRE(net, "imk_unet_set_swap_rb: NULL handle");
    net->desc.swap_rb = swap_rb ? 1 : 0;
    return IMK_OK;
}

extern "C" int imk_unet_forward(imk_unet_t *net, const void *images_dev, int in_dtype, int64_t N,
                                float *probs_dev, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    IMK_REQUIRE(net && images_dev && probs_dev, "imk_unet_forward: NULL argument");
    IMK_REQUIRE(in_dtype == IMK_IN_U8 || in_dtype == IMK_IN_F32, "imk_unet_forward: in_dtype=%d", in_dtype);
    IMK_REQUIRE(N >= 0, "imk_unet_forward: N=%lld", (long long)N);
    const imk_unet_desc &d = net->desc;
    const size_t in_px_bytes = (size_t)d.in_channels * (in_dtype == IMK_IN_U8 ? 1 : 4);
    const int64_t HW = (int64_t)d.height * d.width;
    for (int64_t n0 = 0; n0 < N; n0 += kMaxChunk) {
        const int64_t n = (N - n0 < kMaxChunk) ? N - n0 : kMaxChunk;
        float *probs_c = probs_dev + n0 * HW * d.num_outputmasks;
        const HeadOut ho{0, 0.f, 0.f, 0, probs_c, nullptr};
        const bool head = unet_has_head(net);
        int rc = unet_trunk(net, (const char *)images_dev + n0 * HW * in_px_bytes, in_dtype, d.swap_rb, n, stream, head ? &ho : nullptr);
        if (rc) return rc;
        if (!head && (rc = launch_out_probs(net, n, probs_c, stream))) return rc;
        if ((rc = unet_mark_used(net, stream))) return rc;
    }
    return IMK_OK;
}

extern "C" int imk_unet_predict_host(imk_unet_t *net, const void *images_host, int in_dtype, int64_t N, float *probs_host) {
    IMK_REQUIRE(net && images_host && probs_host, "imk_unet_predict_host: NULL argument");
    IMK_REQUIRE(in_dtype == IMK_IN_U8 || in_dtype == IMK_IN_F32, "imk_unet_predict_host: in_dtype=%d", in_dtype);
    IMK_REQUIRE(N >= 0, "imk_unet_predict_host: N=%lld", (long long)N);
    const imk_unet_desc &d = net->desc;
    const int64_t HW = (int64_t)d.height * d.width;
    const size_t in_img = (size_t)HW * d.in_channels * (in_dtype == IMK_IN_U8 ? 1 : 4);
    const size_t out_img = (size_t)HW * d.num_outputmasks * sizeof(float);
    const int64_t chunk = N < kMaxChunk ? (N > 0 ? N : 1) : kMaxChunk;
    if (net->stage_in_bytes < in_img * chunk) {
        if (net->stage_in) cudaFree(net->stage_in);
        net->stage_in = nullptr; net->stage_in_bytes = 0;
        if (cudaMalloc(&net->stage_in, in_img * chunk) != cudaSuccess) { set_error("predict: cudaMalloc failed"); return IMK_ENOMEM; }
        net->stage_in_bytes = in_img * chunk;
    }
    if (net->stage_probs_bytes < out_img * chunk) {
        if (net->stage_probs) cudaFree(net->stage_probs);
        net->stage_probs = nullptr; net->stage_probs_bytes = 0;
        if (cudaMalloc(&net->stage_probs, out_img * chunk) != cudaSuccess) { set_error("predict: cudaMalloc failed"); return IMK_ENOMEM; }
        net->stage_probs_bytes = out_img * chunk;
    }
    for (int64_t n0 = 0; n0 < N; n0 += chunk) {
        const int64_t n = (N - n0 < chunk) ? N - n0 : chunk;
        IMK_CUDA(cudaMemcpyAsync(net->stage_in, (const char *)images_host + n0 * in_img, in_img * n, cudaMemcpyHostToDevice, 0));
        int rc = imk_unet_forward(net, net->stage_in, in_dtype, n, net->stage_probs, nullptr);
        if (rc) return rc;
        IMK_CUDA(cudaMemcpyAsync((char *)probs_host + n0 * out_img, net->stage_probs, out_img * n, cudaMemcpyDeviceToHost, 0));
    }
    IMK_CUDA(cudaStreamSynchronize(0));
    return IMK_OK;
}

// ---------------------------------------------------------------------------------------
//  fused ensemble calls
// ---------------------------------------------------------------------------------------
namespace imk {
static thread_local unsigned long long *g_presence2 = nullptr;
static thread_local size_t g_presence2_cap = 0;

// The trunks of different models are independent until the ensemble epilogue: they are issued round-robin to the
// caller's stream and a few auxiliary streams, so that the tail of one model's persistent kernel (SMs whose CTA has
// run out of tiles) is filled by the next kernel of another model.  IMK_STREAMS=1 restores the sequential order.
constexpr int kMaxAux = 3;
struct AuxStreams {
    int device = -1, n = 0;
    cudaStream_t s[kMaxAux] = {};
    cudaEvent_t fork = nullptr, join[kMaxAux] = {};
};
static thread_local AuxStreams g_aux;
static int aux_streams(int want, AuxStreams **out) {
    int dev = 0;
    IMK_CUDA(cudaGetDevice(&dev));
    AuxStreams &A = g_aux;
    if (A.device != dev) {                                   // (re)create on this device; the old ones die with their context
        A = AuxStreams{};
        A.device = dev;
        IMK_CUDA(cudaEventCreateWithFlags(&A.fork, cudaEventDisableTiming));
    }
    while (A.n < want && A.n < kMaxAux) {
        IMK_CUDA(cudaStreamCreateWithFlags(&A.s[A.n], cudaStreamNonBlocking));
        IMK_CUDA(cudaEventCreateWithFlags(&A.join[A.n], cudaEventDisableTiming));
        ++A.n;
    }
    *out = &A;
    return IMK_OK;
}

static int check_ensemble(imk_unet_t *const *nets, int M, const char *who) {
    IMK_REQUIRE(nets && M >= 1 && M <= IMK_MAX_MODELS, "%s: M=%d outside 1..%d", who, M, IMK_MAX_MODELS);
    for (int m = 0; m < M; ++m) {
        IMK_REQUIRE(nets[m], "%s: nets[%d] is NULL", who, m);
        const imk_unet_desc &a = nets[0]->desc, &b = nets[m]->desc;
        IMK_REQUIRE(a.height == b.height && a.width == b.width && a.in_channels == b.in_channels &&
                        a.num_outputmasks == b.num_outputmasks && a.act_out == b.act_out,
                    "%s: model %d has a different input / output signature than model 0", who, m);
    }
    return IMK_OK;
}

// largest d with RN(1 / d) >= thr (> thr when strict): the sigmoid decision becomes one compare (0 = not applicable)
static float sigmoid_dstar(float thr, int strict) {
    if (!(thr > 0.f && thr < 1.f)) return 0.f;
    auto fires = [&](float dd) { const volatile float q = 1.0f / dd; return strict ? q > thr : q >= thr; };
    float d0 = 1.0f / thr;
    for (int it = 0; it < 64 && fires(nextafterf(d0, INFINITY)); ++it) d0 = nextafterf(d0, INFINITY);
    for (int it = 0; it < 64 && !fires(d0); ++it) d0 = nextafterf(d0, 0.f);
    return (fires(d0) && !fires(nextafterf(d0, INFINITY))) ? d0 : 0.f;
}

template <int KMAX, bool MC, int KFIX, int C1FIX>
static int launch_ens(const EnsPtrs &ens, int M, int c1p, int K, int act, float thr, int strict, int64_t total_px, int64_t HW,
                      int64_t N, int64_t plane_stride, const uint8_t *img, int c, int block_in, int block_out, uint8_t *img_out, uint8_t *labels,
                      uint8_t *im, int64_t *im_size, int64_t *pred_size, unsigned long long *presence, cudaStream_t stream) {
    size_t smem = (size_t)M * ((K * c1p + K + 3) / 4 * 4) * sizeof(float) + 8 * 32;            // weights + bias per model, class-id bytes per warp
    if constexpr (KFIX > 0) smem += (size_t)M * HeadMma<KFIX, C1FIX>::BFRAG_WORDS * 4 + 8 * (size_t)HeadMma<KFIX, C1FIX>::WARP_BYTES;
    IMK_CUDA(cudaFuncSetAttribute(ensemble_im_kernel<KMAX, MC, KFIX, C1FIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = grid_1d(total_px, 256, (KFIX > 0 && C1FIX == 8) ? 8 : 4);    // the 8-channel head is light on registers: more CTAs in flight
    const float dstar = (!MC && KFIX > 0 && act == IMK_ACT_SIGMOID) ? sigmoid_dstar(thr, strict) : 0.f;
    IMK_PROFILE("ensemble_im", -1, stream);
    ensemble_im_kernel<KMAX, MC, KFIX, C1FIX><<<grid, 256, smem, stream>>>(ens, M, c1p, K, act, thr, strict, dstar, total_px, HW, N, plane_stride, img, c,
                                                               block_in, block_out, img_out, labels, im, im_size, pred_size, presence);
    IMK_LAUNCHED();
    return IMK_OK;
}
}  // namespace imk

static int ensemble_run(imk_unet_t *const *nets, int M, bool multiclass, const uint8_t *images_dev, int64_t N, int swap_rb,
                        float thr, int strict, int block_in, int block_out,
                        uint8_t *img_out, uint8_t *labels, uint8_t *im, int64_t *im_size, int64_t *pred_size,
                        uint8_t *lists_equal, cudaStream_t stream, const char *who) {
    int rc = check_ensemble(nets, M, who);
    if (rc) return rc;
    IMK_REQUIRE(images_dev && labels && im && im_size, "%s: NULL images/labels/im/im_size", who);
    IMK_REQUIRE(N >= 0, "%s: N=%lld", who, (long long)N);
    const imk_unet_desc &d = nets[0]->desc;
    const int K = d.num_outputmasks;
    if (!multiclass) IMK_REQUIRE(K == 1 || K == 3, "%s: binary IM needs K = 1 (ISIC) or 3 (HeLa), model has %d", who, K);
    IMK_REQUIRE(!lists_equal || K <= 64, "%s: lists_equal needs K <= 64", who);
    const int c1p = unet_c9_channels(nets[0]);
    // head path: every model's level-0 decoder kernel also runs the output layer and leaves one decision byte per pixel
    bool use_head = true;
    for (int m = 0; m < M; ++m) use_head = use_head && unet_has_head(nets[m]);
    for (int m = 0; m < M && !use_head; ++m)
        IMK_REQUIRE(unet_c9_channels(nets[m]) == c1p, "%s: models with different int(16*alpha) padding (or engines) cannot share the fused epilogue", who);
    if (N == 0) return IMK_OK;
    const int64_t HW = (int64_t)d.height * d.width;
    IMK_CUDA(cudaMemsetAsync(im_size, 0, sizeof(int64_t) * N, stream));
    if (pred_size && !multiclass) IMK_CUDA(cudaMemsetAsync(pred_size, 0, sizeof(int64_t) * N * K, stream));
    unsigned long long *presence = nullptr;
    if (multiclass && lists_equal) {
        const size_t need = sizeof(unsigned long long) * (size_t)N * M;
        if (need > g_presence2_cap) {
            if (g_presence2) cudaFree(g_presence2);
            g_presence2 = nullptr; g_presence2_cap = 0;
            if (cudaMalloc(&g_presence2, need) != cudaSuccess) { cudaGetLastError(); set_error("%s: cudaMalloc(%zu) failed", who, need); return IMK_ENOMEM; }
            g_presence2_cap = need;
        }
        presence = g_presence2;
        IMK_CUDA(cudaMemsetAsync(presence, 0, need, stream));
    }
    int n_streams = 3;                                            // measured (r3k, ISIC M=5): 2 -> 39.5k, 3 -> 40.6k, 4 -> 40.2k img/s
    if (const char *v = getenv("IMK_STREAMS"); v && v[0]) n_streams = atoi(v);
    n_streams = std::max(1, std::min(std::min(n_streams, M), kMaxAux + 1));
    if (profiling_active()) n_streams = 1;                       // per-kernel times are only meaningful without overlap
    AuxStreams *aux = nullptr;
    if (n_streams > 1 && (rc = aux_streams(n_streams - 1, &aux))) return rc;
    const float dstar = (!multiclass && d.act_out == IMK_ACT_SIGMOID) ? sigmoid_dstar(thr, strict) : 0.f;
    for (int64_t n0 = 0; n0 < N; n0 += kMaxChunk) {
        const int64_t n = (N - n0 < kMaxChunk) ? N - n0 : kMaxChunk;
        EnsPtrs ens{};
        DecPtrs decs{};
        // fork: everything already in the caller's stream (uploads, the previous chunk's epilogue that still reads
        // the workspaces) precedes the auxiliary streams' work
        if (aux) {
            IMK_CUDA(cudaEventRecord(aux->fork, stream));
            for (int a = 0; a < n_streams - 1; ++a) IMK_CUDA(cudaStreamWaitEvent(aux->s[a], aux->fork, 0));
        }
        for (int m = 0; m < M; ++m) {
            const int lane = m % n_streams;
            cudaStream_t sm = lane == 0 ? stream : aux->s[lane - 1];
            if ((rc = unet_reserve(nets[m], n))) return rc;       // nets[m]->dec must exist before the HeadOut is formed
            const HeadOut ho{multiclass ? 2 : 1, thr, dstar, strict, nullptr, nets[m]->dec};
            if ((rc = unet_trunk(nets[m], images_dev + n0 * HW * d.in_channels, IMK_IN_U8, swap_rb ? 1 : 0, n, sm, use_head ? &ho : nullptr))) return rc;
            ens.c9[m] = nets[m]->lvl[0].a;
            ens.w[m] = unet_out_weights(nets[m]);
            ens.b[m] = nets[m]->conv.back().bias;
            decs.d[m] = nets[m]->dec;
        }
        if (aux)                                                 // join before the epilogue
            for (int a = 0; a < n_streams - 1; ++a) {
                IMK_CUDA(cudaEventRecord(aux->join[a], aux->s[a]));
                IMK_CUDA(cudaStreamWaitEvent(stream, aux->join[a], 0));
            }
        // chunk view: pixel arrays offset by n0*HW, per-image statistics by n0; label planes stay N*HW apart
        const uint8_t *img_c = images_dev + n0 * HW * d.in_channels;
        uint8_t *img_out_c = img_out ? img_out + n0 * HW * d.in_channels : nullptr;
        uint8_t *im_c = im + n0 * HW;
        int64_t *im_size_c = im_size + n0;
        const int64_t total_px = n * HW;
        if (use_head) {
            IMK_REQUIRE(total_px < 0x7fffffffLL, "%s: chunk of %lld pixels", who, (long long)total_px);
            const int grid = grid_1d(total_px / 16, 256, 8);
            unsigned long long *pres_c = presence ? presence + n0 * M : nullptr;
            int64_t *pred_c = (pred_size && !multiclass) ? pred_size + n0 : nullptr;
            IMK_PROFILE("ensemble_votes", -1, stream);
            if (multiclass)
                ensemble_votes_kernel<2><<<grid, 256, 0, stream>>>(decs, M, total_px, HW, N, N * HW, img_c, d.in_channels, block_in, block_out,
                                                                   img_out_c, labels + n0 * HW, im_c, im_size_c, nullptr, pres_c);
            else if (K == 1)
                ensemble_votes_kernel<0><<<grid, 256, 0, stream>>>(decs, M, total_px, HW, N, N * HW, img_c, d.in_channels, block_in, block_out,
                                                                   img_out_c, labels + n0 * HW, im_c, im_size_c, pred_c, nullptr);
            else
                ensemble_votes_kernel<1><<<grid, 256, 0, stream>>>(decs, M, total_px, HW, N, N * HW, img_c, d.in_channels, block_in, block_out,
                                                                   img_out_c, labels + n0 * HW, im_c, im_size_c, pred_c, nullptr);
            IMK_LAUNCHED();
        } else {
            rc = dispatch_head(K, c1p, [&](auto kmax, auto kfix, auto c1fix) -> int {
                constexpr int KM = decltype(kmax)::value, KF = decltype(kfix)::value, CF = decltype(c1fix)::value;
                if (multiclass)
                    return launch_ens<KM, true, KF, CF>(ens, M, c1p, K, d.act_out, 0.f, 1, total_px, HW, N, N * HW, img_c, d.in_channels,
                                                        block_in, block_out, img_out_c, labels + n0 * HW, im_c, im_size_c,
                                                        nullptr, presence ? presence + n0 * M : nullptr, stream);
                if constexpr (KM <= 4)                                // binary IM: K = 1 (ISIC) or 3 (HeLa), checked above
                    return launch_ens<KM < 3 ? (KM == 1 ? 1 : 4) : KM, false, KF, CF>(ens, M, c1p, K, d.act_out, thr, strict, total_px, HW, N, N * HW, img_c,
                                                         d.in_channels, block_in, block_out, img_out_c, labels + n0 * HW, im_c, im_size_c,
                                                         pred_size ? pred_size + n0 : nullptr, nullptr, stream);
                set_error("%s: binary IM with K = %d", who, K);
                return IMK_EINVAL;
            });
            if (rc) return rc;
        }
        for (int m = 0; m < M; ++m)
            if ((rc = unet_mark_used(nets[m], stream))) return rc;
    }
    if (multiclass && lists_equal) {
        lists_equal_kernel2<<<(int)((N + 255) / 256), 256, 0, stream>>>(presence, M, N, lists_equal);
        IMK_LAUNCHED();
    }
    return IMK_OK;
}

extern "C" int imk_ensemble_im_binary(imk_unet_t *const *nets, int M, const uint8_t *images_dev, int64_t N, int swap_rb,
                                      float thr, int strict_gt, int block_in, int block_out,
                                      uint8_t *img_out_dev, uint8_t *labels_dev, uint8_t *im_dev,
                                      int64_t *im_size_dev, int64_t *pred_size_dev, void *stream) {
    return ensemble_run(nets, M, false, images_dev, N, swap_rb, thr, strict_gt, block_in, block_out, img_out_dev, labels_dev, im_dev,
                        im_size_dev, pred_size_dev, nullptr, (cudaStream_t)stream, "imk_ensemble_im_binary");
}

extern "C" int imk_ensemble_im_multiclass(imk_unet_t *const *nets, int M, const uint8_t *images_dev, int64_t N, int swap_rb,
                                          int block_in, int block_out,
                                          uint8_t *img_out_dev, uint8_t *label_dev, uint8_t *im_dev,
                                          int64_t *im_size_dev, uint8_t *lists_equal_dev, void *stream) {
    return ensemble_run(nets, M, true, images_dev, N, swap_rb, 0.f, 1, block_in, block_out, img_out_dev, label_dev, im_dev,
                        im_size_dev, nullptr, lists_equal_dev, (cudaStream_t)stream, "imk_ensemble_im_multiclass");
}

// Device-side pieces of the Inconsistency-Mask arithmetic that are shared by the
// standalone IM kernels (imk_im.cu) and the fused ensemble epilogue (imk_unet.cu).
// Semantics: SURVEY.md appendix B, i.e. reference functions.py:3104-3238 and the
// blanking of functions.py:2867-2874 / 2968-2974 / 3054-3061.
#pragma once
#include "imk_common.cuh"

namespace imk {

// functions.py:3157 uses `>`, functions.py:3187-3189 use `>=`.  NaN is false for both.
__device__ __forceinline__ uint32_t decide(float p, float thr, bool strict) {
    return strict ? (p > thr) : (p >= thr);
}

// sigmoid / softmax of one pixel's logits in place -- the ONE definition shared by every path
template <int KMAX>
__device__ __forceinline__ void pixel_activation(float (&p)[KMAX], int K, int act) {
    if (act == IMK_ACT_SIGMOID) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) p[k] = __fdiv_rn(1.0f, __fadd_rn(1.0f, __expf(-p[k])));
    } else {
        float mx = p[0];
#pragma unroll
        for (int k = 1; k < KMAX; ++k) if (k < K) mx = fmaxf(mx, p[k]);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) { p[k] = __expf(__fsub_rn(p[k], mx)); sum = __fadd_rn(sum, p[k]); }
#pragma unroll
        for (int k = 0; k < KMAX; ++k) if (k < K) p[k] = __fdiv_rn(p[k], sum);
    }
}

// np.argmax (functions.py:3225): first index of the maximum, NaN is the maximum (the first NaN wins).
// Floats are mapped to unsigned keys whose integer order is the float order with -0 == +0 and every NaN on top,
// so that one strict integer compare per element implements the whole rule (about 8 instructions per element).
__device__ __forceinline__ uint32_t order_key(float v) {
    v = v + 0.0f;                                                   // -0 -> +0 (np.argmax: equal, the first wins)
    const uint32_t b = __float_as_uint(v);
    const uint32_t key = b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
    return (v != v) ? 0xFFFFFFFFu : key;
}

// argmax_step folds candidate (v, k) into the running (best, arg); every index folded so far must be smaller than k.
__device__ __forceinline__ void argmax_step(float v, int k, float &best, int &arg) {
    // a NaN `best` can never be displaced; a NaN v displaces any non-NaN best
    if (!(best != best) && (v > best || v != v)) { best = v; arg = k; }
}

// exact rule for any input (NaN / signed zeros) through the order keys; K may be a run-time value
template <int KT>
__device__ __forceinline__ int argmax_row_keys(const float *__restrict__ row, int K) {
    uint32_t best = order_key(row[0]);
    int arg = 0;
    if constexpr (KT > 0) {
#pragma unroll
        for (int k = 1; k < KT; ++k) {
            const uint32_t key = order_key(row[k]);
            if (key > best) { best = key; arg = k; }
        }
    } else {
#pragma unroll 4
        for (int k = 1; k < K; ++k) {
            const uint32_t key = order_key(row[k]);
            if (key > best) { best = key; arg = k; }
        }
    }
    return arg;
}

// Fast path for a compile-time class count: a tree of independent merges (FSETP + FSEL + SEL each, ties keep the
// lower index: np.argmax replaces the running maximum only when `!(v <= max)`), plus the row sum, which is NaN
// exactly when the row holds a NaN (or both infinities).  Only such rows take the exact order-key scan, so the
// common case costs ~4 instructions per element with K/2-way instruction-level parallelism.
template <int LO, int N>
__device__ __forceinline__ void argmax_tree(const float *__restrict__ row, float &best, int &arg, float &sum) {
    if constexpr (N == 1) {
        best = row[LO]; arg = LO; sum = best;
    } else {
        float bl, br, sl, sr;
        int al, ar;
        argmax_tree<LO, N / 2>(row, bl, al, sl);
        argmax_tree<LO + N / 2, N - N / 2>(row, br, ar, sr);
        const bool gt = br > bl;
        best = gt ? br : bl; arg = gt ? ar : al; sum = sl + sr;
    }
}

// rows wider than 8 classes are folded chunk by chunk (8-wide trees merged left to right) so that only one
// chunk of values is live at a time
template <int KT>
__device__ __forceinline__ int argmax_row_t(const float *__restrict__ row, int K) {
    if constexpr (KT > 0) {
        constexpr int CH = 8;
        float best, sum;
        int arg;
        argmax_tree<0, (KT < CH ? KT : CH)>(row, best, arg, sum);
#pragma unroll
        for (int k0 = CH; k0 < KT; k0 += CH) {
            float b2, s2;
            int a2;
            if (k0 + CH <= KT) argmax_tree<0, CH>(row + k0, b2, a2, s2);
            else argmax_tree<0, (KT % CH == 0 ? CH : KT % CH)>(row + k0, b2, a2, s2);
            const bool gt = b2 > best;
            best = gt ? b2 : best; arg = gt ? a2 + k0 : arg; sum += s2;
        }
        if (sum != sum) arg = argmax_row_keys<KT>(row, K);
        return arg;
    } else {
        return argmax_row_keys<0>(row, K);
    }
}

__device__ __forceinline__ int argmax_row(const float *__restrict__ row, int K) {
    switch (K) {                                                    // the reference's class counts (config.ini:64, 85) fully unrolled
        case 9: return argmax_row_t<9>(row, K);
        case 35: return argmax_row_t<35>(row, K);
        default: return argmax_row_t<0>(row, K);
    }
}

// ---- SIMD-within-a-word helpers on packed bytes ---------------------------------------
// 0x01 in every byte of w that is non-zero (valid for byte values < 0x80)
__device__ __forceinline__ uint32_t bytes_nonzero01(uint32_t w) {
    return ((w + 0x7F7F7F7Fu) >> 7) & 0x01010101u;
}
// 0/1 bytes -> 0x00/0xFF bytes
__device__ __forceinline__ uint32_t bytes01_to_ff(uint32_t w) { return w * 255u; }
// 4 bits -> 4 bytes of 0/1 (bit e -> byte e)
__device__ __forceinline__ uint32_t bits4_to_bytes01(uint32_t bits) {
    return ((bits & 0xFu) * 0x00204081u) & 0x01010101u;
}

// Blank one 16-byte vector (index q) of a [16 px][C ch] uint8 group given the group's IM
// as 4 words of 0x00/0xFF bytes (imw[j] = pixels 4j..4j+3).  Word i = 4q+j of the group
// holds bytes of pixels (4i+e)/C, which all live in imw[i / C].
template <int C>
__device__ __forceinline__ uint4 blank_vec(uint4 v, const uint32_t (&imw)[4], int q) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = 4 * q + j;
        uint32_t sel = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) sel |= (uint32_t)((4 * (i % C) + e) / C) << (4 * e);
        w[j] &= ~__byte_perm(imw[i / C], 0u, sel);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// img_out[16 px] = im ? 0 : img[16 px]   (or a plain copy when !block_in)
template <int C>
__device__ __forceinline__ void blank_image16(const uint8_t *__restrict__ img, uint8_t *__restrict__ img_out,
                                              int64_t px, const uint32_t (&imw_in)[4], bool block_in) {
    const uint8_t *src = img + px * C;
    uint8_t *dst = img_out + px * C;
    uint4 v[C];
#pragma unroll
    for (int q = 0; q < C; ++q) v[q] = ldg_stream(src + 16 * q);
    uint32_t imw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) imw[j] = block_in ? imw_in[j] : 0u;
#pragma unroll
    for (int q = 0; q < C; ++q) stg_stream(dst + 16 * q, blank_vec<C>(v[q], imw, q));
}

__device__ __forceinline__ void blank_image16_any(const uint8_t *img, uint8_t *img_out, int c,
                                                  int64_t px, const uint32_t (&imw)[4], bool block_in) {
    switch (c) {
        case 1: blank_image16<1>(img, img_out, px, imw, block_in); break;
        case 2: blank_image16<2>(img, img_out, px, imw, block_in); break;
        case 3: blank_image16<3>(img, img_out, px, imw, block_in); break;
        default: blank_image16<4>(img, img_out, px, imw, block_in); break;
    }
}

// Variant for the warp-cooperative epilogues: vector q (0..C-1, runtime) of a 16-pixel group.  With the channel
// count known at compile time every byte selector folds to a constant (4 PRMT + 4 LOP per vector); C == 0 is the
// generic run-time path.
template <int C>
__device__ __forceinline__ uint4 blank_vec_sel(uint4 v, const uint32_t (&imw)[4], int c, int q) {
    if constexpr (C == 1) {
        return blank_vec<1>(v, imw, 0);
    } else if constexpr (C == 2) {
        return q == 0 ? blank_vec<2>(v, imw, 0) : blank_vec<2>(v, imw, 1);
    } else if constexpr (C == 3) {
        return q == 0 ? blank_vec<3>(v, imw, 0) : (q == 1 ? blank_vec<3>(v, imw, 1) : blank_vec<3>(v, imw, 2));
    } else if constexpr (C == 4) {
        return q < 2 ? (q == 0 ? blank_vec<4>(v, imw, 0) : blank_vec<4>(v, imw, 1)) : (q == 2 ? blank_vec<4>(v, imw, 2) : blank_vec<4>(v, imw, 3));
    } else {
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = 4 * q + j;
            const int r = i % c;
            uint32_t sel = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) sel |= (uint32_t)((4 * r + e) / c) << (4 * e);
            const int src = i / c;
            const uint32_t m = src == 0 ? imw[0] : src == 1 ? imw[1] : src == 2 ? imw[2] : imw[3];
            w[j] &= ~__byte_perm(m, 0u, sel);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
}
__device__ __forceinline__ uint4 blank_vec_rt(uint4 v, const uint32_t (&imw)[4], int c, int q) {
    return blank_vec_sel<0>(v, imw, c, q);
}

// Add per-lane counts into per-image int64 slots.  `n` is the lane's image index
// (-1 for lanes with nothing to add).  One RED per (warp, image).
__device__ __forceinline__ void warp_add_stat(int64_t *slots, int64_t n, uint32_t v, bool warp_uniform_n) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    if (warp_uniform_n) {
        const uint32_t s = __reduce_add_sync(full, v);
        if (lane == 0 && s) atomicAdd(reinterpret_cast<unsigned long long *>(slots + n), (unsigned long long)s);
    } else {
        const unsigned grp = __match_any_sync(full, n);
        const uint32_t s = __reduce_add_sync(grp, v);
        if (n >= 0 && lane == (__ffs(grp) - 1) && s)
            atomicAdd(reinterpret_cast<unsigned long long *>(slots + n), (unsigned long long)s);
    }
}

__device__ __forceinline__ void warp_or_stat(unsigned long long *slots, int64_t n, unsigned long long v,
                                             bool warp_uniform_n) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    if (warp_uniform_n) {
        lo = __reduce_or_sync(full, lo);
        hi = __reduce_or_sync(full, hi);
        if (lane == 0) atomicOr(slots + n, ((unsigned long long)hi << 32) | lo);
    } else {
        const unsigned grp = __match_any_sync(full, n);
        lo = __reduce_or_sync(grp, lo);
        hi = __reduce_or_sync(grp, hi);
        if (n >= 0 && lane == (__ffs(grp) - 1)) atomicOr(slots + n, ((unsigned long long)hi << 32) | lo);
    }
}

}  // namespace imk

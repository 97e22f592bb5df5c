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

// np.argmax (functions.py:3225): first index of the maximum, NaN is the maximum (the first NaN wins).
// argmax_step folds candidate (v, k) into the running (best, arg); every index folded so far must be
// smaller than k.  The rule is associative over ORDERED groups, so rows are reduced as trees of 8
// (8 independent shared-memory loads in flight, dependency depth 4 instead of 8).
__device__ __forceinline__ void argmax_step(float v, int k, float &best, int &arg) {
    // a NaN `best` can never be displaced; a NaN v displaces any non-NaN best
    if (!(best != best) && (v > best || v != v)) { best = v; arg = k; }
}

__device__ __forceinline__ int argmax_row(const float *__restrict__ row, int K) {
    float best = 0.f;
    int arg = 0, k = 0;
    for (; k + 8 <= K; k += 8) {
        float v[8];
        int i[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { v[e] = row[k + e]; i[e] = k + e; }
        argmax_step(v[1], i[1], v[0], i[0]); argmax_step(v[3], i[3], v[2], i[2]);
        argmax_step(v[5], i[5], v[4], i[4]); argmax_step(v[7], i[7], v[6], i[6]);
        argmax_step(v[2], i[2], v[0], i[0]); argmax_step(v[6], i[6], v[4], i[4]);
        argmax_step(v[4], i[4], v[0], i[0]);
        if (k == 0) { best = v[0]; arg = i[0]; } else argmax_step(v[0], i[0], best, arg);
    }
    for (; k < K; ++k) {
        const float v = row[k];
        if (k == 0) { best = v; arg = 0; } else argmax_step(v, k, best, arg);
    }
    return arg;
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

// Runtime-indexed variant for the warp-cooperative epilogues: vector q (0..c-1) of a 16-pixel group.
__device__ __forceinline__ uint4 blank_vec_rt(uint4 v, const uint32_t (&imw)[4], int c, int q) {
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

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

// np.argmax step (functions.py:3225): first index of the maximum, NaN is the maximum.
// Call with k ascending; `best` starts as the value at k = 0.
__device__ __forceinline__ void argmax_step(float v, int k, float &best, int &arg) {
    // a NaN `best` can never be displaced; a NaN v displaces any non-NaN best
    if (!(best != best) && (v > best || v != v)) { best = v; arg = k; }
}

// Expand a 16-pixel bit mask (bit i = pixel i is inside the IM) into the byte mask
// of the q-th 16-byte vector of a [16 px][C ch] uint8 group, and blank.
template <int C>
__device__ __forceinline__ uint4 blank_vec(uint4 v, uint32_t imbits, int q) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t keep = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int byte = q * 16 + i * 4 + e;
            const int px = byte / C;
            if (!((imbits >> px) & 1u)) keep |= 0xFFu << (8 * e);
        }
        w[i] &= keep;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// img_out[16 px] = im ? 0 : img[16 px]   (or a plain copy when !block_in)
template <int C>
__device__ __forceinline__ void blank_image16(const uint8_t *__restrict__ img, uint8_t *__restrict__ img_out,
                                              int64_t px, uint32_t imbits, bool block_in) {
    const uint8_t *src = img + px * C;
    uint8_t *dst = img_out + px * C;
    uint4 v[C];
#pragma unroll
    for (int q = 0; q < C; ++q) v[q] = ldg_stream(src + 16 * q);
    if (!block_in) imbits = 0;
#pragma unroll
    for (int q = 0; q < C; ++q) stg_stream(dst + 16 * q, blank_vec<C>(v[q], imbits, q));
}

__device__ __forceinline__ void blank_image16_any(const uint8_t *img, uint8_t *img_out, int c,
                                                  int64_t px, uint32_t imbits, bool block_in) {
    switch (c) {
        case 1: blank_image16<1>(img, img_out, px, imbits, block_in); break;
        case 2: blank_image16<2>(img, img_out, px, imbits, block_in); break;
        case 3: blank_image16<3>(img, img_out, px, imbits, block_in); break;
        default: blank_image16<4>(img, img_out, px, imbits, block_in); break;
    }
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

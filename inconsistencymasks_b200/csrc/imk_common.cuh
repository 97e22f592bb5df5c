// Shared host/device helpers of libimk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include "../../include/imk.h"

namespace imk {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs (planning constant of the cost models)
int num_sms();                 // SM count of the current device (persistent grids are sized from this)

// ---- thread-local error / launch accounting --------------------------------
void set_error(const char *fmt, ...);
int64_t &launch_counter();

#define IMK_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e_ = (call);                                                    \
        if (e_ != cudaSuccess) {                                                    \
            imk::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,            \
                           cudaGetErrorString(e_));                                 \
            return IMK_ECUDA;                                                       \
        }                                                                           \
    } while (0)

#define IMK_REQUIRE(cond, ...)                                                      \
    do {                                                                            \
        if (!(cond)) {                                                              \
            imk::set_error(__VA_ARGS__);                                            \
            return IMK_EINVAL;                                                      \
        }                                                                           \
    } while (0)

// Optional per-kernel timing (imk_profile_begin / imk_profile_end): when active, a pair of
// CUDA events brackets the launch on ITS stream; nothing synchronises until the end call.
struct ProfileScope {
    int slot;
    cudaStream_t stream;
    ProfileScope(const char *name, int tag, cudaStream_t s);
    ~ProfileScope();
};
bool profiling_active();                // true between imk_profile_begin and imk_profile_end (this thread)
#define IMK_PROFILE(name, tag, stream) imk::ProfileScope imk_prof_scope_((name), (tag), (stream))

// Call right after a <<<>>> launch.
#define IMK_LAUNCHED()                                                              \
    do {                                                                            \
        ++imk::launch_counter();                                                    \
        IMK_CUDA(cudaGetLastError());                                               \
    } while (0)

// ---- device helpers ----------------------------------------------------------
// Streaming 128-bit accesses: every byte of the IM path is touched once
// (ld.global.cs / st.global.cs: streaming, evict-first; plain intrinsics so the compiler is
// free to batch many independent loads per thread -- memory-level parallelism is what
// saturates HBM here).
__device__ __forceinline__ uint4 ldg_stream(const void *p) { return __ldcs(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ void stg_stream(void *p, const uint4 &v) { __stcs(reinterpret_cast<uint4 *>(p), v); }
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// The ONE definition of the output activations, shared by the materialised
// (.predict) path and the fused ensemble path so that both produce identical bits.
__device__ __forceinline__ float sigmoid_f32(float x) {
    return 1.0f / (1.0f + __expf(-x));
}

}  // namespace imk

// Microbenchmark: how long after an mbarrier phase completes does a waiting warp run again?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_wake mbar_wake.cu && ./mbar_wake
// Waiter variants: try_wait with a suspend-time hint (what imk_block_tc.cu uses), plain try_wait loop, test_wait polling.
// W waiting warps (lane 0 of each records), one arriving warp that arrives `delay` cycles after the start.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__device__ __forceinline__ void wait(uint64_t *bar, uint32_t parity, uint32_t hint) {
    if (MODE == 0) {
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                     :: "r"(smem_u32(bar)), "r"(parity), "r"(hint) : "memory");
    } else if (MODE == 1) {
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                     :: "r"(smem_u32(bar)), "r"(parity) : "memory");
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                     :: "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

template <int MODE>
__global__ void k(long long *out, int W, int delay, uint32_t hint, int rounds) {
    __shared__ uint64_t bar;
    __shared__ long long t_arr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); }
    __syncthreads();
    long long sum = 0, mx = 0;
    for (int r = 0; r < rounds; ++r) {
        __syncthreads();
        if (warp == W) {                      // the arriving warp
            const long long t0 = clock64();
            while (clock64() - t0 < delay) { }
            if (lane == 0) {
                t_arr = clock64();
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&bar)) : "memory");
            }
        } else {
            wait<MODE>(&bar, (uint32_t)(r & 1), hint);
            const long long t1 = clock64();
            __syncwarp();
            if (lane == 0) { const long long d = t1 - *(volatile long long *)&t_arr; sum += d; if (d > mx) mx = d; }
        }
    }
    if (lane == 0 && warp < W) { out[2 * warp] = sum / rounds; out[2 * warp + 1] = mx; }
}

template <int MODE>
void run(const char *name, int W, int delay, uint32_t hint) {
    long long *d; cudaMalloc(&d, 64 * 16);
    cudaMemset(d, 0, 64 * 16);
    k<MODE><<<1, (W + 1) * 32>>>(d, W, delay, hint, 50);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mn = 1 << 30, mxa = 0, mxm = 0;
    for (int w = 0; w < W; ++w) { if (h[2 * w] < mn) mn = h[2 * w]; if (h[2 * w] > mxa) mxa = h[2 * w]; if (h[2 * w + 1] > mxm) mxm = h[2 * w + 1]; }
    printf("%-34s W=%2d delay=%6d hint=%6u : wake latency avg min %5lld max %5lld cycles (worst single %lld)\n", name, W, delay, hint, mn, mxa, mxm);
    cudaFree(d);
}

int main() {
    for (int W : {1, 4, 16, 24}) {
        for (int delay : {500, 5000, 50000}) {
            run<0>("try_wait + suspend hint", W, delay, 20000);
            run<0>("try_wait + suspend hint", W, delay, 1000);
            run<0>("try_wait + suspend hint", W, delay, 100);
            run<1>("try_wait (no hint)", W, delay, 0);
            run<2>("test_wait polling", W, delay, 0);
        }
    }
    return 0;
}

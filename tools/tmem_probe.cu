// tmem_probe.cu - can Tensor Memory serve as per-thread scratch?  12 warps (384 threads, the shape of a 6-group rows CTA) each
// write and read back 132 32-bit columns of their own TMEM lanes with tcgen05.st / tcgen05.ld (.32x32b), every pattern checked,
// then the round trip is timed.  Addressing under test: lane field = 32 * (warp % 4), column field = 132 * (warp / 4) + c.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmem_probe tools/tmem_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tm_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tm_ld4(uint32_t taddr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(384, 1) probe(unsigned* errors, long long* cycles, int reps) {
    __shared__ uint32_t tm_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tm_base;
    const uint32_t mine = base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(132 * (warp >> 2));
    unsigned bad = 0;
    for (int round = 0; round < 3; ++round) {
#pragma unroll 1
        for (int c = 0; c < 132; c += 4) {
            const uint32_t v = (uint32_t)(blockIdx.x * 1000003 + warp * 4099 + lane * 131 + c + round * 7);
            tm_st4(mine + c, v, v + 1, v + 2, v + 3);
        }
        tm_wait_st();
        __syncthreads();      // everybody has written: a wrong lane / column mapping would have clobbered somebody else's values
#pragma unroll 1
        for (int c = 0; c < 132; c += 4) {
            uint32_t a, b, cc, d;
            tm_ld4(mine + c, a, b, cc, d);
            tm_wait_ld();
            const uint32_t v = (uint32_t)(blockIdx.x * 1000003 + warp * 4099 + lane * 131 + c + round * 7);
            bad += (a != v) + (b != v + 1) + (cc != v + 2) + (d != v + 3);
        }
        __syncthreads();
    }
    // timing: the access pattern of the cross phase - per step load 4 columns, modify, store 4 columns
    const long long t0 = clock64();
    uint32_t acc = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 1
        for (int c = 0; c < 132; c += 4) {
            uint32_t a, b, cc, d;
            tm_ld4(mine + c, a, b, cc, d);
            tm_wait_ld();
            acc += a ^ b ^ cc ^ d;
            tm_st4(mine + c, a + 1, b + 1, cc + 1, d + 1);
        }
        tm_wait_st();
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) bad += 1000000;
    atomicAdd(errors, bad);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    unsigned* d_err; long long* d_cyc;
    cudaMalloc(&d_err, 4); cudaMemset(d_err, 0, 4);
    cudaMalloc(&d_cyc, 8 * p.multiProcessorCount);
    const int reps = 200;
    probe<<<p.multiProcessorCount, 384>>>(d_err, d_cyc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned err = 0; long long cyc = 0;
    cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("status %s, mismatches %u (of %d checked values per thread), %.1f cycles per {ld x4, wait, st x4} step with 12 warps per SM\n",
           cudaGetErrorString(e), err, 3 * 132, (double)cyc / (reps * 33.0));
    return e != cudaSuccess || err != 0;
}

// fp32x2_bench.cu - does Blackwell's packed fp32 arithmetic (add / mul / fma .f32x2 -> FADD2 / FMUL2 / FFMA2) buy issue slots?
// Times independent chains of scalar and packed instructions, alone and mixed with shared-memory / integer work, on every SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp32x2_bench tools/fp32x2_bench.cu && gpurun_out/fp32x2_bench
#include <cuda_runtime.h>
#include <cstdio>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int NCH = 16;   // independent chains per thread

template <int MODE>
__global__ void __launch_bounds__(256) kern(float* out, int iters, float fa, float fb) {
    __shared__ float sm[256 * 4];
    float s = 0.f;
    if (MODE == 0 || MODE == 1) {           // scalar FADD / FFMA: 2 * NCH scalar chains = as many VALUES as the packed variants
        float x[2 * NCH];
#pragma unroll
        for (int k = 0; k < 2 * NCH; ++k) x[k] = threadIdx.x * 1e-3f + k;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 2 * NCH; ++k) x[k] = MODE == 0 ? x[k] + fb : fmaf(x[k], fa, fb);
        }
#pragma unroll
        for (int k = 0; k < 2 * NCH; ++k) s += x[k];
    } else if (MODE == 2 || MODE == 3 || MODE == 4) {   // packed FADD2 / FFMA2 / FMUL2
        u64 x[NCH];
        const u64 a = pack(fa, fa), b = pack(fb, fb);
#pragma unroll
        for (int k = 0; k < NCH; ++k) x[k] = pack(threadIdx.x * 1e-3f + k, k);
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) x[k] = MODE == 2 ? add2(x[k], b) : (MODE == 3 ? fma2(x[k], a, b) : mul2(x[k], a));
        }
#pragma unroll
        for (int k = 0; k < NCH; ++k) s += lo(x[k]);
    } else if (MODE == 5 || MODE == 6) {    // FFT-like mix: per 8 values 6 adds + 2 fma, plus one LDS + one STS per 16 FP ops (scalar / packed)
        sm[threadIdx.x] = fa;
        __syncthreads();
        if (MODE == 5) {
            float x[2 * NCH];
#pragma unroll
            for (int k = 0; k < 2 * NCH; ++k) x[k] = threadIdx.x * 1e-3f + k;
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int k = 0; k < 2 * NCH; ++k) x[k] = (k & 3) == 3 ? fmaf(x[k], fa, fb) : x[k] + x[(k + 2) % (2 * NCH)];
                x[0] += sm[(threadIdx.x + i) & 255];
                sm[256 + threadIdx.x] = x[1];
                x[2] += sm[(threadIdx.x + 2 * i) & 255];
                sm[512 + threadIdx.x] = x[3];
            }
#pragma unroll
            for (int k = 0; k < 2 * NCH; ++k) s += x[k];
        } else {
            u64 x[NCH];
            const u64 a = pack(fa, fa), b = pack(fb, fb);
#pragma unroll
            for (int k = 0; k < NCH; ++k) x[k] = pack(threadIdx.x * 1e-3f + k, k);
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int k = 0; k < NCH; ++k) x[k] = (k & 3) == 3 ? fma2(x[k], a, b) : add2(x[k], x[(k + 1) % NCH]);
                x[0] = add2(x[0], pack(sm[(threadIdx.x + i) & 255], 0.f));
                sm[256 + threadIdx.x] = lo(x[1]);
                x[2] = add2(x[2], pack(sm[(threadIdx.x + 2 * i) & 255], 0.f));
                sm[512 + threadIdx.x] = lo(x[3]);
            }
#pragma unroll
            for (int k = 0; k < NCH; ++k) s += lo(x[k]);
        }
    }
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static double run(float* d, int grid, int iters) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(a);
        kern<MODE><<<grid, 256>>>(d, iters, 0.999f, 1e-3f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r && ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    const int iters = 20000;
    const char* names[] = {"FADD   (32 scalar chains)", "FFMA   (32 scalar chains)", "FADD2  (16 packed chains)", "FFMA2  (16 packed chains)",
                           "FMUL2  (16 packed chains)", "mix scalar (3 FADD : 1 FFMA + LDS/STS)", "mix packed (3 FADD2 : 1 FFMA2 + LDS/STS)"};
    for (int wps = 1; wps <= 8; wps *= 2) {      // blocks of 256 threads per SM = warps per scheduler / 2
        const int grid = p.multiProcessorCount * wps;
        double ms[7] = {run<0>(d, grid, iters), run<1>(d, grid, iters), run<2>(d, grid, iters), run<3>(d, grid, iters),
                        run<4>(d, grid, iters), run<5>(d, grid, iters), run<6>(d, grid, iters)};
        for (int m = 0; m < 7; ++m) {
            const double vals = 32.0 * iters * 256.0 * grid;                  // fp32 values updated
            const double flop = (m == 1 || m == 3) ? 2.0 : 1.0;
            printf("blocks/SM %d  %-44s %8.3f ms  %7.2f Tvalue-ops/s  %7.2f TFLOP/s\n", wps, names[m], ms[m], vals / (ms[m] * 1e-3) / 1e12,
                   (m < 5 ? flop : 1.25) * vals / (ms[m] * 1e-3) / 1e12);
        }
    }
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    return 0;
}

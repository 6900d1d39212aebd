// project.cuh - orthoprojection of camera frames onto the PIV grid with pre-computed index maps.
//
// Replaces pyorc.project.img_to_ortho + _group_average (pyorc/project.py:19-53, :123-157) as used by project_numpy
// (project.py:160-230): nearest-neighbour gather for under-sampled target pixels, float32 group mean over the source
// pixels that fall into an over-sampled target pixel.  The reference scatters (sum/count arrays, then a gather-assign);
// here both maps are merged ONCE per camera configuration into one CSR list per target pixel (b2piv_project_plan), so
// a frame is a pure gather: one thread per target pixel, the source samples of a pixel summed in the reference's order
// (ascending source index) in float32, divided by the count, and written once.  No atomics, no zero-fill pass.
//
// HBM-bound: per frame the touched source pixels (each read once; neighbouring target pixels read neighbouring source
// pixels, so sectors are shared through L1/L2) + the CSR (4 B per sample + 4 B per target pixel, amortised over FR
// frames held in registers) + one store per target pixel.
#pragma once
#include <cuda_runtime.h>

namespace b2piv {

template <typename T>
__device__ __forceinline__ T proj_cast(float v);
template <>
__device__ __forceinline__ float proj_cast<float>(float v) { return v; }
// np.vectorize(otypes=[uint8]) / astype: truncation toward zero (project.py:205-227 `output_dtypes=[da.dtype]`)
template <>
__device__ __forceinline__ unsigned char proj_cast<unsigned char>(float v) { return (unsigned char)(int)v; }

template <typename TI, typename TO, int FR>
__global__ void __launch_bounds__(256) proj_gather_kernel(const TI* __restrict__ frames, long long frame_elems, int n_frames,
                                                          const int* __restrict__ off, const int* __restrict__ src, int n_out,
                                                          TO* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int f0 = blockIdx.y * FR;
    if (j >= n_out) return;
    const int b = off[j], e = off[j + 1];
    float acc[FR];
#pragma unroll
    for (int r = 0; r < FR; ++r) acc[r] = 0.f;
    const TI* base = frames + (long long)f0 * frame_elems;
    if (f0 + FR <= n_frames) {
        for (int k = b; k < e; ++k) {
            const TI* p = base + src[k];
#pragma unroll
            for (int r = 0; r < FR; ++r) acc[r] = __fadd_rn(acc[r], (float)p[(long long)r * frame_elems]);
        }
    } else {
        for (int k = b; k < e; ++k) {
            const TI* p = base + src[k];
#pragma unroll
            for (int r = 0; r < FR; ++r)
                if (f0 + r < n_frames) acc[r] = __fadd_rn(acc[r], (float)p[(long long)r * frame_elems]);
        }
    }
    const int cnt = e - b;
    const float fc = (float)cnt;
    TO* o = out + (long long)f0 * n_out + j;
#pragma unroll
    for (int r = 0; r < FR; ++r)
        if (f0 + r < n_frames) {
            // float32 sum / int64 count is evaluated in float64 by numba and rounded to float32 on the store
            // (project.py:50-52); with a 53-bit intermediate that double rounding is innocuous, i.e. == __fdiv_rn
            const float a = cnt > 1 ? __fdiv_rn(acc[r], fc) : acc[r];
            o[(long long)r * n_out] = proj_cast<TO>(a);
        }
}

}  // namespace b2piv

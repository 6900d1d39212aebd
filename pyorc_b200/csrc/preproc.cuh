// preproc.cuh - frame pre-processing that feeds the PIV engine, device-resident so that frames never
// go back to the host between `normalize()` / `time_diff()` / `smooth()` / `edge_detect()` and `get_piv()`.
// All of it is HBM-bound element / small-stencil work: coalesced 16-byte accesses, grids sized to the SM count.
//
// Reference semantics (pyorc @ be7d7c8):
//   normalize   pyorc/api/frames.py:279-306   frames - mean(sampled frames) -> per-frame min/max stretch -> uint8
//   minmax      pyorc/api/frames.py:343-361   clamp
//   time_diff   pyorc/api/frames.py:403-430   float32 difference of consecutive frames, thresholded, optional abs
//   smooth      pyorc/api/frames.py:432-466 -> pyorc/cv.py:142-159   cv2.GaussianBlur(float32, (k,k), 0)
//   edge_detect pyorc/api/frames.py:308-341 -> pyorc/cv.py:162-183   GaussianBlur(k2) - GaussianBlur(k1)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2piv {

// 16 consecutive elements per thread and iteration (one 16-byte load for uint8, four for float32): every kernel below
// is a pure stream, so the only things that matter are wide coalesced accesses and enough of them in flight.
template <typename T> struct Vec16;
template <> struct Vec16<unsigned char> {
    static __device__ __forceinline__ void load(const unsigned char* p, float* v) {
        const uint4 q = *reinterpret_cast<const uint4*>(p);
        const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (float)((w[i >> 2] >> (8 * (i & 3))) & 0xffu);
    }
    static __device__ __forceinline__ void store(unsigned char* p, const float* v) {   // values already in [0, 255]
        unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i >> 2] |= ((unsigned)(int)v[i] & 0xffu) << (8 * (i & 3));
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};
template <> struct Vec16<float> {
    static __device__ __forceinline__ void load(const float* p, float* v) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 q = reinterpret_cast<const float4*>(p)[j];
            v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
        }
    }
    static __device__ __forceinline__ void store(float* p, const float* v) {
#pragma unroll
        for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(p)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
};
__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- normalize -------------------------------------------------------------------------------------------------
// K1: per-pixel mean over the sampled frames; integer sums are exact, the division runs in double like numpy's
// mean of an integer array (float64 accumulator), then the reference casts to float32.
template <typename T>
__global__ void __launch_bounds__(256) pre_mean_kernel(const T* __restrict__ frames, long long frame_elems, int n_frames, int step,
                                                       float* __restrict__ mean) {
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    for (int f = 0; f < n_frames; f += step) ++cnt;
    const long long nv = (aligned16(frames) && aligned16(mean) && frame_elems % 16 == 0) ? frame_elems : 0;
    for (long long i = t0 * 16; i < nv; i += stride * 16) {
        double acc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = 0.0;
        for (int f = 0; f < n_frames; f += step) {
            float a[16];
            Vec16<T>::load(frames + (long long)f * frame_elems + i, a);
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] += (double)a[k];
        }
        float m[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) m[k] = (float)(acc[k] / (double)cnt);
        Vec16<float>::store(mean + i, m);
    }
    for (long long i = nv + t0; i < frame_elems; i += stride) {
        double acc = 0.0;
        for (int f = 0; f < n_frames; f += step) acc += (double)frames[(long long)f * frame_elems + i];
        mean[i] = (float)(acc / (double)cnt);
    }
}

// order-preserving float <-> unsigned mapping for atomicMin / atomicMax
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// K2: per-frame min / max of (float32(frame) - mean); minmax[2f] = ordered min, [2f+1] = ordered max
template <typename T>
__global__ void __launch_bounds__(256) pre_minmax_kernel(const T* __restrict__ frames, const float* __restrict__ mean, long long frame_elems,
                                                         unsigned* __restrict__ minmax) {
    const int f = blockIdx.y;
    const T* fr = frames + (long long)f * frame_elems;
    float mn = INFINITY, mx = -INFINITY;
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = aligned16(fr) && aligned16(mean) && (frame_elems % 16 == 0);
    if (vec) {
        for (long long i = t0 * 16; i < frame_elems; i += stride * 16) {
            float a[16], m[16];
            Vec16<T>::load(fr + i, a);
            Vec16<float>::load(mean + i, m);
#pragma unroll
            for (int k = 0; k < 16; ++k) { const float d = __fsub_rn(a[k], m[k]); mn = fminf(mn, d); mx = fmaxf(mx, d); }
        }
    } else {
        for (long long i = t0; i < frame_elems; i += stride) {
            const float d = __fsub_rn((float)fr[i], mean[i]);
            mn = fminf(mn, d); mx = fmaxf(mx, d);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&minmax[2 * f], f2ord(mn));
        atomicMax(&minmax[2 * f + 1], f2ord(mx));
    }
}

// K3: ((d - min) / (max - min) * 255).astype(uint8) with numpy's float32 operation order (no contraction)
template <typename T>
__global__ void __launch_bounds__(256) pre_normalize_kernel(const T* __restrict__ frames, const float* __restrict__ mean,
                                                            const unsigned* __restrict__ minmax, long long frame_elems,
                                                            unsigned char* __restrict__ out) {
    const int f = blockIdx.y;
    const T* fr = frames + (long long)f * frame_elems;
    unsigned char* o = out + (long long)f * frame_elems;
    const float mn = ord2f(minmax[2 * f]), mx = ord2f(minmax[2 * f + 1]);
    const float range = __fsub_rn(mx, mn);
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = aligned16(fr) && aligned16(mean) && aligned16(o) && (frame_elems % 16 == 0);
    if (vec) {
        for (long long i = t0 * 16; i < frame_elems; i += stride * 16) {
            float a[16], m[16];
            Vec16<T>::load(fr + i, a);
            Vec16<float>::load(mean + i, m);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float v = __fmul_rn(__fdiv_rn(__fsub_rn(__fsub_rn(a[k], m[k]), mn), range), 255.0f);
                a[k] = (v >= 0.f && v < 256.f) ? v : 0.f;
            }
            Vec16<unsigned char>::store(o + i, a);
        }
    } else {
        for (long long i = t0; i < frame_elems; i += stride) {
            const float d = __fsub_rn((float)fr[i], mean[i]);
            const float v = __fmul_rn(__fdiv_rn(__fsub_rn(d, mn), range), 255.0f);
            o[i] = (v >= 0.f && v < 256.f) ? (unsigned char)(int)v : (unsigned char)0;   // NaN (flat frame) -> 0
        }
    }
}

// ---- time_diff / minmax ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pre_time_diff_kernel(const T* __restrict__ frames, long long frame_elems, long long n_out, float thres,
                                                            int absolute, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = aligned16(frames) && aligned16(out) && (frame_elems % 16 == 0);
    if (vec) {
        for (long long i = t0 * 16; i < n_out; i += stride * 16) {
            float a[16], b[16];
            Vec16<T>::load(frames + i, a);
            Vec16<T>::load(frames + i + frame_elems, b);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float d = __fsub_rn(b[k], a[k]);
                float v = d > thres ? d : 0.f;
                a[k] = absolute ? fabsf(v) : v;
            }
            Vec16<float>::store(out + i, a);
        }
    } else {
        for (long long i = t0; i < n_out; i += stride) {
            const float d = __fsub_rn((float)frames[i + frame_elems], (float)frames[i]);
            float v = d > thres ? d : 0.f;     // where(diff > thres) ... fillna(0.0); NaN compares false -> 0
            if (absolute) v = fabsf(v);
            out[i] = v;
        }
    }
}
template <typename T>
__global__ void __launch_bounds__(256) pre_clamp_kernel(const T* __restrict__ in, long long n, float lo, float hi, T* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nv = (aligned16(in) && aligned16(out)) ? (n / 16) * 16 : 0;
    for (long long i = t0 * 16; i < nv; i += stride * 16) {
        float a[16];
        Vec16<T>::load(in + i, a);
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fmaxf(fminf(a[k], hi), lo);
        Vec16<T>::store(out + i, a);
    }
    for (long long i = nv + t0; i < n; i += stride) {
        const float x = (float)in[i];
        out[i] = (T)fmaxf(fminf(x, hi), lo);   // np.maximum(np.minimum(x, max), min)
    }
}

// ---- Gaussian blur / band filter -----------------------------------------------------------------------------------
// Separable float32 blur with BORDER_REFLECT_101, both passes through shared memory; out = blur(k2) - blur(k1) when
// k1 > 0 (edge_detect) else blur(k2) (smooth).  Coefficients come from the host (OpenCV's getGaussianKernel rule).
constexpr int GB_TX = 32, GB_TY = 16, GB_MAXR = 15;
struct GaussTaps {
    float k1[2 * GB_MAXR + 1];
    float k2[2 * GB_MAXR + 1];
    int r1, r2;   // radii; r1 < 0: no first kernel
};
__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}
template <typename T>
__global__ void __launch_bounds__(GB_TX* GB_TY) pre_gauss_kernel(const T* __restrict__ frames, int H, int W, GaussTaps taps,
                                                                 float* __restrict__ out) {
    extern __shared__ float gsm[];
    const int R = taps.r2 > taps.r1 ? taps.r2 : taps.r1;
    const int tw = GB_TX + 2 * R, th = GB_TY + 2 * R;
    float* tile = gsm;                       // [th][tw] source
    float* rows1 = tile + th * tw;           // [th][GB_TX] row-filtered with k1
    float* rows2 = rows1 + th * GB_TX;       // [th][GB_TX] row-filtered with k2
    const long long fe = (long long)H * W;
    const T* fr = frames + (long long)blockIdx.z * fe;
    float* o = out + (long long)blockIdx.z * fe;
    const int x0 = blockIdx.x * GB_TX, y0 = blockIdx.y * GB_TY;
    // 2-D strided loops (no div / mod); reflect-101 only matters for the border tiles
    for (int ty = threadIdx.y; ty < th; ty += GB_TY) {
        const int gy = reflect101(y0 + ty - R, H);
        const T* src = fr + (long long)gy * W;
        for (int tx = threadIdx.x; tx < tw; tx += GB_TX) tile[ty * tw + tx] = (float)src[reflect101(x0 + tx - R, W)];
    }
    __syncthreads();
    for (int ty = threadIdx.y; ty < th; ty += GB_TY) {
        const float* trow = tile + ty * tw + threadIdx.x + R;
        float a2 = 0.f, a1 = 0.f;
        for (int j = -taps.r2; j <= taps.r2; ++j) a2 = __fadd_rn(a2, __fmul_rn(taps.k2[j + taps.r2], trow[j]));
        if (taps.r1 >= 0)
            for (int j = -taps.r1; j <= taps.r1; ++j) a1 = __fadd_rn(a1, __fmul_rn(taps.k1[j + taps.r1], trow[j]));
        rows2[ty * GB_TX + threadIdx.x] = a2;
        rows1[ty * GB_TX + threadIdx.x] = a1;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x < W && y < H) {
        float a2 = 0.f, a1 = 0.f;
        for (int j = -taps.r2; j <= taps.r2; ++j) a2 = __fadd_rn(a2, __fmul_rn(taps.k2[j + taps.r2], rows2[(threadIdx.y + R + j) * GB_TX + threadIdx.x]));
        if (taps.r1 >= 0)
            for (int j = -taps.r1; j <= taps.r1; ++j) a1 = __fadd_rn(a1, __fmul_rn(taps.k1[j + taps.r1], rows1[(threadIdx.y + R + j) * GB_TX + threadIdx.x]));
        o[(long long)y * W + x] = taps.r1 >= 0 ? __fsub_rn(a2, a1) : a2;
    }
}

}  // namespace b2piv

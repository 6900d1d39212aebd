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

// Four consecutive elements per thread and access (one 4-byte load for uint8, one 16-byte load for float32, so that a
// warp instruction always covers one contiguous 128- / 512-byte span on BOTH sides of a mixed uint8 / float32 kernel),
// PRE_U independent accesses in flight per thread: every kernel below is a pure stream, so the only things that
// matter are coalescing and enough bytes in flight (a first version with 16 elements = 64 contiguous float bytes per
// thread ran at a quarter of this: lanes 64 B apart use half of each 32-byte sector per instruction).
constexpr int PRE_U = 4;
template <typename T> struct Vec4;
template <> struct Vec4<unsigned char> {
    static __device__ __forceinline__ void load(const unsigned char* p, float* v) {
        const uchar4 q = *reinterpret_cast<const uchar4*>(p);
        v[0] = (float)q.x; v[1] = (float)q.y; v[2] = (float)q.z; v[3] = (float)q.w;
    }
    static __device__ __forceinline__ void store(unsigned char* p, const float* v) {   // values already in [0, 255]
        *reinterpret_cast<uchar4*>(p) = make_uchar4((unsigned char)(int)v[0], (unsigned char)(int)v[1], (unsigned char)(int)v[2],
                                                    (unsigned char)(int)v[3]);
    }
};
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float* v) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    static __device__ __forceinline__ void store(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
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
    const long long nv = (aligned16(frames) && aligned16(mean) && frame_elems % 4 == 0) ? frame_elems : 0;
    for (long long i = t0 * 4; i < nv; i += stride * 4) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        int f = 0;
        for (; f + (PRE_U - 1) * step < n_frames; f += PRE_U * step) {
            float a[PRE_U][4];
#pragma unroll
            for (int u = 0; u < PRE_U; ++u) Vec4<T>::load(frames + (long long)(f + u * step) * frame_elems + i, a[u]);
#pragma unroll
            for (int u = 0; u < PRE_U; ++u)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] += (double)a[u][k];
        }
        for (; f < n_frames; f += step) {
            float a[4];
            Vec4<T>::load(frames + (long long)f * frame_elems + i, a);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] += (double)a[k];
        }
        float m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = (float)(acc[k] / (double)cnt);
        Vec4<float>::store(mean + i, m);
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

// K2: per-frame min / max of (float32(frame) - mean); minmax[f] = ordered min, minmax[n_frames + f] = ordered max
// (so that two cudaMemsetAsync calls initialise them: 0xff.. for the minima, 0 for the maxima).
// A block covers PRE_FPB consecutive frames (blockIdx.y) so that the float32 mean - 4x the bytes of a uint8 frame - is
// loaded once per PRE_FPB frames instead of once per frame (per-frame blocks ran at the L2 rate of the mean, 1.5 TB/s).
constexpr int PRE_FPB = 8;
template <typename T>
__global__ void __launch_bounds__(256) pre_minmax_kernel(const T* __restrict__ frames, const float* __restrict__ mean, long long frame_elems,
                                                         int n_frames, unsigned* __restrict__ minmax) {
    const int f0 = blockIdx.y * PRE_FPB;
    const int nf = n_frames - f0 < PRE_FPB ? n_frames - f0 : PRE_FPB;
    const T* fr = frames + (long long)f0 * frame_elems;
    float mn[PRE_FPB], mx[PRE_FPB];
#pragma unroll
    for (int k = 0; k < PRE_FPB; ++k) { mn[k] = INFINITY; mx[k] = -INFINITY; }
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = aligned16(frames) && aligned16(mean) && (frame_elems % 4 == 0);
    if (vec) {
        for (long long i = t0 * 4; i < frame_elems; i += stride * 4) {
            float m[4], a[PRE_FPB][4];
            Vec4<float>::load(mean + i, m);
#pragma unroll
            for (int k = 0; k < PRE_FPB; ++k)
                if (k < nf) Vec4<T>::load(fr + (long long)k * frame_elems + i, a[k]);
#pragma unroll
            for (int k = 0; k < PRE_FPB; ++k)
                if (k < nf) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float d = __fsub_rn(a[k][j], m[j]); mn[k] = fminf(mn[k], d); mx[k] = fmaxf(mx[k], d); }
                }
        }
    } else {
        for (long long i = t0; i < frame_elems; i += stride) {
            const float m = mean[i];
#pragma unroll
            for (int k = 0; k < PRE_FPB; ++k)
                if (k < nf) { const float d = __fsub_rn((float)fr[(long long)k * frame_elems + i], m); mn[k] = fminf(mn[k], d); mx[k] = fmaxf(mx[k], d); }
        }
    }
    __shared__ float smn[8][PRE_FPB], smx[8][PRE_FPB];
#pragma unroll
    for (int k = 0; k < PRE_FPB; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5][k] = mn[k]; smx[threadIdx.x >> 5][k] = mx[k]; }
    }
    __syncthreads();
    if (threadIdx.x < nf) {
        float a = smn[0][threadIdx.x], b = smx[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < 8; ++w) { a = fminf(a, smn[w][threadIdx.x]); b = fmaxf(b, smx[w][threadIdx.x]); }
        atomicMin(&minmax[f0 + threadIdx.x], f2ord(a));
        atomicMax(&minmax[n_frames + f0 + threadIdx.x], f2ord(b));
    }
}

// K3: ((d - min) / (max - min) * 255).astype(uint8) with numpy's float32 operation order (no contraction)
template <typename T>
__global__ void __launch_bounds__(256) pre_normalize_kernel(const T* __restrict__ frames, const float* __restrict__ mean,
                                                            const unsigned* __restrict__ minmax, long long frame_elems, int n_frames,
                                                            unsigned char* __restrict__ out) {
    const int f0 = blockIdx.y * PRE_FPB;
    const int nf = n_frames - f0 < PRE_FPB ? n_frames - f0 : PRE_FPB;
    const T* fr = frames + (long long)f0 * frame_elems;
    unsigned char* o = out + (long long)f0 * frame_elems;
    float mn[PRE_FPB], range[PRE_FPB];
#pragma unroll
    for (int k = 0; k < PRE_FPB; ++k) {
        const int f = k < nf ? f0 + k : f0;
        mn[k] = ord2f(minmax[f]);
        range[k] = __fsub_rn(ord2f(minmax[n_frames + f]), mn[k]);
    }
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = aligned16(frames) && aligned16(mean) && aligned16(out) && (frame_elems % 4 == 0);
    if (vec) {
        for (long long i = t0 * 4; i < frame_elems; i += stride * 4) {
            float m[4], a[PRE_FPB][4];
            Vec4<float>::load(mean + i, m);
#pragma unroll
            for (int k = 0; k < PRE_FPB; ++k)
                if (k < nf) Vec4<T>::load(fr + (long long)k * frame_elems + i, a[k]);
#pragma unroll
            for (int k = 0; k < PRE_FPB; ++k)
                if (k < nf) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float v = __fmul_rn(__fdiv_rn(__fsub_rn(__fsub_rn(a[k][j], m[j]), mn[k]), range[k]), 255.0f);
                        a[k][j] = (v >= 0.f && v < 256.f) ? v : 0.f;   // NaN (flat frame) -> 0
                    }
                    Vec4<unsigned char>::store(o + (long long)k * frame_elems + i, a[k]);
                }
        }
    } else {
        for (long long i = t0; i < frame_elems; i += stride) {
            const float m = mean[i];
#pragma unroll
            for (int k = 0; k < PRE_FPB; ++k)
                if (k < nf) {
                    const float d = __fsub_rn((float)fr[(long long)k * frame_elems + i], m);
                    const float v = __fmul_rn(__fdiv_rn(__fsub_rn(d, mn[k]), range[k]), 255.0f);
                    o[(long long)k * frame_elems + i] = (v >= 0.f && v < 256.f) ? (unsigned char)(int)v : (unsigned char)0;
                }
        }
    }
}

// ---- time_diff / minmax ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pre_time_diff_kernel(const T* __restrict__ frames, long long frame_elems, long long n_out, float thres,
                                                            int absolute, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = aligned16(frames) && aligned16(out) && (frame_elems % 4 == 0);
    if (vec) {
        for (long long i = t0 * 4; i < n_out; i += stride * 4 * PRE_U) {
            float a[PRE_U][4], b[PRE_U][4];
#pragma unroll
            for (int u = 0; u < PRE_U; ++u) {
                const long long ii = i + u * stride * 4;
                if (ii < n_out) { Vec4<T>::load(frames + ii, a[u]); Vec4<T>::load(frames + ii + frame_elems, b[u]); }
            }
#pragma unroll
            for (int u = 0; u < PRE_U; ++u) {
                const long long ii = i + u * stride * 4;
                if (ii < n_out) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float d = __fsub_rn(b[u][k], a[u][k]);
                        const float v = d > thres ? d : 0.f;   // where(diff > thres) ... fillna(0.0); NaN compares false -> 0
                        a[u][k] = absolute ? fabsf(v) : v;
                    }
                    Vec4<float>::store(out + ii, a[u]);
                }
            }
        }
    } else {
        for (long long i = t0; i < n_out; i += stride) {
            const float d = __fsub_rn((float)frames[i + frame_elems], (float)frames[i]);
            float v = d > thres ? d : 0.f;
            if (absolute) v = fabsf(v);
            out[i] = v;
        }
    }
}
template <typename T>
__global__ void __launch_bounds__(256) pre_clamp_kernel(const T* __restrict__ in, long long n, float lo, float hi, T* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nv = (aligned16(in) && aligned16(out)) ? (n / 4) * 4 : 0;
    for (long long i = t0 * 4; i < nv; i += stride * 4 * PRE_U) {
        float a[PRE_U][4];
#pragma unroll
        for (int u = 0; u < PRE_U; ++u) {
            const long long ii = i + u * stride * 4;
            if (ii < nv) Vec4<T>::load(in + ii, a[u]);
        }
#pragma unroll
        for (int u = 0; u < PRE_U; ++u) {
            const long long ii = i + u * stride * 4;
            if (ii < nv) {
#pragma unroll
                for (int k = 0; k < 4; ++k) a[u][k] = fmaxf(fminf(a[u][k], hi), lo);   // np.maximum(np.minimum(x, max), min)
                Vec4<T>::store(out + ii, a[u]);
            }
        }
    }
    for (long long i = nv + t0; i < n; i += stride) {
        const float x = (float)in[i];
        out[i] = (T)fmaxf(fminf(x, hi), lo);
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

// Fast path for the common kernel sizes (pyorc defaults: smooth wdw=1 -> 3 taps, edge_detect wdw_1=1, wdw_2=2 -> 3 and 5
// taps): a 128-thread block marches down a 128-column strip row by row.  Per row: one coalesced load per thread into a
// double-buffered shared row (ONE barrier per row), the row filter from shared memory, and the column filter from a
// sliding window of row sums kept in registers (the march is fully unrolled, so the window shifts are register renames).
// Every input pixel is read (2R + GS_SH) / GS_SH times instead of being staged through three shared-memory tiles with
// two barriers per 32x16 output pixels (pre_gauss_kernel above, kept for the other sizes), which was latency-bound at
// 8 % of the HBM rate.  Same accumulation order as above (ascending tap index, separate multiply and add).
constexpr int GS_BW = 128, GS_SH = 32;
__device__ __forceinline__ int reflect_clamp(int i, int n) {
    i = i < 0 ? -i : i;
    i = i >= n ? 2 * n - 2 - i : i;
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}
template <typename T, int R1, int R2>   // R1 < 0: smooth, i.e. blur(k2) only
__global__ void __launch_bounds__(GS_BW) pre_gauss_strip_kernel(const T* __restrict__ frames, int H, int W, GaussTaps taps,
                                                                float* __restrict__ out) {
    constexpr int R = R2 > R1 ? R2 : R1;
    constexpr int N2 = R + R2 + 1, N1 = R1 >= 0 ? R + R1 + 1 : 1, NIT = GS_SH + 2 * R;
    __shared__ float buf[2][GS_BW + 2 * R];
    const int t = threadIdx.x, x0 = blockIdx.x * GS_BW, y0 = blockIdx.y * GS_SH;
    const long long fe = (long long)H * W;
    const T* fr = frames + (long long)blockIdx.z * fe;
    float* o = out + (long long)blockIdx.z * fe;
    const int gxa = reflect_clamp(x0 - R + t, W), gxb = reflect_clamp(x0 - R + GS_BW + t, W);
    const bool halo = t < 2 * R;
    float pa, pb = 0.f;
    {
        const T* row = fr + (long long)reflect_clamp(y0 - R, H) * W;
        pa = (float)row[gxa];
        if (halo) pb = (float)row[gxb];
    }
    float q2[N2], q1[N1];   // q[d] = row sum of the input row d iterations ago
#pragma unroll
    for (int d = 0; d < N2; ++d) q2[d] = 0.f;
#pragma unroll
    for (int d = 0; d < N1; ++d) q1[d] = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        float* b = buf[it & 1];
        b[t] = pa;
        if (halo) b[GS_BW + t] = pb;
        if (it + 1 < NIT) {   // next input row in flight while this one is filtered
            const T* row = fr + (long long)reflect_clamp(y0 - R + it + 1, H) * W;
            pa = (float)row[gxa];
            if (halo) pb = (float)row[gxb];
        }
        __syncthreads();
        const float* c = b + t + R;
#pragma unroll
        for (int d = N2 - 1; d > 0; --d) q2[d] = q2[d - 1];
        float a2 = 0.f;
#pragma unroll
        for (int j = -R2; j <= R2; ++j) a2 = __fadd_rn(a2, __fmul_rn(taps.k2[j + R2], c[j]));
        q2[0] = a2;
        if (R1 >= 0) {
#pragma unroll
            for (int d = N1 - 1; d > 0; --d) q1[d] = q1[d - 1];
            float a1 = 0.f;
#pragma unroll
            for (int j = -R1; j <= R1; ++j) a1 = __fadd_rn(a1, __fmul_rn(taps.k1[j + (R1 >= 0 ? R1 : 0)], c[j]));
            q1[0] = a1;
        }
        if (it >= 2 * R) {
            const int yo = y0 + it - 2 * R;   // needs input rows yo - r .. yo + r = delays R + r .. R - r
            float v2 = 0.f;
#pragma unroll
            for (int j = -R2; j <= R2; ++j) v2 = __fadd_rn(v2, __fmul_rn(taps.k2[j + R2], q2[R - j]));
            float v = v2;
            if (R1 >= 0) {
                float v1 = 0.f;
#pragma unroll
                for (int j = -R1; j <= R1; ++j) v1 = __fadd_rn(v1, __fmul_rn(taps.k1[j + (R1 >= 0 ? R1 : 0)], q1[(R1 >= 0 ? R - j : 0)]));
                v = __fsub_rn(v2, v1);
            }
            if (yo < H && x0 + t < W) o[(long long)yo * W + x0 + t] = v;
        }
    }
}

}  // namespace b2piv

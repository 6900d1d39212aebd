// mask.cuh - pyorc's velocimetry mask stack and result packing on device-resident result fields.
//
// The PIV engine leaves v_x, v_y, corr, s2n as float32 [time][y][x] in HBM; pyorc then runs up to eleven xarray passes
// over them (pyorc/api/mask.py:147-403) and packs them to int16 for storage (pyorc/const.py:80-83).  Everything here is
// HBM-bound element / small-stencil / time-reduction work: coalesced along x, grids sized to the SM count, float32
// arithmetic in the reference's operation order (no contraction: explicit __f*_rn) so that a mask is decided on the same
// bits as numpy decides it.  Masks are uint8 (1 = keep), [time][y][x] or [y][x].
//
//   minmax     mask.py:147-161     s_min < sqrt(vx^2 + vy^2) < s_max
//   angle      mask.py:163-186     |atan2(vx, vy) - expected| < tolerance
//   count      mask.py:188-201     #valid(t) > tolerance * T                                   -> [y][x]
//   corr, s2n  mask.py:203-225     field > tolerance
//   outliers   mask.py:227-252     |(v - mean_t) / std_t| < tolerance, per component, or / and
//   variance   mask.py:254-285     |std_t / max(mean_t, 1e30)| < tolerance (the reference's clamp) -> [y][x]
//   rolling    mask.py:287-303     s > tolerance * max(s over a centred time window), NaN as 0
//   window_nan / window_mean / window_replace   mask.py:305-403 + helpers.stack_window (helpers.py:638-679)
//   where      mask.py:131-144     field = mask ? field : NaN
//   encode / decode int16          const.py:80-83 (scale_factor 0.01, _FillValue -9999), xarray's CF coder
//   rotate_u_v helpers.py:602-630  (to_ugrid, api/velocimetry.py:284-289)
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace b2piv {

__device__ __forceinline__ float speed_rn(float vx, float vy) {
    return __fsqrt_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)));
}

// ---- element-wise masks ----------------------------------------------------------------------------------------------
// OP 0: minmax, 1: angle, 2: threshold of a single field (corr / s2n)
template <int OP>
__global__ void __launch_bounds__(256) mask_elem_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float p0,
                                                        float p1, unsigned char* __restrict__ m) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        bool keep;
        if (OP == 0) {
            const float s = speed_rn(a[i], b[i]);
            keep = (s > p0) && (s < p1);
        } else if (OP == 1) {
            const float ang = atan2f(a[i], b[i]);
            keep = fabsf(__fsub_rn(ang, p0)) < p1;
        } else {
            keep = a[i] > p0;
        }
        m[i] = keep ? 1 : 0;
    }
}

// ---- time statistics (skipna): count, mean, std (ddof = 0) per location ------------------------------------------------
// One thread per location (coalesced along x), SEQUENTIAL float32 sums over time - the order numpy uses for a reduction
// over the leading axis - two passes like np.nanvar (mean first, then squared deviations).  MS_U loads in flight.
constexpr int MS_U = 8;
__global__ void __launch_bounds__(128) time_stats_kernel(const float* __restrict__ f, int T, long long nxy, int* __restrict__ count,
                                                         float* __restrict__ mean, float* __restrict__ stdv) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nxy; i += stride) {
        float sum = 0.f;
        int cnt = 0;
        int t = 0;
        for (; t + MS_U <= T; t += MS_U) {
            float v[MS_U];
#pragma unroll
            for (int k = 0; k < MS_U; ++k) v[k] = f[(long long)(t + k) * nxy + i];
#pragma unroll
            for (int k = 0; k < MS_U; ++k) {
                const bool ok = v[k] == v[k];
                sum = __fadd_rn(sum, ok ? v[k] : 0.f);
                cnt += ok ? 1 : 0;
            }
        }
        for (; t < T; ++t) {
            const float v = f[(long long)t * nxy + i];
            const bool ok = v == v;
            sum = __fadd_rn(sum, ok ? v : 0.f);
            cnt += ok ? 1 : 0;
        }
        const float avg = __fdiv_rn(sum, (float)cnt);          // 0 / 0 = NaN for an all-NaN location, like numpy
        if (count) count[i] = cnt;
        if (mean) mean[i] = avg;
        if (stdv) {
            float sq = 0.f;
            t = 0;
            for (; t + MS_U <= T; t += MS_U) {
                float v[MS_U];
#pragma unroll
                for (int k = 0; k < MS_U; ++k) v[k] = f[(long long)(t + k) * nxy + i];
#pragma unroll
                for (int k = 0; k < MS_U; ++k) {
                    const float d = (v[k] == v[k]) ? __fsub_rn(v[k], avg) : 0.f;
                    sq = __fadd_rn(sq, __fmul_rn(d, d));
                }
            }
            for (; t < T; ++t) {
                const float v = f[(long long)t * nxy + i];
                const float d = (v == v) ? __fsub_rn(v, avg) : 0.f;
                sq = __fadd_rn(sq, __fmul_rn(d, d));
            }
            stdv[i] = cnt > 0 ? __fsqrt_rn(__fdiv_rn(sq, (float)cnt)) : CUDART_NAN_F;
        }
    }
}

__global__ void __launch_bounds__(256) mask_count_kernel(const int* __restrict__ count, long long nxy, double thr,
                                                         unsigned char* __restrict__ m) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nxy; i += stride) m[i] = ((double)count[i] > thr) ? 1 : 0;
}

__global__ void __launch_bounds__(256) mask_outliers_kernel(const float* __restrict__ vx, const float* __restrict__ vy, int T, long long nxy,
                                                            const float* __restrict__ xm, const float* __restrict__ xs,
                                                            const float* __restrict__ ym, const float* __restrict__ ys, float tol,
                                                            int mode_and, unsigned char* __restrict__ m) {
    const long long n = (long long)T * nxy, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long j = i % nxy;
        const bool xc = fabsf(__fdiv_rn(__fsub_rn(vx[i], xm[j]), xs[j])) < tol;
        const bool yc = fabsf(__fdiv_rn(__fsub_rn(vy[i], ym[j]), ys[j])) < tol;
        m[i] = (mode_and ? (xc && yc) : (xc || yc)) ? 1 : 0;
    }
}

// np.maximum propagates NaN
__device__ __forceinline__ float np_maximum(float a, float b) { return (a != a || b != b) ? CUDART_NAN_F : fmaxf(a, b); }

__global__ void __launch_bounds__(256) mask_variance_kernel(long long nxy, const float* __restrict__ xm, const float* __restrict__ xs,
                                                            const float* __restrict__ ym, const float* __restrict__ ys, float tol,
                                                            int mode_and, unsigned char* __restrict__ m) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nxy; i += stride) {
        const bool xc = fabsf(__fdiv_rn(xs[i], np_maximum(xm[i], 1e30f))) < tol;
        const bool yc = fabsf(__fdiv_rn(ys[i], np_maximum(ym[i], 1e30f))) < tol;
        m[i] = (mode_and ? (xc && yc) : (xc || yc)) ? 1 : 0;
    }
}

// rolling: window of label t covers [t - w/2, t + w - 1 - w/2]; a window that leaves the axis gives NaN -> False
__global__ void __launch_bounds__(256) mask_rolling_kernel(const float* __restrict__ vx, const float* __restrict__ vy, int T, long long nxy,
                                                           int wdw, float tol, unsigned char* __restrict__ m) {
    const long long n = (long long)T * nxy, stride = (long long)gridDim.x * blockDim.x;
    const int lo = wdw / 2, hi = wdw - 1 - wdw / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int t = (int)(i / nxy);
        const long long j = i - (long long)t * nxy;
        bool keep = false;
        if (t - lo >= 0 && t + hi < T) {
            float mx = 0.f;   // NaN counts as 0 and speeds are >= 0
            for (int q = t - lo; q <= t + hi; ++q) {
                const float s = speed_rn(vx[(long long)q * nxy + j], vy[(long long)q * nxy + j]);
                mx = (s == s) ? fmaxf(mx, s) : mx;
            }
            const float s0 = speed_rn(vx[i], vy[i]);
            keep = s0 > __fmul_rn(tol, mx);
        }
        m[i] = keep ? 1 : 0;
    }
}

// ---- spatial windows (helpers.stack_window): strides xs in [wx0, wx1], ys in [wy0, wy1) - the reference's ranges --------
// value of stride (xs, ys) at (y, x) = field[y - ys][x - xs], NaN outside.  Order: xs outer, ys inner (the concat order).
struct WindowArgs {
    int T, ny, nx;
    int wx0, wx1, wy0, wy1;   // wy1 EXCLUSIVE, like the reference
};

__device__ __forceinline__ void window_sum(const float* __restrict__ f, const WindowArgs& w, long long base, int y, int x, float& sum,
                                           int& cnt) {
    sum = 0.f;
    cnt = 0;
    for (int xs = w.wx0; xs <= w.wx1; ++xs) {
        const int xx = x - xs;
        for (int ys = w.wy0; ys < w.wy1; ++ys) {
            const int yy = y - ys;
            if (xx < 0 || xx >= w.nx || yy < 0 || yy >= w.ny) continue;
            const float v = f[base + (long long)yy * w.nx + xx];
            if (v == v) { sum = __fadd_rn(sum, v); ++cnt; }
        }
    }
}

__global__ void __launch_bounds__(256) mask_window_nan_kernel(const float* __restrict__ vx, WindowArgs w, double thr,
                                                              unsigned char* __restrict__ m) {
    const long long nxy = (long long)w.ny * w.nx, n = (long long)w.T * nxy, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long t = i / nxy, j = i - t * nxy;
        float sum;
        int cnt;
        window_sum(vx, w, t * nxy, (int)(j / w.nx), (int)(j % w.nx), sum, cnt);
        m[i] = ((double)cnt >= thr) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) mask_window_mean_kernel(const float* __restrict__ vx, const float* __restrict__ vy, WindowArgs w,
                                                               float tol, int mode_and, unsigned char* __restrict__ m) {
    const long long nxy = (long long)w.ny * w.nx, n = (long long)w.T * nxy, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long t = i / nxy, j = i - t * nxy;
        const int y = (int)(j / w.nx), x = (int)(j % w.nx);
        float sx, sy;
        int cx, cy;
        window_sum(vx, w, t * nxy, y, x, sx, cx);
        window_sum(vy, w, t * nxy, y, x, sy, cy);
        const float mx = __fdiv_rn(sx, (float)cx), my = __fdiv_rn(sy, (float)cy);
        const bool xc = __fdiv_rn(fabsf(__fsub_rn(vx[i], mx)), mx) < tol;
        const bool yc = __fdiv_rn(fabsf(__fsub_rn(vy[i], my)), my) < tol;
        m[i] = (mode_and ? (xc && yc) : (xc || yc)) ? 1 : 0;
    }
}

// one iteration of window_replace for one field (out-of-place: every mean reads the field before this iteration)
__global__ void __launch_bounds__(256) window_replace_kernel(const float* __restrict__ in, WindowArgs w, float* __restrict__ out) {
    const long long nxy = (long long)w.ny * w.nx, n = (long long)w.T * nxy, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = in[i];
        if (v != v) {
            const long long t = i / nxy, j = i - t * nxy;
            float s;
            int c;
            window_sum(in, w, t * nxy, (int)(j / w.nx), (int)(j % w.nx), s, c);
            v = __fdiv_rn(s, (float)c);
        }
        out[i] = v;
    }
}

// ---- where(mask): up to four fields in one pass, mask [T][nxy] or [nxy] (broadcast over time) ---------------------------
struct Fields4 { float* f[4]; int n; };
__global__ void __launch_bounds__(256) mask_apply_kernel(Fields4 fs, long long n, long long nxy, const unsigned char* __restrict__ m,
                                                         int mask_has_time) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const bool keep = m[mask_has_time ? i : i % nxy] != 0;
        if (!keep) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < fs.n) fs.f[k][i] = CUDART_NAN_F;
        }
    }
}

// ---- f-4: CF packing ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) encode_i16_kernel(const float* __restrict__ a, long long n, float scale, int fill,
                                                         short* __restrict__ q) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = rintf(__fdiv_rn(a[i], scale));          // round half to even, like np.rint
        q[i] = (v != v) ? (short)fill : (short)fminf(fmaxf(v, -32768.f), 32767.f);
    }
}
__global__ void __launch_bounds__(256) decode_i16_kernel(const short* __restrict__ q, long long n, float scale, int fill,
                                                         float* __restrict__ a) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        a[i] = (q[i] == (short)fill) ? CUDART_NAN_F : __fmul_rn((float)q[i], scale);
}
// u2 = c u - s v ; v2 = s u + c v in float64 (numpy promotes float32 fields times a float64 array element)
__global__ void __launch_bounds__(256) rotate_uv_kernel(const float* __restrict__ u, const float* __restrict__ v, long long n, double c,
                                                        double s, double* __restrict__ u2, double* __restrict__ v2) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double a = (double)u[i], b = (double)v[i];
        u2[i] = __dadd_rn(__dmul_rn(c, a), __dmul_rn(-s, b));
        v2[i] = __dadd_rn(__dmul_rn(s, a), __dmul_rn(c, b));
    }
}

}  // namespace b2piv

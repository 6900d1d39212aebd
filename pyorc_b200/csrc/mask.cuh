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
// Pure streams: four consecutive elements per thread and access (16-byte loads, 4-byte mask stores) and two accesses in
// flight, scalar tail / unaligned fall-back (same shape as preproc.cuh's streams).
__device__ __forceinline__ bool mask_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__device__ __forceinline__ bool mask_al4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3) == 0; }

// OP 0: minmax, 1: angle, 2: threshold of a single field (corr / s2n)
template <int OP>
__device__ __forceinline__ unsigned char mask_elem(float a, float b, float p0, float p1) {
    if (OP == 0) {
        const float s = speed_rn(a, b);
        return ((s > p0) && (s < p1)) ? 1 : 0;
    } else if (OP == 1) {
        return (fabsf(__fsub_rn(atan2f(a, b), p0)) < p1) ? 1 : 0;
    }
    return (a > p0) ? 1 : 0;
}
template <int OP>
__global__ void __launch_bounds__(256) mask_elem_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float p0,
                                                        float p1, unsigned char* __restrict__ m) {
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nv = (mask_al16(a) && mask_al16(b) && mask_al4(m)) ? (n & ~3ll) : 0;
    long long i = t0 * 4;
    for (; i + stride * 4 < nv; i += stride * 8) {
        const float4 a0 = *reinterpret_cast<const float4*>(a + i), a1 = *reinterpret_cast<const float4*>(a + i + stride * 4);
        float4 b0 = a0, b1 = a1;
        if (OP != 2) { b0 = *reinterpret_cast<const float4*>(b + i); b1 = *reinterpret_cast<const float4*>(b + i + stride * 4); }
        *reinterpret_cast<uchar4*>(m + i) = make_uchar4(mask_elem<OP>(a0.x, b0.x, p0, p1), mask_elem<OP>(a0.y, b0.y, p0, p1),
                                                        mask_elem<OP>(a0.z, b0.z, p0, p1), mask_elem<OP>(a0.w, b0.w, p0, p1));
        *reinterpret_cast<uchar4*>(m + i + stride * 4) = make_uchar4(mask_elem<OP>(a1.x, b1.x, p0, p1), mask_elem<OP>(a1.y, b1.y, p0, p1),
                                                                     mask_elem<OP>(a1.z, b1.z, p0, p1), mask_elem<OP>(a1.w, b1.w, p0, p1));
    }
    for (; i < nv; i += stride * 4) {
        const float4 a0 = *reinterpret_cast<const float4*>(a + i);
        float4 b0 = a0;
        if (OP != 2) b0 = *reinterpret_cast<const float4*>(b + i);
        *reinterpret_cast<uchar4*>(m + i) = make_uchar4(mask_elem<OP>(a0.x, b0.x, p0, p1), mask_elem<OP>(a0.y, b0.y, p0, p1),
                                                        mask_elem<OP>(a0.z, b0.z, p0, p1), mask_elem<OP>(a0.w, b0.w, p0, p1));
    }
    for (long long j = nv + t0; j < n; j += stride) m[j] = mask_elem<OP>(a[j], OP != 2 ? b[j] : 0.f, p0, p1);
}

// ---- time statistics (skipna): count, mean, std (ddof = 0) per location ------------------------------------------------
// One thread per location (coalesced along x), SEQUENTIAL float32 sums over time - the order numpy uses for a reduction
// over the leading axis - two passes like np.nanvar (mean first, then squared deviations).  MS_U loads in flight.
constexpr int MS_U = 32;
__global__ void __launch_bounds__(64) time_stats_kernel(const float* __restrict__ f, int T, long long nxy, int* __restrict__ count,
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

// Segmented variant for T <= 32 * L (L = 32, i.e. up to 1024 time steps): the SUMS stay sequential in time - bit-identical to the kernel above and
// to numpy - but the LOADS do not: a block owns 32 consecutive locations (lanes, coalesced) and cuts time into S segments
// of L steps, one warp per segment; every thread first pulls its L values into registers (S * L loads in flight per
// location instead of MS_U), then the running sum is handed from warp to warp through shared memory in time order.  The
// squared deviations of the second pass reuse the registers, so the field is read from HBM once (the reference's
// two-pass std reads it twice).  ncu on the one-thread-per-location kernel: 10 % of the warp slots occupied, 98 % of
// the stalls `long_scoreboard`, 1.5 TB/s.
template <int L>
__global__ void __launch_bounds__(1024) time_stats_seg_kernel(const float* __restrict__ f, int T, long long nxy, int* __restrict__ count,
                                                              float* __restrict__ mean, float* __restrict__ stdv) {
    __shared__ float s_acc[32];
    __shared__ int s_cnt[32];
    const int lane = threadIdx.x, seg = threadIdx.y, S = blockDim.y;
    const long long n_blocks = (nxy + 31) / 32;
    for (long long b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const long long loc = b * 32 + lane;
        const bool live = loc < nxy;
        float v[L];
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const int t = seg * L + k;
            v[k] = (live && t < T) ? f[(long long)t * nxy + loc] : CUDART_NAN_F;   // padding is skipped like a NaN sample
        }
        for (int sg = 0; sg < S; ++sg) {       // first pass: sum and count, in time order
            if (seg == sg) {
                float sum = sg == 0 ? 0.f : s_acc[lane];
                int cnt = sg == 0 ? 0 : s_cnt[lane];
#pragma unroll
                for (int k = 0; k < L; ++k) {
                    const bool ok = v[k] == v[k];
                    sum = __fadd_rn(sum, ok ? v[k] : 0.f);
                    cnt += ok ? 1 : 0;
                }
                s_acc[lane] = sum;
                s_cnt[lane] = cnt;
            }
            __syncthreads();
        }
        const int cnt = s_cnt[lane];
        const float avg = __fdiv_rn(s_acc[lane], (float)cnt);
        __syncthreads();
        if (stdv) {
            for (int sg = 0; sg < S; ++sg) {   // second pass: squared deviations from the registers, in time order
                if (seg == sg) {
                    float sq = sg == 0 ? 0.f : s_acc[lane];
#pragma unroll
                    for (int k = 0; k < L; ++k) {
                        const float d = (v[k] == v[k]) ? __fsub_rn(v[k], avg) : 0.f;
                        sq = __fadd_rn(sq, __fmul_rn(d, d));
                    }
                    s_acc[lane] = sq;
                }
                __syncthreads();
            }
        }
        if (seg == 0 && live) {
            if (count) count[loc] = cnt;
            if (mean) mean[loc] = avg;
            if (stdv) stdv[loc] = cnt > 0 ? __fsqrt_rn(__fdiv_rn(s_acc[lane], (float)cnt)) : CUDART_NAN_F;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) mask_count_kernel(const int* __restrict__ count, long long nxy, int min_count,
                                                         unsigned char* __restrict__ m) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nxy; i += stride) m[i] = (count[i] >= min_count) ? 1 : 0;
}

// grid: x over locations (grid-stride), y over time steps (grid-stride) - no 64-bit divisions per element
__global__ void __launch_bounds__(256) mask_outliers_kernel(const float* __restrict__ vx, const float* __restrict__ vy, int T, long long nxy,
                                                            const float* __restrict__ xm, const float* __restrict__ xs,
                                                            const float* __restrict__ ym, const float* __restrict__ ys, float tol,
                                                            int mode_and, unsigned char* __restrict__ m) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nxy; j += stride) {
        const float mx = xm[j], sx = xs[j], my = ym[j], sy = ys[j];
        for (int t = blockIdx.y; t < T; t += gridDim.y) {
            const long long i = (long long)t * nxy + j;
            const bool xc = fabsf(__fdiv_rn(__fsub_rn(vx[i], mx), sx)) < tol;
            const bool yc = fabsf(__fdiv_rn(__fsub_rn(vy[i], my), sy)) < tol;
            m[i] = (mode_and ? (xc && yc) : (xc || yc)) ? 1 : 0;
        }
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

// rolling: window of label t covers [t - w/2, t + w - 1 - w/2]; a window that leaves the axis gives NaN -> False.
// A thread walks down time for one location with the last `wdw` speeds in a register ring (wdw <= ROLL_MAX), so every
// element is read once; larger windows recompute the speeds (served by L1 / L2).
constexpr int ROLL_MAX = 16;
template <int WD>   // WD > 0: compile-time window (register ring); WD = 0: generic
__global__ void __launch_bounds__(128) mask_rolling_kernel(const float* __restrict__ vx, const float* __restrict__ vy, int T, long long nxy,
                                                           int wdw, float tol, int t_chunk, unsigned char* __restrict__ m) {
    const int w = WD > 0 ? WD : wdw;
    const int lo = w / 2, hi = w - 1 - w / 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nxy; j += stride) {
        for (int c = blockIdx.y; c * t_chunk < T; c += gridDim.y) {
            const int t0 = c * t_chunk, t1 = min(T, t0 + t_chunk);
            if constexpr (WD > 0) {
                float ring[WD > 0 ? WD : 1];   // ring[k] = speed (NaN as 0) of time t - lo + k, shifted every step
#pragma unroll
                for (int k = 0; k < WD - 1; ++k) {
                    const int q = t0 - lo + k;
                    float sp = 0.f;
                    if (q >= 0 && q < T) { sp = speed_rn(vx[(long long)q * nxy + j], vy[(long long)q * nxy + j]); sp = (sp == sp) ? sp : 0.f; }
                    ring[k] = sp;
                }
                for (int t = t0; t < t1; ++t) {
                    const int q = t + hi;
                    float sp = 0.f;
                    if (q < T) { sp = speed_rn(vx[(long long)q * nxy + j], vy[(long long)q * nxy + j]); sp = (sp == sp) ? sp : 0.f; }
                    ring[WD - 1] = sp;
                    float mx = 0.f;
#pragma unroll
                    for (int k = 0; k < WD; ++k) mx = fmaxf(mx, ring[k]);
                    const long long i = (long long)t * nxy + j;
                    const float s0 = speed_rn(vx[i], vy[i]);     // own sample again: NaN must stay NaN here (L1 hit)
                    m[i] = (t - lo >= 0 && t + hi < T && s0 > __fmul_rn(tol, mx)) ? 1 : 0;
#pragma unroll
                    for (int k = 0; k < WD - 1; ++k) ring[k] = ring[k + 1];
                }
            } else {
                for (int t = t0; t < t1; ++t) {
                    const long long i = (long long)t * nxy + j;
                    bool keep = false;
                    if (t - lo >= 0 && t + hi < T) {
                        float mx = 0.f;   // NaN counts as 0 and speeds are >= 0
                        for (int q = t - lo; q <= t + hi; ++q) {
                            const float sp = speed_rn(vx[(long long)q * nxy + j], vy[(long long)q * nxy + j]);
                            mx = (sp == sp) ? fmaxf(mx, sp) : mx;
                        }
                        keep = speed_rn(vx[i], vy[i]) > __fmul_rn(tol, mx);
                    }
                    m[i] = keep ? 1 : 0;
                }
            }
        }
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
    // strides that land inside the field: xs in [x - nx + 1, x], ys in [y - ny + 1, y] - clipped once, the loops are then
    // free of conditions (order unchanged: xs outer ascending, ys inner ascending)
    const int xs0 = max(w.wx0, x - w.nx + 1), xs1 = min(w.wx1, x);
    const int ys0 = max(w.wy0, y - w.ny + 1), ys1 = min(w.wy1 - 1, y);
    const float* p = f + base + (long long)y * w.nx + x;
    for (int xs = xs0; xs <= xs1; ++xs) {
        const float* q = p - xs - ys0 * w.nx;
        for (int ys = ys0; ys <= ys1; ++ys, q -= w.nx) {
            const float v = *q;
            if (v == v) { sum = __fadd_rn(sum, v); ++cnt; }
        }
    }
}

// Compile-time neighbourhood (the reference's defaults wdw = 1 and wdw = 2): fully unrolled, 32-bit offsets, predicated loads.
// ncu on the run-time version: 293 instructions per warp and output, 85 % issue-bound - the loop machinery, not memory.
template <int WX0, int WX1, int WY0, int WY1>
__device__ __forceinline__ void window_sum_ct(const float* __restrict__ p, int ny, int nx, int y, int x, float& sum, int& cnt) {
    sum = 0.f;
    cnt = 0;
#pragma unroll
    for (int xs = WX0; xs <= WX1; ++xs) {
#pragma unroll
        for (int ys = WY0; ys < WY1; ++ys) {
            const int xx = x - xs, yy = y - ys;
            const bool in = (unsigned)xx < (unsigned)nx && (unsigned)yy < (unsigned)ny;
            const float v = in ? p[-ys * nx - xs] : CUDART_NAN_F;
            if (v == v) { sum = __fadd_rn(sum, v); ++cnt; }
        }
    }
}
// dispatch on the neighbourhood: 0 = run time, 1 = (-1, 1, -1, 1), 2 = (-2, 2, -2, 2)
__device__ __forceinline__ void window_sum_any(int kind, const float* __restrict__ f, const WindowArgs& w, long long base, int y, int x,
                                               float& sum, int& cnt) {
    if (kind == 1) window_sum_ct<-1, 1, -1, 1>(f + base + y * w.nx + x, w.ny, w.nx, y, x, sum, cnt);
    else if (kind == 2) window_sum_ct<-2, 2, -2, 2>(f + base + y * w.nx + x, w.ny, w.nx, y, x, sum, cnt);
    else window_sum(f, w, base, y, x, sum, cnt);
}
__host__ __device__ inline int window_kind(const WindowArgs& w) {
    if (w.wx0 == -1 && w.wx1 == 1 && w.wy0 == -1 && w.wy1 == 1) return 1;
    if (w.wx0 == -2 && w.wx1 == 2 && w.wy0 == -2 && w.wy1 == 2) return 2;
    return 0;
}

// blocks of 32 x 8 outputs, grid.z strided over time: neighbours come from L1
#define B2_WINDOW_LOOP                                                                     \
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;         \
    const long long nxy = (long long)w.ny * w.nx;                                          \
    if (x >= w.nx || y >= w.ny) return;                                                    \
    for (int t = blockIdx.z; t < w.T; t += gridDim.z)

__global__ void __launch_bounds__(256) mask_window_nan_kernel(const float* __restrict__ vx, WindowArgs w, int min_count,
                                                              unsigned char* __restrict__ m) {
    const int kind = window_kind(w);
    B2_WINDOW_LOOP {
        float sum;
        int cnt;
        window_sum_any(kind, vx, w, t * nxy, y, x, sum, cnt);
        m[t * nxy + (long long)y * w.nx + x] = (cnt >= min_count) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) mask_window_mean_kernel(const float* __restrict__ vx, const float* __restrict__ vy, WindowArgs w,
                                                               float tol, int mode_and, unsigned char* __restrict__ m) {
    const int kind = window_kind(w);
    B2_WINDOW_LOOP {
        const long long i = t * nxy + (long long)y * w.nx + x;
        float sx, sy;
        int cx, cy;
        window_sum_any(kind, vx, w, t * nxy, y, x, sx, cx);
        window_sum_any(kind, vy, w, t * nxy, y, x, sy, cy);
        const float mx = __fdiv_rn(sx, (float)cx), my = __fdiv_rn(sy, (float)cy);
        const bool xc = __fdiv_rn(fabsf(__fsub_rn(vx[i], mx)), mx) < tol;
        const bool yc = __fdiv_rn(fabsf(__fsub_rn(vy[i], my)), my) < tol;
        m[i] = (mode_and ? (xc && yc) : (xc || yc)) ? 1 : 0;
    }
}

// one iteration of window_replace for one field (out-of-place: every mean reads the field before this iteration)
__global__ void __launch_bounds__(256) window_replace_kernel(const float* __restrict__ in, WindowArgs w, float* __restrict__ out) {
    const int kind = window_kind(w);
    B2_WINDOW_LOOP {
        const long long i = t * nxy + (long long)y * w.nx + x;
        float v = in[i];
        if (v != v) {
            float s;
            int c;
            window_sum_any(kind, in, w, t * nxy, y, x, s, c);
            v = __fdiv_rn(s, (float)c);
        }
        out[i] = v;
    }
}

// ---- where(mask): up to four fields in one pass, mask [T][nxy] or [nxy] (broadcast over time) ---------------------------
struct Fields4 { float* f[4]; int n; };
__global__ void __launch_bounds__(256) mask_apply_kernel(Fields4 fs, int T, long long nxy, const unsigned char* __restrict__ m,
                                                         int mask_has_time) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nxy; j += stride) {
        const bool keep_xy = mask_has_time ? true : (m[j] != 0);
        if (!mask_has_time && keep_xy) continue;
        for (int t = blockIdx.y; t < T; t += gridDim.y) {
            const long long i = (long long)t * nxy + j;
            if (mask_has_time && m[i] != 0) continue;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < fs.n) fs.f[k][i] = CUDART_NAN_F;
        }
    }
}

// ---- f-4: CF packing ------------------------------------------------------------------------------------------------
__device__ __forceinline__ short encode_one(float a, float scale, int fill) {
    const float v = rintf(__fdiv_rn(a, scale));          // round half to even, like np.rint
    return (v != v) ? (short)fill : (short)fminf(fmaxf(v, -32768.f), 32767.f);
}
__global__ void __launch_bounds__(256) encode_i16_kernel(const float* __restrict__ a, long long n, float scale, int fill,
                                                         short* __restrict__ q) {
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nv = (mask_al16(a) && (reinterpret_cast<uintptr_t>(q) & 7) == 0) ? (n & ~3ll) : 0;
    for (long long i = t0 * 4; i < nv; i += stride * 4) {
        const float4 v = *reinterpret_cast<const float4*>(a + i);
        *reinterpret_cast<short4*>(q + i) = make_short4(encode_one(v.x, scale, fill), encode_one(v.y, scale, fill), encode_one(v.z, scale, fill),
                                                        encode_one(v.w, scale, fill));
    }
    for (long long j = nv + t0; j < n; j += stride) q[j] = encode_one(a[j], scale, fill);
}
__device__ __forceinline__ float decode_one(short q, float scale, int fill) {
    return (q == (short)fill) ? CUDART_NAN_F : __fmul_rn((float)q, scale);
}
__global__ void __launch_bounds__(256) decode_i16_kernel(const short* __restrict__ q, long long n, float scale, int fill,
                                                         float* __restrict__ a) {
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nv = (mask_al16(a) && (reinterpret_cast<uintptr_t>(q) & 7) == 0) ? (n & ~3ll) : 0;
    for (long long i = t0 * 4; i < nv; i += stride * 4) {
        const short4 v = *reinterpret_cast<const short4*>(q + i);
        *reinterpret_cast<float4*>(a + i) = make_float4(decode_one(v.x, scale, fill), decode_one(v.y, scale, fill), decode_one(v.z, scale, fill),
                                                        decode_one(v.w, scale, fill));
    }
    for (long long j = nv + t0; j < n; j += stride) a[j] = decode_one(q[j], scale, fill);
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

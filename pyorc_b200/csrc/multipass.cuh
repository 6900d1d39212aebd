// multipass.cuh - the glue of the two-pass scheme of BASELINE.json configs[2] ("2-pass deform").
//
// There is no reference implementation (ffpiv is single pass), so the scheme is DEFINED in DESIGN.md §8 (and restated on the CPU for the tests):
// pass 1 (coarse windows) -> universal outlier detection on 3 x 3 neighbourhoods -> bilinear predictor at the fine window
// centres, rounded to whole pixels -> pass 2 with frame k+1's window displaced by the predictor -> predictor + residual.
// Both passes are the ordinary fused correlation kernels; this file holds the two small kernels in between.  They work
// on a few thousand vectors per frame pair, in float64 and in the oracle's operation order (no contraction), so that the
// integer shifts - which decide what pass 2 correlates - are identical to the oracle's for identical pass-1 fields.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace b2piv {

// median as the element of rank (n - 1) / 2 (insertion sort of at most 8 values)
__device__ __forceinline__ double lower_median8(double* a, int n) {
    for (int i = 1; i < n; ++i) {
        const double x = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > x) { a[j + 1] = a[j]; --j; }
        a[j + 1] = x;
    }
    return a[(n - 1) / 2];
}

// Universal outlier detection (Westerweel & Scarano 2005) with replacement by the neighbourhood median.
// in: u, v float32 [n_pairs][rows][cols] (NaN = invalid); out: validated float64 fields, same shape
__global__ void __launch_bounds__(128) mp_validate_kernel(const float* __restrict__ u, const float* __restrict__ v, int n_pairs, int rows,
                                                          int cols, double eps, double thr, double* __restrict__ ou, double* __restrict__ ov) {
    const long long n = (long long)n_pairs * rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cols), r = (int)((i / cols) % rows);
        const long long base = i - (long long)r * cols - c;
        double nu[8], nv[8];
        int m = 0;
        for (int dr = -1; dr <= 1; ++dr)
            for (int dc = -1; dc <= 1; ++dc) {
                const int rr = r + dr, cc = c + dc;
                if ((dr == 0 && dc == 0) || rr < 0 || rr >= rows || cc < 0 || cc >= cols) continue;
                const float a = u[base + (long long)rr * cols + cc], b = v[base + (long long)rr * cols + cc];
                if (isfinite(a) && isfinite(b)) { nu[m] = (double)a; nv[m] = (double)b; ++m; }
            }
        const double cu = (double)u[i], cv = (double)v[i];
        const bool bad = !(isfinite(cu) && isfinite(cv));
        double xu = bad ? 0.0 : cu, xv = bad ? 0.0 : cv;
        if (m > 0) {
            double du[8], dv[8], su[8], sv[8];
            for (int k = 0; k < m; ++k) { su[k] = nu[k]; sv[k] = nv[k]; }
            const double mu = lower_median8(su, m), mv = lower_median8(sv, m);
            for (int k = 0; k < m; ++k) { du[k] = fabs(__dsub_rn(nu[k], mu)); dv[k] = fabs(__dsub_rn(nv[k], mv)); }
            const double ru = __dadd_rn(lower_median8(du, m), eps), rv = __dadd_rn(lower_median8(dv, m), eps);
            if (bad || __ddiv_rn(fabs(__dsub_rn(cu, mu)), ru) > thr || __ddiv_rn(fabs(__dsub_rn(cv, mv)), rv) > thr) { xu = mu; xv = mv; }
        }
        ou[i] = xu; ov[i] = xv;
    }
}

struct MpGrid {
    int rows, cols;      // field shape
    int wy, wx, sy, sx;  // window size and stride
};

// Bilinear predictor at the fine window centres -> whole-pixel shifts (dy, dx), clamped to keep the window in the frame.
// shift: short [n_pairs][rows2 * cols2][2]
__global__ void __launch_bounds__(128) mp_predictor_kernel(const double* __restrict__ u, const double* __restrict__ v, int n_pairs, MpGrid g1,
                                                           MpGrid g2, int H, int W, short* __restrict__ shift) {
    const long long n = (long long)n_pairs * g2.rows * g2.cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % g2.cols), r = (int)((i / g2.cols) % g2.rows);
        const long long k = i / ((long long)g2.cols * g2.rows);
        const int y0 = r * g2.sy, x0 = c * g2.sx;
        double fy = __ddiv_rn(__dsub_rn(__dadd_rn((double)y0, (double)g2.wy / 2.0), (double)g1.wy / 2.0), (double)g1.sy);
        double fx = __ddiv_rn(__dsub_rn(__dadd_rn((double)x0, (double)g2.wx / 2.0), (double)g1.wx / 2.0), (double)g1.sx);
        fy = fmin(fmax(fy, 0.0), (double)(g1.rows - 1));
        fx = fmin(fmax(fx, 0.0), (double)(g1.cols - 1));
        int iy = (int)floor(fy), ix = (int)floor(fx);
        iy = min(iy, max(g1.rows - 2, 0));
        ix = min(ix, max(g1.cols - 2, 0));
        const double ty = __dsub_rn(fy, (double)iy), tx = __dsub_rn(fx, (double)ix);
        const int iy1 = min(iy + 1, g1.rows - 1), ix1 = min(ix + 1, g1.cols - 1);
        const double* fu = u + k * g1.rows * g1.cols;
        const double* fv = v + k * g1.rows * g1.cols;
        double res[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double* f = q == 0 ? fu : fv;
            const double f00 = f[iy * g1.cols + ix], f01 = f[iy * g1.cols + ix1], f10 = f[iy1 * g1.cols + ix], f11 = f[iy1 * g1.cols + ix1];
            const double top = __dadd_rn(f00, __dmul_rn(__dsub_rn(f01, f00), tx));
            const double bot = __dadd_rn(f10, __dmul_rn(__dsub_rn(f11, f10), tx));
            res[q] = __dadd_rn(top, __dmul_rn(__dsub_rn(bot, top), ty));
        }
        long long dx = (long long)rint(res[0]), dy = (long long)rint(res[1]);   // half to even, like np.rint
        dy = dy < -y0 ? -y0 : (dy > H - g2.wy - y0 ? H - g2.wy - y0 : dy);
        dx = dx < -x0 ? -x0 : (dx > W - g2.wx - x0 ? W - g2.wx - x0 : dx);
        shift[2 * i + 0] = (short)dy;
        shift[2 * i + 1] = (short)dx;
    }
}

// ---- window DEFORMATION (BASELINE.json configs[2] "2-pass deform") -----------------------------------------------------------
// The validated pass-1 field is interpolated to EVERY pixel (bilinear between the coarse window centres, edge values outside)
// and frame k+1 is resampled at (y + dv, x + du) (bilinear between the four neighbours, coordinates clamped to the frame):
// B'_k(y, x) ~ frame k+1 warped back onto frame k, so that pass 2 - an ordinary correlation of frame k with B'_k on the fine
// grid - only sees the residual, whatever the velocity GRADIENT inside a window (a discrete offset removes the mean only).
// The arithmetic is float64 in the operation order of the CPU definition (two-pass scheme), rounded once to float32, so the
// warped frames are identical to the definition's.  Output: an interleaved float32 stack [2 P][H][W] = (frame k, B'_k).
__device__ __forceinline__ double mp_interp(const double* __restrict__ f, int cols, int iy, int iy1, int ix, int ix1, double ty, double tx) {
    const double f00 = f[iy * cols + ix], f01 = f[iy * cols + ix1], f10 = f[iy1 * cols + ix], f11 = f[iy1 * cols + ix1];
    const double top = __dadd_rn(f00, __dmul_rn(__dsub_rn(f01, f00), tx));
    const double bot = __dadd_rn(f10, __dmul_rn(__dsub_rn(f11, f10), tx));
    return __dadd_rn(top, __dmul_rn(__dsub_rn(bot, top), ty));
}
// fractional coarse index of position `pos` (pixel-edge coordinates) along one axis, and its cell
__device__ __forceinline__ void mp_cell(double pos, int w1, int s1, int n1, int* i0, int* i1, double* t) {
    double f = __ddiv_rn(__dsub_rn(pos, (double)w1 / 2.0), (double)s1);
    f = fmin(fmax(f, 0.0), (double)(n1 - 1));
    int i = (int)floor(f);
    i = min(i, max(n1 - 2, 0));
    *t = __dsub_rn(f, (double)i);
    *i0 = i;
    *i1 = min(i + 1, n1 - 1);
}
template <typename T>
__global__ void __launch_bounds__(256) mp_deform_kernel(const T* __restrict__ frames, long long frame_stride_el, int pitch_el, int n_pairs, int H,
                                                        int W, const double* __restrict__ u, const double* __restrict__ v, MpGrid g1,
                                                        float* __restrict__ stack) {
    const long long fe = (long long)H * W, n = (long long)n_pairs * fe;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const long long k = i / fe;
        const T* fa = frames + k * frame_stride_el;
        const T* fb = fa + frame_stride_el;
        int iy, iy1, ix, ix1;
        double ty, tx;
        mp_cell((double)y + 0.5, g1.wy, g1.sy, g1.rows, &iy, &iy1, &ty);
        mp_cell((double)x + 0.5, g1.wx, g1.sx, g1.cols, &ix, &ix1, &tx);
        const long long fo = k * g1.rows * g1.cols;
        const double du = mp_interp(u + fo, g1.cols, iy, iy1, ix, ix1, ty, tx);
        const double dv = mp_interp(v + fo, g1.cols, iy, iy1, ix, ix1, ty, tx);
        // sample frame k+1 at (y + dv, x + du), clamped to the frame
        const double yy = fmin(fmax(__dadd_rn((double)y, dv), 0.0), (double)(H - 1)), xx = fmin(fmax(__dadd_rn((double)x, du), 0.0), (double)(W - 1));
        int y0 = (int)floor(yy), x0 = (int)floor(xx);
        y0 = min(y0, max(H - 2, 0)); x0 = min(x0, max(W - 2, 0));
        const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
        const double wy = __dsub_rn(yy, (double)y0), wx = __dsub_rn(xx, (double)x0);
        const double b00 = (double)fb[(long long)y0 * pitch_el + x0], b01 = (double)fb[(long long)y0 * pitch_el + x1];
        const double b10 = (double)fb[(long long)y1 * pitch_el + x0], b11 = (double)fb[(long long)y1 * pitch_el + x1];
        const double top = __dadd_rn(b00, __dmul_rn(__dsub_rn(b01, b00), wx));
        const double bot = __dadd_rn(b10, __dmul_rn(__dsub_rn(b11, b10), wx));
        stack[(2 * k) * fe + (long long)y * W + x] = (float)fa[(long long)y * pitch_el + x];
        stack[(2 * k + 1) * fe + (long long)y * W + x] = (float)__dadd_rn(top, __dmul_rn(__dsub_rn(bot, top), wy));
    }
}
// the same predictor at the centres of the fine windows, NOT rounded: float32 [n_pairs][rows2 * cols2][2] = (dv, du), what the
// residual of pass 2 is added to
__global__ void __launch_bounds__(128) mp_predictor_float_kernel(const double* __restrict__ u, const double* __restrict__ v, int n_pairs, MpGrid g1,
                                                                 MpGrid g2, float* __restrict__ pred) {
    const long long n = (long long)n_pairs * g2.rows * g2.cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % g2.cols), r = (int)((i / g2.cols) % g2.rows);
        const long long k = i / ((long long)g2.cols * g2.rows);
        int iy, iy1, ix, ix1;
        double ty, tx;
        mp_cell(__dadd_rn((double)(r * g2.sy), (double)g2.wy / 2.0), g1.wy, g1.sy, g1.rows, &iy, &iy1, &ty);
        mp_cell(__dadd_rn((double)(c * g2.sx), (double)g2.wx / 2.0), g1.wx, g1.sx, g1.cols, &ix, &ix1, &tx);
        const long long fo = k * g1.rows * g1.cols;
        pred[2 * i + 0] = (float)mp_interp(v + fo, g1.cols, iy, iy1, ix, ix1, ty, tx);
        pred[2 * i + 1] = (float)mp_interp(u + fo, g1.cols, iy, iy1, ix, ix1, ty, tx);
    }
}

}  // namespace b2piv

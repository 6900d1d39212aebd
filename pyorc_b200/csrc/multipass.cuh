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

}  // namespace b2piv

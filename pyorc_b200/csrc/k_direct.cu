// k_direct.cu - any-size windows by direct circular cross-correlation (piv_direct.cuh), one CTA per (pair, window).
#include "engine.h"
#include "piv_direct.cuh"

using namespace b2piv;

__global__ void __launch_bounds__(DNT) piv_direct_kernel(Params p, int wy, int wx, long long n_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DView s = direct_view(smem_raw, wy, wx);
    const int tid = threadIdx.x;
    const int nw = p.n_rows * p.n_cols;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int pair = (int)(item / nw), widx = (int)(item % nw);
        direct_load(s, tid, p, pair, widx);            __syncthreads();
        direct_center(s, tid, p);                      __syncthreads();
        direct_correlate(s, tid, p, pair, widx);       __syncthreads();
        direct_peak(s, tid, p, pair, widx);            __syncthreads();
    }
}

// ensemble variant: CTA owns a window, walks the frame pairs, masked planes summed in registers
__global__ void __launch_bounds__(DNT) piv_direct_ens_kernel(Params p, EnsParams ep, int wy, int wx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DView s = direct_view(smem_raw, wy, wx);
    const int tid = threadIdx.x;
    const int nw = p.n_rows * p.n_cols, npx = wy * wx;
    constexpr int EPT = 16;   // 64*64 / 256
    for (int widx = blockIdx.x; widx < nw; widx += gridDim.x) {
        float acc[EPT];
#pragma unroll
        for (int k = 0; k < EPT; ++k) acc[k] = 0.f;
        float cnt = 0.f;
        for (int pr = 0; pr < p.n_pairs; ++pr) {
            direct_load(s, tid, p, pr, widx);          __syncthreads();
            direct_center(s, tid, p);                  __syncthreads();
            Params q = p; q.planes = nullptr;
            direct_correlate(s, tid, q, pr, widx);     __syncthreads();
            const unsigned long long key = d_tot_max(s, 4);
            float cmax = __uint_as_float((unsigned)(key >> 32));
            float s2n = cmax / (d_tot_sum(s, 5) / (float)npx);
            bool ok = (cmax >= ep.corr_min) && (s2n >= ep.s2n_min) && isfinite(cmax);
            if (p.keep && !p.keep[widx]) ok = false;
            if (ok) {
#pragma unroll
                for (int k = 0; k < EPT; ++k) {
                    const int e = tid + k * DNT;
                    if (e < npx) acc[k] += s.plane[e];
                }
                if (cmax > 1e-6f) cnt += 1.f;
            } else {
                cmax = 0.f; s2n = 0.f;
            }
            if (tid == 0) { p.cmax[(long long)pr * nw + widx] = cmax; p.s2n[(long long)pr * nw + widx] = s2n; }
            __syncthreads();
        }
        float* dst = ep.plane_sum + (long long)widx * npx;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            const int e = tid + k * DNT;
            if (e < npx) dst[e] += acc[k];
        }
        if (tid == 0) ep.count[widx] += cnt;
    }
}

int launch_direct(b2piv_engine* e, const Params& p, cudaStream_t st) {
    const long long n_items = (long long)p.n_rows * p.n_cols * p.n_pairs;
    if (n_items <= 0) return B2PIV_OK;
    const size_t smem = direct_smem_bytes(e->wy, e->wx);
    CK(cudaFuncSetAttribute(piv_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, piv_direct_kernel, DNT, smem));
    if (occ < 1) return fail(e, B2PIV_ERR_CUDA, "direct kernel does not fit on an SM");
    long long grid = (long long)occ * e->sm_count;
    if (grid > n_items) grid = n_items;
    piv_direct_kernel<<<(unsigned)grid, DNT, smem, st>>>(p, e->wy, e->wx, n_items);
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}
int launch_direct_ens(b2piv_engine* e, const Params& p, const EnsParams& ep, cudaStream_t st) {
    const int nw = p.n_rows * p.n_cols;
    if (p.n_pairs <= 0) return B2PIV_OK;
    const size_t smem = direct_smem_bytes(e->wy, e->wx);
    CK(cudaFuncSetAttribute(piv_direct_ens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, piv_direct_ens_kernel, DNT, smem));
    if (occ < 1) return fail(e, B2PIV_ERR_CUDA, "direct kernel does not fit on an SM");
    long long grid = (long long)occ * e->sm_count;
    if (grid > nw) grid = nw;
    piv_direct_ens_kernel<<<(unsigned)grid, DNT, smem, st>>>(p, ep, e->wy, e->wx);
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Large windows (one side above 64 px - pyorc accepts any even window size, pyorc/api/frames.py:159-171; the FFT kernels
// cover the powers of two and everything up to 64 px): the same direct circular correlation, organised for 65 .. 128 px.
//   shared memory: a [wy][wx] and b doubled along x [wy][2 wx] (the x wrap-around disappears, the y one is an index),
//                  190 KB at 126 x 126; the plane goes to a per-CTA scratch in global memory (L2);
//   a thread owns the lags (ly .. ly + 7, lx): walking down y at fixed x, b[(y + ly + k) % wy][x + lx] of step y is
//   b[.. + k - 1] of step y + 1 - a sliding register window, ONE new value per step - and a[y][x] is a broadcast: two
//   shared-memory reads per eight multiply-adds, lanes on consecutive lx (conflict-free).
// O(N^2) per window like piv_direct_kernel: ~0.3 M windows/s at 100 x 100, a compatibility path, not a fast one.
// ------------------------------------------------------------------------------------------------------------
constexpr int DB_K = 8;   // lags per thread along y

static size_t direct_big_smem_host(int wy, int wx) { return (size_t)(3 * wy * wx) * sizeof(float) + 64 * sizeof(unsigned long long) + 16; }

// one (frame pair, window): correlation plane -> scratch (fftshifted, clipped); returns via shared reductions max key (slot 4)
// and sum (slot 5) like direct_correlate
__device__ __forceinline__ void direct_big_plane(DView& s, int tid, const Params& p, int pair, int widx, float* plane) {
    const int wy = s.wy, wx = s.wx;
    // ---- load + sums (b at [y][x] of the doubled rows) ----
    {
        const unsigned char* base = (const unsigned char*)p.frames + (long long)pair * p.frame_stride;
        const int r = widx / p.n_cols, c = widx % p.n_cols;
        const long long off = (long long)(r * p.sy) * p.pitch;
        const int x0 = c * p.sx;
        float fa = 0.f, fb = 0.f;
        for (int e = tid; e < wy * wx; e += DNT) {
            const int y = e / wx, x = e % wx;
            float a, b;
            if (!p.is_f32) {
                const unsigned char* ra = base + off + (long long)y * p.pitch + x0 + x;
                a = (float)ra[0]; b = (float)ra[p.frame_stride];
            } else {
                const float* ra = (const float*)(base + off + (long long)y * p.pitch) + x0 + x;
                a = ra[0];
                b = *(const float*)((const unsigned char*)ra + p.frame_stride);
            }
            s.a[e] = a;
            s.b2[y * 2 * wx + x] = b;
            fa += a; fb += b;
        }
        d_dep_sum(s, tid, 0, fa);
        d_dep_sum(s, tid, 1, fb);
    }
    __syncthreads();
    // ---- centre (+clip), second moments, double b along x ----
    {
        const float n = (float)(wy * wx);
        const float ma = d_tot_sum(s, 0) / n, mb = d_tot_sum(s, 1) / n;
        float qa = 0.f, qb = 0.f;
        for (int e = tid; e < wy * wx; e += DNT) {
            const int y = e / wx, x = e % wx;
            float a = s.a[e] - ma, b = s.b2[y * 2 * wx + x] - mb;
            qa += a * a; qb += b * b;
            if (p.clip_norm) { a = a < 0.f ? 0.f : a; b = b < 0.f ? 0.f : b; }
            s.a[e] = a;
            s.b2[y * 2 * wx + x] = b;
            s.b2[y * 2 * wx + x + wx] = b;
        }
        d_dep_sum(s, tid, 2, qa);
        d_dep_sum(s, tid, 3, qb);
    }
    __syncthreads();
    // ---- correlate ----
    const double n = (double)(wy * wx);
    const double va = (double)d_tot_sum(s, 2) / n, vb = (double)d_tot_sum(s, 3) / n;
    const float scale = (va > 0.0 && vb > 0.0) ? (float)(1.0 / (n * sqrt(va) * sqrt(vb))) : 0.f;
    unsigned long long best = 0ull;
    float sum = 0.f;
    const int n_strips = (wy + DB_K - 1) / DB_K;
    for (int item = tid; item < n_strips * wx; item += DNT) {
        const int strip = item / wx, lx = item % wx;     // lanes of a warp: consecutive lx
        const int ly0 = strip * DB_K;
        float acc[DB_K];
#pragma unroll
        for (int k = 0; k < DB_K; ++k) acc[k] = 0.f;
        if (scale != 0.f) {
            for (int x = 0; x < wx; ++x) {
                const float* bc = s.b2 + x + lx;           // column x + lx of the doubled rows
                const float* ac = s.a + x;
                float w[DB_K], part[DB_K];    // two-level sums (per column, then over the columns): rounding error ~ sqrt(wy) + sqrt(wx)
#pragma unroll
                for (int k = 0; k < DB_K; ++k) part[k] = 0.f;
                int row = ly0 % wy;
#pragma unroll
                for (int k = 0; k < DB_K - 1; ++k) { w[k] = bc[row * 2 * wx]; row = row + 1 == wy ? 0 : row + 1; }
                for (int y = 0; y < wy; ++y) {
                    w[DB_K - 1] = bc[row * 2 * wx];
                    row = row + 1 == wy ? 0 : row + 1;
                    const float av = ac[y * wx];
#pragma unroll
                    for (int k = 0; k < DB_K; ++k) part[k] = fmaf(av, w[k], part[k]);
#pragma unroll
                    for (int k = 0; k < DB_K - 1; ++k) w[k] = w[k + 1];
                }
#pragma unroll
                for (int k = 0; k < DB_K; ++k) acc[k] += part[k];
            }
        }
#pragma unroll
        for (int k = 0; k < DB_K; ++k) {
            const int ly = ly0 + k;
            if (ly < wy) {
                const float v = scale == 0.f ? 0.f : clip01(acc[k] * scale);
                const int i = (ly + wy / 2) % wy, j = (lx + wx / 2) % wx;      // fftshifted position of lag (ly, lx)
                const int e = i * wx + j;
                plane[e] = v;
                sum += v;
                const unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xffffffffu - (unsigned)e);
                best = key > best ? key : best;
                if (p.planes) p.planes[((long long)pair * p.n_rows * p.n_cols + widx) * (wy * wx) + e] = v;
            }
        }
    }
    d_dep_max(s, tid, 4, best);
    d_dep_sum(s, tid, 5, sum);
    __syncthreads();      // reductions complete, the plane (global scratch, written by this CTA) is visible to this CTA
}

__global__ void __launch_bounds__(DNT) piv_direct_big_kernel(Params p, int wy, int wx, long long n_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DView s = direct_view(smem_raw, wy, wx);       // a, b2 (only [wy][2 wx] of it is used), red, scal; s.plane is NOT valid here
    const int tid = threadIdx.x;
    const int nw = p.n_rows * p.n_cols;
    float* plane = p.scratch + (size_t)blockIdx.x * wy * wx;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int pair = (int)(item / nw), widx = (int)(item % nw);
        direct_big_plane(s, tid, p, pair, widx, plane);
        s.plane = plane;                            // direct_peak reads the neighbours of the peak through the view
        direct_peak(s, tid, p, pair, widx);
        __syncthreads();
    }
}

// ensemble variant: a CTA owns a window and walks the frame pairs; a plane that passes the thresholds is added to the window's
// accumulator plane (only this CTA ever touches it)
__global__ void __launch_bounds__(DNT) piv_direct_big_ens_kernel(Params p, EnsParams ep, int wy, int wx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DView s = direct_view(smem_raw, wy, wx);
    const int tid = threadIdx.x;
    const int nw = p.n_rows * p.n_cols, npx = wy * wx;
    float* plane = p.scratch + (size_t)blockIdx.x * npx;
    for (int widx = blockIdx.x; widx < nw; widx += gridDim.x) {
        float cnt = 0.f;
        float* dst = ep.plane_sum + (long long)widx * npx;
        for (int pr = 0; pr < p.n_pairs; ++pr) {
            Params q = p; q.planes = nullptr;
            direct_big_plane(s, tid, q, pr, widx, plane);
            const unsigned long long key = d_tot_max(s, 4);
            float cmax = __uint_as_float((unsigned)(key >> 32));
            float s2n = cmax / (d_tot_sum(s, 5) / (float)npx);
            bool ok = (cmax >= ep.corr_min) && (s2n >= ep.s2n_min) && isfinite(cmax);
            if (p.keep && !p.keep[widx]) ok = false;
            if (ok) {
                for (int e = tid; e < npx; e += DNT) dst[e] += plane[e];
                if (cmax > 1e-6f) cnt += 1.f;
            } else {
                cmax = 0.f; s2n = 0.f;
            }
            if (tid == 0) { p.cmax[(long long)pr * nw + widx] = cmax; p.s2n[(long long)pr * nw + widx] = s2n; }
            __syncthreads();
        }
        if (tid == 0) ep.count[widx] += cnt;
    }
}

static int direct_big_scratch(b2piv_engine* e, Params& p, long long grid) {
    const int rc = ensure(e, &e->d_direct_ws, &e->cap_direct_ws, (size_t)grid * e->wy * e->wx * sizeof(float));
    if (rc) return rc;
    p.scratch = e->d_direct_ws;
    return B2PIV_OK;
}

int launch_direct_big(b2piv_engine* e, const Params& p0, cudaStream_t st) {
    Params p = p0;
    const long long n_items = (long long)p.n_rows * p.n_cols * p.n_pairs;
    if (n_items <= 0) return B2PIV_OK;
    const size_t smem = direct_big_smem_host(e->wy, e->wx);
    CK(cudaFuncSetAttribute(piv_direct_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, piv_direct_big_kernel, DNT, smem));
    if (occ < 1) return fail(e, B2PIV_ERR_CUDA, "large-window direct kernel does not fit on an SM");
    long long grid = (long long)occ * e->sm_count;
    if (grid > n_items) grid = n_items;
    int rc = direct_big_scratch(e, p, grid);
    if (rc) return rc;
    piv_direct_big_kernel<<<(unsigned)grid, DNT, smem, st>>>(p, e->wy, e->wx, n_items);
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}
int launch_direct_big_ens(b2piv_engine* e, const Params& p0, const EnsParams& ep, cudaStream_t st) {
    Params p = p0;
    const int nw = p.n_rows * p.n_cols;
    if (p.n_pairs <= 0) return B2PIV_OK;
    const size_t smem = direct_big_smem_host(e->wy, e->wx);
    CK(cudaFuncSetAttribute(piv_direct_big_ens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, piv_direct_big_ens_kernel, DNT, smem));
    if (occ < 1) return fail(e, B2PIV_ERR_CUDA, "large-window direct kernel does not fit on an SM");
    long long grid = (long long)occ * e->sm_count;
    if (grid > nw) grid = nw;
    int rc = direct_big_scratch(e, p, grid);
    if (rc) return rc;
    piv_direct_big_ens_kernel<<<(unsigned)grid, DNT, smem, st>>>(p, ep, e->wy, e->wx);
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}

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

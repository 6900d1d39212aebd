// k_generic_ens.cu - ensemble variant of the shared-memory FFT kernel: a CTA owns a window pair and walks the frame pairs.
#include "engine.h"

using namespace b2piv;

template <class C>
__global__ void __launch_bounds__(C::NT) piv_ens_kernel(Params p, EnsParams ep, const float2* __restrict__ twx,
                                                        const float2* __restrict__ twy, int n_witems) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<C>& s = *reinterpret_cast<Smem<C>*>(smem_raw);
    const int tid = threadIdx.x;
    constexpr int EPT = C::NPX / C::NT;
    phase_init<C>(s, tid, twx, twy);
    __syncthreads();
    const int nw = p.n_rows * p.n_cols;
    for (int wi = blockIdx.x; wi < n_witems; wi += gridDim.x) {
        float acc[C::NWIN][EPT];
        float cnt[C::NWIN];
#pragma unroll
        for (int w = 0; w < C::NWIN; ++w) {
            cnt[w] = 0.f;
#pragma unroll
            for (int k = 0; k < EPT; ++k) acc[w][k] = 0.f;
        }
        Item it = decode_item<C>(p, wi);  // pair 0 of this window item
        for (int pr = 0; pr < p.n_pairs; ++pr) {
            it.pair = pr;
            phase_load<C>(s, tid, p, it);            __syncthreads();
            phase_stats<C>(s, tid, p);               __syncthreads();
            phase_center<C>(s, tid, p);              __syncthreads();
            phase_stats_f32<C>(s, tid, p);
            if (C::PADDED) { phase_embed<C>(s, tid, p); __syncthreads(); }
            fft_pass<C, C::NWIN, 0, 0, 0>(s, tid);   __syncthreads();
            fft_pass<C, C::NWIN, 0, 1, 0>(s, tid);   __syncthreads();
            fft_pass<C, C::NWIN, 1, 0, 0>(s, tid);   __syncthreads();
            fft_pass<C, C::NWIN, 1, 1, 0>(s, tid);   __syncthreads();
            phase_cross<C>(s, tid);                  __syncthreads();
            fft_pass<C, 1, 1, 1, 1>(s, tid);         __syncthreads();
            fft_pass<C, 1, 1, 0, 1>(s, tid);         __syncthreads();
            fft_pass<C, 1, 0, 1, 1>(s, tid);         __syncthreads();
            fft_pass<C, 1, 0, 0, 1>(s, tid);         __syncthreads();
            phase_reduce<C>(s, tid, p, it);          __syncthreads();
#pragma unroll
            for (int w = 0; w < C::NWIN; ++w) {
                if (w == 1 && !it.valid1) continue;
                const unsigned long long key = total_max_u64<C>(s, 2 * w + 0);
                float cmax = __uint_as_float((unsigned)(key >> 32));
                float s2n = cmax / (total_sum_f32<C>(s, 2 * w + 1) / (float)(win_ny<C>(p) * win_nx<C>(p)));
                bool ok = (cmax >= ep.corr_min) && (s2n >= ep.s2n_min) && isfinite(cmax) && (s.scale[w] != 0.f);
                if (p.keep && !p.keep[it.w[w]]) ok = false;   // NaN plane in the reference -> masked out
                if (ok) {
#pragma unroll
                    for (int k = 0; k < EPT; ++k) {
                        const int e = tid + k * C::NT;
                        if (e < win_ny<C>(p) * win_nx<C>(p)) acc[w][k] += shifted_value<C>(s, w, e / win_nx<C>(p), e % win_nx<C>(p), win_ny<C>(p), win_nx<C>(p));
                    }
                    if (cmax > 1e-6f) cnt[w] += 1.f;
                } else {
                    cmax = 0.f; s2n = 0.f;
                }
                if (tid == 0) {
                    const long long o = (long long)pr * nw + it.w[w];
                    p.cmax[o] = cmax; p.s2n[o] = s2n;
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int w = 0; w < C::NWIN; ++w) {
            if (w == 1 && !it.valid1) continue;
            float* dst = ep.plane_sum + (long long)it.w[w] * (win_ny<C>(p) * win_nx<C>(p));
#pragma unroll
            for (int k = 0; k < EPT; ++k)
                if (tid + k * C::NT < win_ny<C>(p) * win_nx<C>(p)) dst[tid + k * C::NT] += acc[w][k];
            if (tid == 0) ep.count[it.w[w]] += cnt[w];
        }
    }
}

template <class C>
static int launch_ens(b2piv_engine* e, const Params& p, const EnsParams& ep, cudaStream_t st) {
    const int nw = p.n_rows * p.n_cols;
    const int n_witems = (C::NWIN == 2) ? (nw + 1) / 2 : nw;
    if (p.n_pairs <= 0) return B2PIV_OK;
    const size_t smem = sizeof(Smem<C>);
    auto kern = piv_ens_kernel<C>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NT, smem));
    if (occ < 1) return fail(e, B2PIV_ERR_CUDA, "kernel does not fit on an SM");
    long long grid = (long long)occ * e->sm_count;
    if (grid > n_witems) grid = n_witems;
    kern<<<(unsigned)grid, C::NT, smem, st>>>(p, ep, e->d_twx, e->d_twy, n_witems);
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}

// window shapes compiled in (NT threads; two windows per work item unless the planes do not fit in 227 KB)
#define B2PIV_CONFIGS(X)      \
    X(16, 16, 64, 2)          \
    X(32, 32, 128, 2)         \
    X(64, 64, 256, 2)         \
    X(32, 64, 128, 2)         \
    X(64, 32, 128, 2)         \
    X(64, 128, 256, 2)        \
    X(128, 64, 256, 2)        \
    X(128, 128, 512, 1)

int launch_generic_ens(b2piv_engine* e, const Params& p, const EnsParams& ep, cudaStream_t st) {
    int py, px;
    plane_shape(e, &py, &px);
    const bool padded = !(py == e->wy && px == e->wx);
#define X(Y, XX, T, NW) if (py == Y && px == XX) return padded ? launch_ens<Cfg<Y, XX, T, NW, true>>(e, p, ep, st) : launch_ens<Cfg<Y, XX, T, NW, false>>(e, p, ep, st);
    B2PIV_CONFIGS(X)
#undef X
    return fail(e, B2PIV_ERR_UNSUPPORTED, "window size not compiled in");
}

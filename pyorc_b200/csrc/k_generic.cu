// k_generic.cu - shared-memory FFT kernel (piv_core.cuh): power-of-two planes 16..128, uint8 / float32, optional padding.
#include "engine.h"

using namespace b2piv;

// Per-time-step: persistent CTAs stride over (frame pair, window pair) work items.
template <class C>
__global__ void __launch_bounds__(C::NT) piv_pairs_kernel(Params p, const float2* __restrict__ twx,
                                                          const float2* __restrict__ twy, int n_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<C>& s = *reinterpret_cast<Smem<C>*>(smem_raw);
    const int tid = threadIdx.x;
    phase_init<C>(s, tid, twx, twy);
    __syncthreads();
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const Item it = decode_item<C>(p, item);
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
        phase_peak<C>(s, tid, p, it);            __syncthreads();
    }
}

// ---- kernel dispatch over the compiled window configurations ------------------------------------------------
template <class C>
static int launch_pairs(b2piv_engine* e, const Params& p, cudaStream_t st) {
    const int nw = p.n_rows * p.n_cols;
    const int per_pair = (C::NWIN == 2) ? (nw + 1) / 2 : nw;
    const long long n_items = (long long)per_pair * p.n_pairs;
    if (n_items <= 0) return B2PIV_OK;
    const size_t smem = sizeof(Smem<C>);
    auto kern = piv_pairs_kernel<C>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NT, smem));
    if (occ < 1) return fail(e, B2PIV_ERR_CUDA, "kernel does not fit on an SM");
    long long grid = (long long)occ * e->sm_count;
    if (grid > n_items) grid = n_items;
    kern<<<(unsigned)grid, C::NT, smem, st>>>(p, e->d_twx, e->d_twy, (int)n_items);
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

bool fft_config(int wy, int wx) {
#define X(Y, XX, T, NW) if (wy == Y && wx == XX) return true;
    B2PIV_CONFIGS(X)
#undef X
    return false;
}

// FFT plane for a window size that is not itself a compiled FFT shape: smallest power of two >= 2n per axis (exact
// circular correlation by padding, piv_core.cuh phase_embed); squared up when the rectangular shape is not compiled.
static int pad_pow2(int n) { int w = 16; while (w < 2 * n) w <<= 1; return w; }
void plane_shape(const b2piv_engine* e, int* py, int* px) {
    if (fft_config(e->wy, e->wx)) { *py = e->wy; *px = e->wx; return; }
    int a = pad_pow2(e->wy), b = pad_pow2(e->wx);
    if (!fft_config(a, b)) a = b = (a > b ? a : b);
    *py = a; *px = b;
}

// shared-memory FFT kernel (piv_core.cuh) on the window's own plane or, for sizes that are not a compiled FFT shape, padded
int launch_generic(b2piv_engine* e, const Params& p, cudaStream_t st) {
    e->last_variant = 1;
    int py, px;
    plane_shape(e, &py, &px);
    const bool padded = !(py == e->wy && px == e->wx);
#define X(Y, XX, T, NW) if (py == Y && px == XX) return padded ? launch_pairs<Cfg<Y, XX, T, NW, true>>(e, p, st) : launch_pairs<Cfg<Y, XX, T, NW, false>>(e, p, st);
    B2PIV_CONFIGS(X)
#undef X
    return fail(e, B2PIV_ERR_UNSUPPORTED, "window size not compiled in");
}

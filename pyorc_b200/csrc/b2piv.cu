// b2piv.cu - kernels + C ABI (include/b2piv.h) of the B200-native LSPIV engine.  sm_100a only.
//
// Replaces the ffpiv/rocket-fft CPU path that pyorc/velocimetry/ffpiv.py calls (cross_corr, u_v_displacement)
// with ONE fused kernel per frame-pair batch: window gather -> normalise -> packed complex 2-D FFT -> cross
// spectrum -> inverse FFT -> fftshift,/N,clip -> max / mean / first-argmax -> 3-point Gaussian sub-pixel fit.
#include "../../include/b2piv.h"
#include "piv_core.cuh"
#include "piv_rows.cuh"
#include "piv_rows128.cuh"
#include "piv_direct.cuh"
#include "preproc.cuh"
#include "project.cuh"
#include "mask.cuh"
#include "multipass.cuh"

#include <cuda.h>   // CUtensorMap (types only; the encoder is resolved at run time, libcuda is not linked)

#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace b2piv;

// ------------------------------------------------------------------------------------------------------------
// Kernels
// ------------------------------------------------------------------------------------------------------------
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


// ------------------------------------------------------------------------------------------------------------
// Row-per-thread kernel (piv_rows.cuh): TMA-staged uint8 tiles, register-resident W-point FFTs, forward spectra
// shared between consecutive frame pairs.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, unsigned long long* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((unsigned long long)map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

// G groups of W threads share one CTA and run in LOCKSTEP (CTA-wide barriers): the loop body is >100 KB of
// straight-line code, far beyond the instruction caches, so the warps of an SM should stream the SAME instructions
// (ncu on a one-group-per-CTA version: 30 % of issue slots lost to `no_instructions`).
// ROLLED: the four 1-D FFT passes of a frame share one copy of the unrolled FFT; otherwise two copies (one
// "FFT, transpose, FFT" block executed twice).
// Compute phases run unconditionally (an inactive group - only at the tail of the grid - works on garbage and never
// stores results); only TMA traffic and global stores are predicated, so no shuffle sits in a divergent region.
template <class R, int G, bool ROLLED, bool ALIGNED, bool F32, bool ENS = false, bool PAD = false>
__global__ void __launch_bounds__(R::NT* G) piv_rows_kernel(const __grid_constant__ CUtensorMap tmap, RParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // the swizzled TMA tiles need 1024-byte aligned bases: align by hand (launch adds 1 KB of slack)
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    constexpr int W = R::W;
    const int g = threadIdx.x / R::NT;     // group within the CTA
    const int tid = threadIdx.x % R::NT;   // thread within the group (= row / column slot)
    RSmem<R>& s = reinterpret_cast<RSmem<R>*>(base)[g];
    if (tid == 0) {
        mbar_init(&s.mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity = 0;
    RRegs<R> r;
    r.half_alpha_prev[0] = r.half_alpha_prev[1] = 0.f;
    for (long long ubase = (long long)blockIdx.x * G; ubase < p.n_units; ubase += (long long)gridDim.x * G) {
        int maxn = 0;  // frames of the longest unit of this round (uniform over the CTA)
#pragma unroll
        for (int j = 0; j < G; ++j) {
            if (ubase + j < p.n_units) {
                const RUnit t = decode_unit(p, (int)(ubase + j));
                maxn = max(maxn, t.f1 - t.f0 + 1);
            }
        }
        const bool has_unit = (ubase + g) < p.n_units;
        const RUnit un = decode_unit(p, has_unit ? (int)(ubase + g) : 0);
        const int nfr = has_unit ? un.f1 - un.f0 + 1 : 0;
        constexpr int TILE_BYTES = ALIGNED ? R::TILE : R::TILE_U;
        constexpr int WIN_BYTES = TILE_BYTES / 2;
        const int xa0 = ALIGNED ? un.x0[0] : (un.x0[0] & ~15), xa1 = ALIGNED ? un.x0[1] : (un.x0[1] & ~15);
        const int xoff0 = un.x0[0] - xa0, xoff1 = un.x0[1] - xa1;
        // TMA of the tile(s) a frame starts with: both uint8 windows, or (float32) window 0 - plus window 1 when both
        // fit the buffer; `issue_f32` loads the W/32 swizzled 128-byte-wide boxes of one float32 window
        auto issue_f32 = [&](int w, int frame, int toff) {
#pragma unroll
            for (int h = 0; h < W / 32; ++h)
                tma_load_3d(s.tile() + toff + h * R::FBOX, &tmap, &s.mbar, un.x0[w] + 32 * h, un.y0[w], frame);
        };
        auto issue_frame_start = [&](int frame) {
            fence_proxy_async();
            if constexpr (!F32) {
                mbar_expect_tx(&s.mbar, TILE_BYTES);
                tma_load_3d(s.tile(), &tmap, &s.mbar, xa0, un.y0[0], frame);
                tma_load_3d(s.tile() + WIN_BYTES, &tmap, &s.mbar, xa1, un.y0[1], frame);
            } else {
                mbar_expect_tx(&s.mbar, R::F_PHASES == 2 ? R::FWIN : 2 * R::FWIN);
                issue_f32(0, frame, 0);
                if (R::F_PHASES == 1) issue_f32(1, frame, R::FWIN);
            }
        };
        if (has_unit && tid == 0) issue_frame_start(un.f0);
        for (int k = 0; k < maxn; ++k) {
            const bool active = k < nfr;
            const bool have_prev = k > 0;
            const int f = un.f0 + k;
            if (active) {
                while (!mbar_try_wait(&s.mbar, parity)) {}
                parity ^= 1u;
            }
            if constexpr (PAD) {
                static_assert(!PAD || (!ALIGNED && !F32), "padded mode uses the 16-byte wider uint8 boxes");
                rows_p1_pad<R>(s, r, tid, p, xoff0, xoff1);
                __syncthreads();  // A
                rows_p2_pre_pad<R>(s, r, tid, p);
            } else if constexpr (!F32) {
                rows_p1<R, ALIGNED>(s, r, tid, xoff0, xoff1);
                __syncthreads();  // A: integer moments visible, tile (aliased on X) fully consumed
                rows_p2_pre<R>(s, r, tid, p.clip_norm);
            } else {
                rows_f1<R>(s, r, tid, 0, 0);
                if (R::F_PHASES == 1) rows_f1<R>(s, r, tid, 1, R::FWIN);
                __syncthreads();  // A: tile consumed, row sums visible
                if (R::F_PHASES == 2) {
                    if (active && tid == 0) {   // window 1 of this frame into the same buffer
                        fence_proxy_async();
                        mbar_expect_tx(&s.mbar, R::FWIN);
                        issue_f32(1, f, 0);
                    }
                    rows_f2<R>(s, r, tid, 0);   // overlaps the TMA round trip
                    if (active) {
                        while (!mbar_try_wait(&s.mbar, parity)) {}
                        parity ^= 1u;
                    }
                    rows_f1<R>(s, r, tid, 1, 0);
                    __syncthreads();  // A2: tile consumed again
                    rows_f2<R>(s, r, tid, 1);
                } else {
                    rows_f2<R>(s, r, tid, 0);
                    rows_f2<R>(s, r, tid, 1);
                }
                __syncthreads();  // A3: centred second moments visible
                rows_f3<R>(s, r, tid, p.clip_norm);
            }
            // The first frame of a unit has no previous spectra: it still runs the whole pipeline (on whatever the
            // park buffer holds) and simply stores no result - one wasted inverse transform per ~26 frames buys a loop
            // body without data-dependent branches, so no shuffle needs convergence bookkeeping.
            // FFT(rows) T FFT(cols) | cross | FFT(cols) T FFT(rows): the transpose T is its own inverse and leaves the
            // registers in natural order, so the sequence is two identical halves (or four identical FFTs).
            if (ROLLED) {
#pragma unroll 1
                for (int st = 0; st < 4; ++st) {
                    fft_reg<W, 0>(r.v);
                    if ((st & 1) == 0) transpose_device<R>(s, r, tid, st != 0);
                    else if (st == 1) rows_p3b_device<R, PAD>(s, r, tid, true, &p);
                }
            } else {
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    fft_reg<W, 0>(r.v);
                    transpose_device<R>(s, r, tid, half != 0);
                    fft_reg<W, 0>(r.v);
                    if (half == 0) rows_p3b_device<R, PAD>(s, r, tid, true, &p);
                }
            }
            const bool dead0 = (r.half_alpha_prev[0] == 0.f) || (r.half_alpha_new[0] == 0.f);
            const bool dead1 = (r.half_alpha_prev[1] == 0.f) || (r.half_alpha_new[1] == 0.f);
            rows_p5_post<R, PAD>(s, r, tid, dead0, dead1, &p);
            __syncthreads();  // E1: block max / sum; X (and the tile aliased on it) is free again
            if (active && tid == 0 && k + 1 < nfr) issue_frame_start(f + 1);
            if constexpr (ENS) {
                rows_ens<R, PAD>(s, r, tid, p, un, f - 1, active && have_prev);   // thresholds + accumulate; no peak search per pair
            } else {
                rows_p6<R, PAD>(s, r, tid, &p);
                __syncthreads();  // E2: first-argmax key
                if (active && have_prev) rows_dump_planes<R, PAD>(r, tid, p, un, f - 1);
                rows_p7<R, PAD>(s, r, tid, &p);
                __syncthreads();  // F: neighbour rows dumped
                if (active && have_prev) rows_p8<R, PAD>(s, r, tid, p, un, f - 1);
            }
            r.half_alpha_prev[0] = r.half_alpha_new[0];
            r.half_alpha_prev[1] = r.half_alpha_new[1];
        }
        __syncthreads();  // round boundary: the next round's first TMA overwrites X
    }
}

// Displaced second pass (two-pass scheme, multipass.cuh) on the row-per-thread machinery: per frame the DISPLACED windows
// (byte-granular TMA boxes, funnel-shifted rows) are transformed and crossed with the parked spectra of the previous frame's
// undisplaced windows, then the undisplaced windows of this frame are transformed and parked - 1.5 complex FFTs per
// window and pair instead of 1.0, still no window stack and no correlation plane in HBM.  G groups per CTA in lockstep; one
// copy of the FFT in a rolled stage loop (stages 0-3: displaced tile -> cross -> inverse -> peak; 4-5: undisplaced -> park).
template <class R, int G>
__global__ void __launch_bounds__(R::NT* G) piv_rows_shift_kernel(const __grid_constant__ CUtensorMap tmap, RParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    constexpr int W = R::W;
    constexpr int WIN_BYTES = R::TILE_U / 2;
    const int g = threadIdx.x / R::NT;
    const int tid = threadIdx.x % R::NT;
    RShiftSmem<R>& ss = reinterpret_cast<RShiftSmem<R>*>(base)[g];
    RSmem<R>& s = ss.base;
    if (tid == 0) {
        mbar_init(&s.mbar, 1);
        mbar_init(&ss.mbar_d, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t par_u = 0, par_d = 0;
    RRegs<R> r;
    r.half_alpha_prev[0] = r.half_alpha_prev[1] = 0.f;
    const long long nw = (long long)p.n_rows * p.n_cols;
    for (long long ubase = (long long)blockIdx.x * G; ubase < p.n_units; ubase += (long long)gridDim.x * G) {
        int maxn = 0;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            if (ubase + j < p.n_units) {
                const RUnit t = decode_unit(p, (int)(ubase + j));
                maxn = max(maxn, t.f1 - t.f0 + 1);
            }
        }
        const bool has_unit = (ubase + g) < p.n_units;
        const RUnit un = decode_unit(p, has_unit ? (int)(ubase + g) : 0);
        const int nfr = has_unit ? un.f1 - un.f0 + 1 : 0;
        const int xa0 = un.x0[0] & ~15, xa1 = un.x0[1] & ~15;
        const int xoff_u0 = un.x0[0] - xa0, xoff_u1 = un.x0[1] - xa1;
        auto issue_u = [&](int frame) {
            fence_proxy_async();
            mbar_expect_tx(&s.mbar, R::TILE_U);
            tma_load_3d(ss.tile_u, &tmap, &s.mbar, xa0, un.y0[0], frame);
            tma_load_3d(ss.tile_u + WIN_BYTES, &tmap, &s.mbar, xa1, un.y0[1], frame);
        };
        // displaced tile of `frame` = `b` windows of pair frame - 1
        auto issue_d = [&](int frame) {
            const short* sh0 = p.shift + 2 * ((long long)(frame - 1) * nw + un.w[0]);
            const short* sh1 = p.shift + 2 * ((long long)(frame - 1) * nw + un.w[1]);
            fence_proxy_async();
            mbar_expect_tx(&ss.mbar_d, R::TILE_U);
            tma_load_3d(ss.tile_d, &tmap, &ss.mbar_d, (un.x0[0] + sh0[1]) & ~15, un.y0[0] + sh0[0], frame);
            tma_load_3d(ss.tile_d + WIN_BYTES, &tmap, &ss.mbar_d, (un.x0[1] + sh1[1]) & ~15, un.y0[1] + sh1[0], frame);
        };
        if (has_unit && tid == 0) {
            issue_u(un.f0);
            if (nfr > 1) issue_d(un.f0 + 1);
        }
        for (int k = 0; k < maxn; ++k) {
            const bool active = k < nfr;
            const int f = un.f0 + k;
#pragma unroll 1
            for (int stg = (k > 0 ? 0 : 4); stg < 6; ++stg) {
                if (stg == 0) {
                    int xo0 = 0, xo1 = 0;
                    if (active) {
                        xo0 = (un.x0[0] + p.shift[2 * ((long long)(f - 1) * nw + un.w[0]) + 1]) & 15;
                        xo1 = (un.x0[1] + p.shift[2 * ((long long)(f - 1) * nw + un.w[1]) + 1]) & 15;
                        while (!mbar_try_wait(&ss.mbar_d, par_d)) {}
                        par_d ^= 1u;
                    }
                    rows_p1_shift<R>(s, r, tid, ss.tile_d, xo0, xo1);
                    __syncthreads();  // A: integer moments visible, displaced tile consumed
                    if (active && tid == 0 && k + 1 < nfr) issue_d(f + 1);
                    rows_p2_pre<R>(s, r, tid, p.clip_norm);
                } else if (stg == 4) {
                    if (active) {
                        while (!mbar_try_wait(&s.mbar, par_u)) {}
                        par_u ^= 1u;
                    }
                    __syncthreads();  // the reductions of the previous stage (s.red) have been read by everyone
                    rows_p1<R, false>(s, r, tid, xoff_u0, xoff_u1, ss.tile_u);
                    __syncthreads();  // A': moments visible, undisplaced tile consumed
                    if (active && tid == 0 && k + 1 < nfr) issue_u(f + 1);
                    rows_p2_pre<R>(s, r, tid, p.clip_norm);
                }
                fft_reg<W, 0>(r.v);
                if ((stg & 1) == 0) {
                    transpose_device<R>(s, r, tid, true);
                } else if (stg == 1) {
                    rows_cross_only_device<R>(s, r, tid);
                } else if (stg == 3) {
                    const bool dead0 = (r.half_alpha_prev[0] == 0.f) || (r.half_alpha_new[0] == 0.f);
                    const bool dead1 = (r.half_alpha_prev[1] == 0.f) || (r.half_alpha_new[1] == 0.f);
                    rows_p5_post<R, false>(s, r, tid, dead0, dead1, &p);
                    __syncthreads();  // E1
                    rows_p6<R, false>(s, r, tid, &p);
                    __syncthreads();  // E2
                    if (active) rows_dump_planes<R, false>(r, tid, p, un, f - 1);
                    rows_p7<R, false>(s, r, tid, &p);
                    __syncthreads();  // F
                    if (active) rows_p8<R, false>(s, r, tid, p, un, f - 1);
                } else {   // stg == 5
                    rows_park_only_device<R>(s, r, tid);
                    r.half_alpha_prev[0] = r.half_alpha_new[0];
                    r.half_alpha_prev[1] = r.half_alpha_new[1];
                }
            }
        }
        __syncthreads();
    }
}

// 128 x 128 windows: four polyphase sub-groups of 64 threads run the 64 x 64 pipeline above and meet in the cross-spectrum
// phase (piv_rows128.cuh).  One CTA = one group of 256 threads = one pair of adjacent windows followed through a run of
// frames; 213 KB of shared memory (4 x transpose blocks + parked spectra), one CTA per SM.
__global__ void __launch_bounds__(256, 1) piv_rows128_kernel(const __grid_constant__ CUtensorMap tmap, RParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    R128Smem& s = *reinterpret_cast<R128Smem*>(base);
    const int tid = threadIdx.x;
    const int sub = tid >> 6;      // polyphase component (p1, p2) = (sub >> 1, sub & 1)
    const int t = tid & 63;        // thread within the sub-group (= line slot of the 64 x 64 pipeline)
    RSmem<R6>& ss = s.sub[sub];
    if (tid == 0) {
        mbar_init(&s.mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity = 0;
    RRegs<R6> r;
    r.half_alpha_prev[0] = r.half_alpha_prev[1] = 0.f;
    for (long long unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const RUnit un = decode_unit(p, (int)unit);
        const int nfr = un.f1 - un.f0 + 1;
        auto issue_frame = [&](int frame) {
            fence_proxy_async();
            mbar_expect_tx(&s.mbar, 2 * 128 * 128);
#pragma unroll
            for (int w = 0; w < 2; ++w)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    tma_load_3d(s.sub[j].tile() + w * 4096, &tmap, &s.mbar, un.x0[w], un.y0[w] + 32 * j, frame);
        };
        if (tid == 0) issue_frame(un.f0);
        for (int k = 0; k < nfr; ++k) {
            const bool have_prev = k > 0;
            const int f = un.f0 + k;
            while (!mbar_try_wait(&s.mbar, parity)) {}
            parity ^= 1u;
            r128_p1(s, r, sub, t);
            __syncthreads();  // A: integer moments visible, tile (aliased on the transpose blocks) fully consumed
            r128_p2(s, r, p.clip_norm);
            // forward: FFT(rows) T FFT(cols) per component; cross spectra across components; inverse: FFT(cols) T FFT(rows)
            // (one copy of the unrolled FFT: the loop body is far beyond the instruction caches, every KB counts)
#pragma unroll 1
            for (int stg = 0; stg < 4; ++stg) {
                fft_reg<64, 0>(r.v);
                if ((stg & 1) == 0) transpose_device<R6>(ss, r, t, stg != 0);
                else if (stg == 1) r128_cross(s, r, sub, t);
            }
            const bool dead0 = (r.half_alpha_prev[0] == 0.f) || (r.half_alpha_new[0] == 0.f);
            const bool dead1 = (r.half_alpha_prev[1] == 0.f) || (r.half_alpha_new[1] == 0.f);
            rows_p5_post<R6, false>(ss, r, t, dead0, dead1, &p);
            __syncthreads();  // E1: block max / sum of all components; the transpose blocks are free again
            if (tid == 0 && k + 1 < nfr) issue_frame(f + 1);
            r128_p6(s, r, sub, t);
            __syncthreads();  // E2: first-argmax keys
            if (have_prev) r128_dump_planes(r, sub, t, p, un, f - 1);
            r128_p7(s, r, sub, t);
            __syncthreads();  // F: neighbour rows dumped
            if (have_prev) r128_p8(s, r, tid, p, un, f - 1);
            r.half_alpha_prev[0] = r.half_alpha_new[0];
            r.half_alpha_prev[1] = r.half_alpha_new[1];
        }
        __syncthreads();  // unit boundary: the next unit's first TMA overwrites the transpose blocks
    }
}

// Ensemble accumulate: one CTA owns a window pair and walks over all frame pairs of the chunk; the masked
// correlation planes are summed in REGISTERS and added to the HBM accumulator once per launch
// (pyorc/velocimetry/ffpiv.py:200-243 thresholds, :361-363 accumulation).
struct EnsParams {
    float corr_min, s2n_min;
    float* plane_sum;   // [n_windows][WY][WX] fftshifted coordinates
    float* count;       // [n_windows]
};

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


// ------------------------------------------------------------------------------------------------------------
// Any-size windows (piv_direct.cuh): direct circular cross-correlation, one CTA per (pair, window).
// ------------------------------------------------------------------------------------------------------------
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

// Ensemble finish: count filter -> mean plane -> first-argmax + Gaussian (ffpiv.py:280-282, :324). One CTA/window.
__global__ void __launch_bounds__(256) ens_finish_kernel(const float* __restrict__ plane_sum, const float* __restrict__ count,
                                                         int wy, int wx, float min_count, int border_nan, float eps,
                                                         float* __restrict__ u, float* __restrict__ v) {
    __shared__ unsigned long long red[8];
    const int w = blockIdx.x, tid = threadIdx.x;
    const float cnt = count[w];
    const float* pl = plane_sum + (long long)w * wy * wx;
    const bool dead = !(cnt >= min_count);          // corr_sum[count < min] = nan
    unsigned long long best = 0ull;
    bool anynan = false;
    for (int e = tid; e < wy * wx; e += 256) {
        const float val = pl[e] / cnt;              // 0/0 -> NaN like np.divide
        if (isnan(val)) anynan = true;
        const unsigned long long key = ((unsigned long long)__float_as_uint(val < 0.f ? 0.f : val) << 32) |
                                       (unsigned long long)(0xffffffffu - (unsigned)e);
        if (!isnan(val)) best = key > best ? key : best;
    }
    anynan = __syncthreads_or(anynan);
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((tid & 31) == 0) red[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < 8; ++i) best = red[i] > best ? red[i] : best;
        float uu, vv;
        if (dead || anynan) {
            uu = vv = nanf("");
        } else {
            const int idx = (int)(0xffffffffu - (unsigned)(best & 0xffffffffull));
            const int pi = idx / wx, pj = idx % wx;
            if (pi == 0 || pi == wy - 1 || pj == 0 || pj == wx - 1) {
                if (border_nan) uu = vv = nanf("");
                else { uu = (float)(pj - wx / 2); vv = (float)(pi - wy / 2); }
            } else {
                const float lc = logf(pl[pi * wx + pj] / cnt + eps);
                const float ll = logf(pl[(pi - 1) * wx + pj] / cnt + eps), lr = logf(pl[(pi + 1) * wx + pj] / cnt + eps);
                const float ld = logf(pl[pi * wx + pj - 1] / cnt + eps), lu = logf(pl[pi * wx + pj + 1] / cnt + eps);
                vv = ((float)pi + (ll - lr) / (2.f * ll - 4.f * lc + 2.f * lr)) - (float)(wy / 2);
                uu = ((float)pj + (ld - lu) / (2.f * ld - 4.f * lc + 2.f * lu)) - (float)(wx / 2);
            }
        }
        u[w] = uu; v[w] = vv;
    }
}

// signal_threshold: fraction of non-zero pixels of a window over all frames of the call (ffpiv.py:93-97).
__global__ void __launch_bounds__(256) signal_keep_kernel(const unsigned char* __restrict__ frames, long long frame_stride,
                                                          int pitch, int is_f32, int n_frames, int n_cols, int wy, int wx,
                                                          int sy, int sx, float thr, unsigned char* __restrict__ keep) {
    __shared__ unsigned red[8];
    const int w = blockIdx.x, tid = threadIdx.x;
    const int r = w / n_cols, c = w % n_cols;
    unsigned cnt = 0;
    for (int f = 0; f < n_frames; ++f) {
        const unsigned char* base = frames + f * frame_stride + (long long)(r * sy) * pitch;
        for (int e = tid; e < wy * wx; e += 256) {
            const int y = e / wx, x = c * sx + e % wx;
            if (is_f32) cnt += (((const float*)(base + (long long)y * pitch))[x] != 0.f);
            else        cnt += (base[(long long)y * pitch + x] != 0);
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((tid & 31) == 0) red[tid >> 5] = cnt;
    __syncthreads();
    if (tid == 0) {
        unsigned t = 0;
        for (int i = 0; i < 8; ++i) t += red[i];
        const double score = (double)t / ((double)n_frames * wy * wx);
        keep[w] = score >= (double)thr ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Engine
// ------------------------------------------------------------------------------------------------------------
// Worker threads that copy ordinary (pageable) host frames into the engine's page-locked staging buffers.  pyorc hands
// `frame_chunk.values` - plain numpy memory - to the engine (pyorc/velocimetry/ffpiv.py:223,451); a cudaMemcpyAsync from
// pageable memory is staged by the driver on ONE thread (measured: 11 GB/s, 18.7 ms per 100-pair 1080p step against 4.1 ms
// from pinned memory), so the staging is done here, sliced over a few threads, one chunk ahead of the H2D copy.
class CopyPool {
public:
    explicit CopyPool(int n) {
        for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { run(i); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // rows x row_bytes from src (pitch spitch) to dst (pitch dpitch), split by rows over the workers; returns when done
    void copy2d(unsigned char* dst, size_t dpitch, const unsigned char* src, size_t spitch, size_t row_bytes, size_t rows) {
        std::unique_lock<std::mutex> lk(m_);
        dst_ = dst; src_ = src; dpitch_ = dpitch; spitch_ = spitch; row_bytes_ = row_bytes; rows_ = rows;
        pending_ = (int)workers_.size();
        ++gen_;
        cv_.notify_all();
        done_.wait(lk, [this] { return pending_ == 0; });
    }
    int size() const { return (int)workers_.size(); }

private:
    void run(int idx) {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
            if (stop_) return;
            seen = gen_;
            const size_t n = workers_.size(), per = (rows_ + n - 1) / n;
            const size_t r0 = per * idx < rows_ ? per * idx : rows_, r1 = r0 + per < rows_ ? r0 + per : rows_;
            unsigned char* d = dst_; const unsigned char* sp = src_;
            const size_t dp = dpitch_, spp = spitch_, rb = row_bytes_;
            lk.unlock();
            if (dp == rb && spp == rb) {
                if (r1 > r0) memcpy(d + r0 * rb, sp + r0 * rb, (r1 - r0) * rb);
            } else {
                for (size_t r = r0; r < r1; ++r) memcpy(d + r * dp, sp + r * spp, rb);
            }
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    bool stop_ = false;
    unsigned long long gen_ = 0;
    int pending_ = 0;
    unsigned char* dst_ = nullptr; const unsigned char* src_ = nullptr;
    size_t dpitch_ = 0, spitch_ = 0, row_bytes_ = 0, rows_ = 0;
};

struct b2piv_engine {
    int device = 0;
    std::string err;
    // options
    int clip_norm = 0, border_nan = 1, copy_chunks = 0;   // copy_chunks = 0: auto (about 10 MB of frames per H2D chunk)
      // clip_norm = 0 is what ffpiv does (pinned, tests/test_golden.py)
    int variant = 0;    // 0: auto, 1: generic shared-memory FFT kernel, 2: row-per-thread TMA kernel (error if
                        // ineligible), 3: direct any-size kernel
    int run_len = 0;    // frame pairs per work unit of the rows kernel (0: auto)
    int groups = 0;     // window-pair groups per CTA of the rows kernel (0: default)
    int rolled = -1;    // rows kernel: 1 = one shared FFT body (rolled stage loop), 0 = two copies, -1 = default per size
    int last_variant = 0;
    float gauss_eps = 1e-7f;
    // plan
    bool planned = false;
    int H = 0, W = 0, wy = 0, wx = 0, oy = 0, ox = 0, dtype = 0, n_rows = 0, n_cols = 0;
    float2 *d_twx = nullptr, *d_twy = nullptr;   // point into tw_cache
    std::map<int, float2*> tw_cache;              // transform length -> exp(-2 pi i j / n) table on the device
    int sm_count = 0;
    // streams / events
    cudaStream_t s_copy = nullptr, s_comp = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;
    std::vector<cudaEvent_t> ev_chunk;
    // device workspace for *_host calls
    unsigned char* d_frames = nullptr; size_t cap_frames = 0;
    // page-locked staging ring + copy threads for pageable host frames (pipeline_host)
    unsigned char* h_stage[3] = {nullptr, nullptr, nullptr}; size_t cap_stage = 0;
    cudaEvent_t ev_stage[3] = {nullptr, nullptr, nullptr};
    CopyPool* pool = nullptr;
    int stage_threads = 0;   // 0: auto (min(8, hardware threads))
    float* d_out = nullptr; size_t cap_out = 0;       // 4 result fields
    float* d_planes = nullptr; size_t cap_planes = 0;
    float* d_planes_nat = nullptr; size_t cap_planes_nat = 0;   // padded rows kernel: W x W planes in natural lag order
    unsigned char* d_keep = nullptr; size_t cap_keep = 0;
    // ensemble accumulators
    float* d_pre_mean = nullptr; size_t cap_pre_mean = 0;   // pre-processing workspaces
    unsigned* d_pre_mm = nullptr; size_t cap_pre_mm = 0;
    // orthoprojection plan (CSR gather lists, project.cuh)
    int* d_proj_off = nullptr; size_t cap_proj_off = 0;
    int* d_proj_src = nullptr; size_t cap_proj_src = 0;
    int proj_h = 0, proj_w = 0, proj_out_h = 0, proj_out_w = 0; long long proj_samples = 0;
    PeerOut peer = {};                                       // fused gather over peer memory (b2piv_set_peer_outputs)
    double* d_mp_ws = nullptr; size_t cap_mp_ws = 0;        // two-pass scheme: validated pass-1 fields
    float* d_mask_ws = nullptr; size_t cap_mask_ws = 0;     // mask stack: time statistics / window_replace ping-pong
    float* d_ens_sum = nullptr; float* d_ens_cnt = nullptr; size_t cap_ens = 0, cap_ens_windows = 0; bool ens_open = false;
    // stats
    float last_kernel_ms = 0.f;
    long long launches = 0;
};

static std::string g_create_err;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            e->err = std::string(#call) + ": " + cudaGetErrorString(_e);                           \
            return B2PIV_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

static int fail(b2piv_engine* e, int code, const std::string& msg) {
    e->err = msg;
    return code;
}

template <class T>
static int ensure(b2piv_engine* e, T** ptr, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *ptr) return B2PIV_OK;
    if (*ptr) CK(cudaFree(*ptr));
    *ptr = nullptr; *cap = 0;
    CK(cudaMalloc((void**)ptr, bytes));
    *cap = bytes;
    return B2PIV_OK;
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

static bool fft_config(int wy, int wx) {
#define X(Y, XX, T, NW) if (wy == Y && wx == XX) return true;
    B2PIV_CONFIGS(X)
#undef X
    return false;
}
static bool supported(int wy, int wx) { return fft_config(wy, wx) || (wy >= 4 && wx >= 4 && wy <= 64 && wx <= 64); }

static int launch_direct(b2piv_engine* e, const Params& p, cudaStream_t st) {
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
static int launch_direct_ens(b2piv_engine* e, const Params& p, const EnsParams& ep, cudaStream_t st) {
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


// ---- row-per-thread kernel: tensor map + launch ---------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)ptr;
    }
    return fn;
}

static bool rows_eligible(const b2piv_engine* e, const void* d_frames, long long frame_stride, int pitch) {
    if (e->wy != e->wx || (e->wy != 64 && e->wy != 32)) return false;   // uint8 and float32 frames both qualify
    if ((pitch & 15) || (frame_stride & 15) || (((uintptr_t)d_frames) & 15)) return false;
    // every TMA box must start on a 16-byte boundary in global memory: x strides that are a multiple of 16 use exact
    // swizzled boxes, multiples of 4 (32x32 at 75 % overlap: stride 8) a 16-byte wider box read at an offset
    if ((e->wx - e->ox) & 3) return false;
    return get_encode_tiled() != nullptr;
}

// triage path of the padded rows kernel: W x W planes in natural lag order -> the reference's fftshifted ny x nx planes
__global__ void planes_reorder_kernel(const float* __restrict__ nat, float* __restrict__ out, long long n_planes, int W, int ny, int nx) {
    const long long n = n_planes * ny * nx;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(i % nx), iy = (int)((i / nx) % ny);
        const long long pl = i / ((long long)nx * ny);
        const int hy = ny / 2, hx = nx / 2;
        const int qy = iy < hy ? iy + ny - hy : iy - hy, qx = ix < hx ? ix + nx - hx : ix - hx;   // lag (j + n - n/2) % n
        out[i] = nat[(pl * W + qy) * W + qx];
    }
}

template <class R, int G, bool ROLLED, bool ALIGNED, bool F32, bool ENS = false, bool PAD = false>
static int launch_rows(b2piv_engine* e, const Params& gp, cudaStream_t st, const EnsParams* ep = nullptr) {
    constexpr int W = R::W;
    const int n_frames = gp.n_pairs + 1;
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {(cuuint64_t)e->W, (cuuint64_t)e->H, (cuuint64_t)n_frames};
    const cuuint64_t strides[2] = {(cuuint64_t)gp.pitch, (cuuint64_t)gp.frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)(F32 ? 32 : (ALIGNED ? W : R::WB)), (cuuint32_t)W, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle swz = F32 ? CU_TENSOR_MAP_SWIZZLE_128B
                                       : (!ALIGNED ? CU_TENSOR_MAP_SWIZZLE_NONE : (W == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B));
    const CUresult cr = get_encode_tiled()(&tmap, F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3,
                                           const_cast<void*>(gp.frames), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(e, B2PIV_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)cr));
    RParams p;
    memset(&p, 0, sizeof(p));
    p.frames = (const unsigned char*)gp.frames; p.frame_stride = gp.frame_stride; p.pitch = gp.pitch;
    p.n_rows = gp.n_rows; p.n_cols = gp.n_cols; p.sy = gp.sy; p.sx = gp.sx; p.n_pairs = gp.n_pairs;
    p.clip_norm = gp.clip_norm; p.border_nan = gp.border_nan; p.gauss_eps = gp.gauss_eps; p.keep = gp.keep;
    p.u = gp.u; p.v = gp.v; p.cmax = gp.cmax; p.s2n = gp.s2n; p.planes = gp.planes; p.peer = gp.peer;
    const size_t smem = sizeof(RSmem<R>) * G + 1024;
    if (ENS) { p.corr_min = ep->corr_min; p.s2n_min = ep->s2n_min; p.ens_sum = ep->plane_sum; p.ens_count = ep->count; }
    p.ny = PAD ? e->wy : W; p.nx = PAD ? e->wx : W;
    if (PAD) {   // spectrum factor of the 2 x 2 tiling (piv_rows.cuh, "Padded mode")
        p.pad_scale = (float)(1.0 / ((double)R::NPX * p.ny * p.nx));
        const double two_pi = 6.283185307179586476925286766559;
        for (int k = 0; k <= W / 2; ++k) { const double th = two_pi * (double)((k * p.ny) % W) / W; p.pad_ty[k] = make_float2((float)(1.0 + cos(th)), (float)(-sin(th))); }
        for (int k = 0; k < W; ++k) { const double th = two_pi * (double)((k * p.nx) % W) / W; p.pad_tx[k] = make_float2((float)(1.0 + cos(th)), (float)(-sin(th))); }
        for (int k = 0; k < W / 4; ++k) { const int left = p.nx - 4 * k; p.pad_mask[k] = left >= 4 ? 0xffffffffu : (left <= 0 ? 0u : ((1u << (8 * left)) - 1u)); }
        for (int x = 0; x < W; ++x) p.pad_cm[x] = x < p.nx ? 1.f : 0.f;
        if (gp.planes) {
            const int rcp = ensure(e, &e->d_planes_nat, &e->cap_planes_nat, (size_t)gp.n_pairs * gp.n_rows * gp.n_cols * W * W * sizeof(float));
            if (rcp) return rcp;
            p.planes = e->d_planes_nat;
        }
    }
    auto kern = piv_rows_kernel<R, G, ROLLED, ALIGNED, F32, ENS, PAD>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, R::NT * G, smem));
    if (occ < 1) return fail(e, B2PIV_ERR_CUDA, "rows kernel does not fit on an SM");
    const long long resident = (long long)occ * e->sm_count * G;   // resident work units
    const int nw = gp.n_rows * gp.n_cols, n_wp = (nw + 1) / 2;
    int run = e->run_len;
    if (run <= 0) {  // aim for >= 8 waves of work units; every unit start costs one extra forward transform
        long long chunks = (8 * resident + n_wp - 1) / n_wp;
        if (chunks < 1) chunks = 1;
        run = (int)((gp.n_pairs + chunks - 1) / chunks);
        if (run < 8) run = 8;
    }
    if (run > gp.n_pairs || ENS) run = gp.n_pairs;   // ensemble: one unit owns its windows' accumulators for the whole launch
    p.run_len = run;
    const long long n_units = (long long)n_wp * ((gp.n_pairs + run - 1) / run);
    p.n_units = (int)n_units;
    long long grid = (n_units + G - 1) / G;
    if (grid > (long long)occ * e->sm_count) grid = (long long)occ * e->sm_count;
    kern<<<(unsigned)grid, R::NT * G, smem, st>>>(tmap, p);
    CK(cudaGetLastError());
    e->launches++;
    if (PAD && gp.planes) {
        const long long n_planes = (long long)gp.n_pairs * gp.n_rows * gp.n_cols;
        planes_reorder_kernel<<<e->sm_count * 8, 256, 0, st>>>(e->d_planes_nat, gp.planes, n_planes, W, p.ny, p.nx);
        CK(cudaGetLastError());
        e->launches++;
    }
    return B2PIV_OK;
}

// displaced second pass on the row-per-thread kernel: square 32x32 uint8 windows, x stride a multiple of 4
static bool rows_shift_eligible(const b2piv_engine* e, const void* d_frames, long long frame_stride, int pitch) {
    if (e->wy != 32 || e->wx != 32 || e->dtype != B2PIV_U8) return false;
    if ((pitch & 15) || (frame_stride & 15) || (((uintptr_t)d_frames) & 15)) return false;
    if ((e->wx - e->ox) & 3) return false;
    return get_encode_tiled() != nullptr;
}
static int launch_rows_shift(b2piv_engine* e, const Params& gp, cudaStream_t st) {
    using R = RCfg<32>;
    constexpr int G = 4;
    const int n_frames = gp.n_pairs + 1;
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {(cuuint64_t)e->W, (cuuint64_t)e->H, (cuuint64_t)n_frames};
    const cuuint64_t strides[2] = {(cuuint64_t)gp.pitch, (cuuint64_t)gp.frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)R::WB, (cuuint32_t)R::W, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = get_encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(gp.frames), dims, strides, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(e, B2PIV_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)cr));
    RParams p;
    memset(&p, 0, sizeof(p));
    p.frames = (const unsigned char*)gp.frames; p.frame_stride = gp.frame_stride; p.pitch = gp.pitch;
    p.n_rows = gp.n_rows; p.n_cols = gp.n_cols; p.sy = gp.sy; p.sx = gp.sx; p.n_pairs = gp.n_pairs;
    p.clip_norm = gp.clip_norm; p.border_nan = gp.border_nan; p.gauss_eps = gp.gauss_eps; p.keep = gp.keep;
    p.u = gp.u; p.v = gp.v; p.cmax = gp.cmax; p.s2n = gp.s2n; p.planes = gp.planes; p.peer = gp.peer; p.shift = gp.shift;
    p.ny = p.nx = R::W;
    const size_t smem = sizeof(RShiftSmem<R>) * G + 1024;
    auto kern = piv_rows_shift_kernel<R, G>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, R::NT * G, smem));
    if (occ < 1) return fail(e, B2PIV_ERR_CUDA, "displaced rows kernel does not fit on an SM");
    const long long resident = (long long)occ * e->sm_count * G;
    const int nw = gp.n_rows * gp.n_cols, n_wp = (nw + 1) / 2;
    int run = e->run_len;
    if (run <= 0) {
        long long chunks = (8 * resident + n_wp - 1) / n_wp;
        if (chunks < 1) chunks = 1;
        run = (int)((gp.n_pairs + chunks - 1) / chunks);
        if (run < 8) run = 8;
    }
    if (run > gp.n_pairs) run = gp.n_pairs;
    p.run_len = run;
    const long long n_units = (long long)n_wp * ((gp.n_pairs + run - 1) / run);
    p.n_units = (int)n_units;
    long long grid = (n_units + G - 1) / G;
    if (grid > (long long)occ * e->sm_count) grid = (long long)occ * e->sm_count;
    kern<<<(unsigned)grid, R::NT * G, smem, st>>>(tmap, p);
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}

// 128 x 128 uint8 windows on the polyphase row-per-thread kernel (piv_rows128.cuh)
static bool rows128_eligible(const b2piv_engine* e, const void* d_frames, long long frame_stride, int pitch) {
    if (e->wy != 128 || e->wx != 128 || e->dtype != B2PIV_U8) return false;
    if ((pitch & 15) || (frame_stride & 15) || (((uintptr_t)d_frames) & 15)) return false;
    if ((e->wx - e->ox) & 15) return false;   // swizzled boxes start on 16-byte boundaries
    return get_encode_tiled() != nullptr;
}
static int launch_rows128(b2piv_engine* e, const Params& gp, cudaStream_t st) {
    const int n_frames = gp.n_pairs + 1;
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {(cuuint64_t)e->W, (cuuint64_t)e->H, (cuuint64_t)n_frames};
    const cuuint64_t strides[2] = {(cuuint64_t)gp.pitch, (cuuint64_t)gp.frame_stride};
    const cuuint32_t box[3] = {128, 32, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = get_encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(gp.frames), dims, strides, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(e, B2PIV_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)cr));
    RParams p;
    memset(&p, 0, sizeof(p));
    p.frames = (const unsigned char*)gp.frames; p.frame_stride = gp.frame_stride; p.pitch = gp.pitch;
    p.n_rows = gp.n_rows; p.n_cols = gp.n_cols; p.sy = gp.sy; p.sx = gp.sx; p.n_pairs = gp.n_pairs;
    p.clip_norm = gp.clip_norm; p.border_nan = gp.border_nan; p.gauss_eps = gp.gauss_eps; p.keep = gp.keep;
    p.u = gp.u; p.v = gp.v; p.cmax = gp.cmax; p.s2n = gp.s2n; p.planes = gp.planes; p.peer = gp.peer;
    p.ny = p.nx = 128;
    const size_t smem = sizeof(R128Smem) + 1024;
    CK(cudaFuncSetAttribute(piv_rows128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, piv_rows128_kernel, 256, smem));
    if (occ < 1) return fail(e, B2PIV_ERR_CUDA, "128x128 rows kernel does not fit on an SM");
    const long long resident = (long long)occ * e->sm_count;
    const int nw = gp.n_rows * gp.n_cols, n_wp = (nw + 1) / 2;
    int run = e->run_len;
    if (run <= 0) {  // aim for >= 8 waves of work units; every unit start costs one extra forward transform
        long long chunks = (8 * resident + n_wp - 1) / n_wp;
        if (chunks < 1) chunks = 1;
        run = (int)((gp.n_pairs + chunks - 1) / chunks);
        if (run < 8) run = 8;
    }
    if (run > gp.n_pairs) run = gp.n_pairs;
    p.run_len = run;
    const long long n_units = (long long)n_wp * ((gp.n_pairs + run - 1) / run);
    p.n_units = (int)n_units;
    long long grid = n_units < resident ? n_units : resident;
    piv_rows128_kernel<<<(unsigned)grid, 256, smem, st>>>(tmap, p);
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}

// Padded mode of the row-per-thread kernel: any uint8 window (square or not, any stride) whose larger side is at most
// 32 px, i.e. at most half of a 64 x 64 (or 32 x 32) plane.  Native 32 x 32 / 64 x 64 windows never come here.
static bool pad_eligible(const b2piv_engine* e, const void* d_frames, long long frame_stride, int pitch) {
    const int m = e->wy > e->wx ? e->wy : e->wx;
    if (e->dtype != B2PIV_U8 || 2 * m > 64 || e->wy < 2 || e->wx < 2) return false;
    if ((pitch & 15) || (frame_stride & 15) || (((uintptr_t)d_frames) & 15)) return false;
    return get_encode_tiled() != nullptr;
}
template <bool ENS>
static int launch_rows_pad(b2piv_engine* e, const Params& p, cudaStream_t st, const EnsParams* ep) {
    const int m = e->wy > e->wx ? e->wy : e->wx;
    if (2 * m <= 32) return launch_rows<RCfg<32>, 4, false, false, false, ENS, true>(e, p, st, ep);
    return launch_rows<RCfg<64>, 1, true, false, false, ENS, true>(e, p, st, ep);
}

// FFT plane for a window size that is not itself a compiled FFT shape: smallest power of two >= 2n per axis (exact
// circular correlation by padding, piv_core.cuh phase_embed); squared up when the rectangular shape is not compiled.
static int pad_pow2(int n) { int w = 16; while (w < 2 * n) w <<= 1; return w; }
static void plane_shape(const b2piv_engine* e, int* py, int* px) {
    if (fft_config(e->wy, e->wx)) { *py = e->wy; *px = e->wx; return; }
    int a = pad_pow2(e->wy), b = pad_pow2(e->wx);
    if (!fft_config(a, b)) a = b = (a > b ? a : b);
    *py = a; *px = b;
}

// shared-memory FFT kernel (piv_core.cuh) on the window's own plane or, for sizes that are not a compiled FFT shape, padded
static int launch_generic(b2piv_engine* e, const Params& p, cudaStream_t st) {
    e->last_variant = 1;
    int py, px;
    plane_shape(e, &py, &px);
    const bool padded = !(py == e->wy && px == e->wx);
#define X(Y, XX, T, NW) if (py == Y && px == XX) return padded ? launch_pairs<Cfg<Y, XX, T, NW, true>>(e, p, st) : launch_pairs<Cfg<Y, XX, T, NW, false>>(e, p, st);
    B2PIV_CONFIGS(X)
#undef X
    return fail(e, B2PIV_ERR_UNSUPPORTED, "window size not compiled in");
}

static int dispatch_pairs(b2piv_engine* e, const Params& p, cudaStream_t st) {
    if (p.n_pairs <= 0) return B2PIV_OK;
    // displaced second-pass windows (multipass.cuh) break the frame-to-frame spectrum sharing of the row-per-thread kernels
    if (p.shift) {
        if (e->variant != 1 && rows_shift_eligible(e, p.frames, p.frame_stride, p.pitch)) {
            e->last_variant = 2;
            return launch_rows_shift(e, p, st);
        }
        if (e->variant == 2) return fail(e, B2PIV_ERR_UNSUPPORTED, "displaced rows kernel needs square 32x32 uint8 windows, 16-byte aligned base/pitch and an x stride that is a multiple of 4");
        return launch_generic(e, p, st);
    }
    const bool can_rows = rows_eligible(e, p.frames, p.frame_stride, p.pitch);
    if (e->variant == 2 && !can_rows && e->wy != 128)
        return fail(e, B2PIV_ERR_UNSUPPORTED, "rows kernel needs a square 32/64 window, 16-byte aligned base/pitch and an x stride that is a multiple of 4");
    if (e->variant == 2 && e->wy == 128 && !rows128_eligible(e, p.frames, p.frame_stride, p.pitch))
        return fail(e, B2PIV_ERR_UNSUPPORTED, "128x128 rows kernel needs uint8 frames, 16-byte aligned base/pitch and an x stride that is a multiple of 16");
    if (rows128_eligible(e, p.frames, p.frame_stride, p.pitch) && (e->variant == 0 || e->variant == 2)) {
        e->last_variant = 2;
        return launch_rows128(e, p, st);
    }
    if (can_rows && e->variant != 1) {
        e->last_variant = 2;
        // groups per CTA (lockstep width) and rolled / unrolled stage loop.  Measured on B200 (profiles/r01): 64x64 is
        // fastest with one group per CTA (four 64-thread CTAs per SM) and four specialised FFT copies, 32x32 with four
        // single-warp groups per CTA.
        // Compiled variants (measured on B200, profiles/r01/quick_sweeps.log): 64x64 is fastest with one group per CTA
        // (four 64-thread CTAs per SM) and one shared FFT body, 32x32 with four single-warp groups and two FFT copies.
        // `groups` / `rolled` select the alternatives kept for A/B runs.
        const bool aligned = ((e->wx - e->ox) & 15) == 0;
        if (e->dtype == B2PIV_F32) {
            if (e->wy == 64) return launch_rows<RCfg<64>, 1, true, true, true>(e, p, st);
            return launch_rows<RCfg<32>, 4, false, true, true>(e, p, st);
        }
        if (e->wy == 64) {
            if (e->groups == 4) return aligned ? launch_rows<RCfg<64>, 4, false, true, false>(e, p, st) : launch_rows<RCfg<64>, 4, false, false, false>(e, p, st);
            if (e->rolled == 0) return aligned ? launch_rows<RCfg<64>, 1, false, true, false>(e, p, st) : launch_rows<RCfg<64>, 1, false, false, false>(e, p, st);
            return aligned ? launch_rows<RCfg<64>, 1, true, true, false>(e, p, st) : launch_rows<RCfg<64>, 1, true, false, false>(e, p, st);
        }
        if (e->groups == 1) return aligned ? launch_rows<RCfg<32>, 1, false, true, false>(e, p, st) : launch_rows<RCfg<32>, 1, false, false, false>(e, p, st);
        return aligned ? launch_rows<RCfg<32>, 4, false, true, false>(e, p, st) : launch_rows<RCfg<32>, 4, false, false, false>(e, p, st);
    }
    // sizes that are not a compiled FFT shape: uint8 windows up to 32 px run zero-padded through the row-per-thread kernel
    // (piv_rows.cuh "Padded mode"; variant 4 forces it wherever it applies; measured 3.3x the direct kernel at 10x10 and
    // 2.6x the shared-memory kernel at 26x26).  Otherwise (float32 frames, caller-owned tensors with an odd pitch, larger
    // windows): tiny windows by direct correlation, the rest padded through the shared-memory FFT kernel; variant 3
    // forces the direct kernel
    const bool tiny = !fft_config(e->wy, e->wx) && e->wy * e->wx <= 144;
    if (e->variant == 4 && !pad_eligible(e, p.frames, p.frame_stride, p.pitch))
        return fail(e, B2PIV_ERR_UNSUPPORTED, "padded rows kernel needs uint8 frames, a window of at most 32 px and 16-byte aligned base/pitch");
    if (pad_eligible(e, p.frames, p.frame_stride, p.pitch) && (e->variant == 4 || (e->variant == 0 && !fft_config(e->wy, e->wx)))) {
        e->last_variant = 4;
        return launch_rows_pad<false>(e, p, st, nullptr);
    }
    if ((e->variant == 3 || tiny) && e->wy <= 64 && e->wx <= 64) {
        e->last_variant = 3;
        return launch_direct(e, p, st);
    }
    return launch_generic(e, p, st);
}
static int dispatch_ens(b2piv_engine* e, const Params& p, const EnsParams& ep, cudaStream_t st) {
    if (p.n_pairs <= 0) return B2PIV_OK;
    const bool can_rows = rows_eligible(e, p.frames, p.frame_stride, p.pitch);
    if (e->variant == 2 && !can_rows)
        return fail(e, B2PIV_ERR_UNSUPPORTED, "rows kernel needs a square 32/64 window, 16-byte aligned base/pitch and an x stride that is a multiple of 4");
    if (can_rows && e->variant != 1 && e->variant != 3) {
        e->last_variant = 2;
        const bool aligned = ((e->wx - e->ox) & 15) == 0;
        if (e->dtype == B2PIV_F32) {
            if (e->wy == 64) return launch_rows<RCfg<64>, 1, true, true, true, true>(e, p, st, &ep);
            return launch_rows<RCfg<32>, 4, false, true, true, true>(e, p, st, &ep);
        }
        if (e->wy == 64) return aligned ? launch_rows<RCfg<64>, 1, true, true, false, true>(e, p, st, &ep) : launch_rows<RCfg<64>, 1, true, false, false, true>(e, p, st, &ep);
        return aligned ? launch_rows<RCfg<32>, 4, false, true, false, true>(e, p, st, &ep) : launch_rows<RCfg<32>, 4, false, false, false, true>(e, p, st, &ep);
    }
    const bool tiny = !fft_config(e->wy, e->wx) && e->wy * e->wx <= 144;
    if (e->variant == 4 && !pad_eligible(e, p.frames, p.frame_stride, p.pitch))
        return fail(e, B2PIV_ERR_UNSUPPORTED, "padded rows kernel needs uint8 frames, a window of at most 32 px and 16-byte aligned base/pitch");
    if (pad_eligible(e, p.frames, p.frame_stride, p.pitch) && (e->variant == 4 || (e->variant == 0 && !fft_config(e->wy, e->wx)))) {
        e->last_variant = 4;
        return launch_rows_pad<true>(e, p, st, &ep);
    }
    e->last_variant = 1;
    if ((e->variant == 3 || tiny) && e->wy <= 64 && e->wx <= 64) return launch_direct_ens(e, p, ep, st);
    int py, px;
    plane_shape(e, &py, &px);
    const bool padded = !(py == e->wy && px == e->wx);
#define X(Y, XX, T, NW) if (py == Y && px == XX) return padded ? launch_ens<Cfg<Y, XX, T, NW, true>>(e, p, ep, st) : launch_ens<Cfg<Y, XX, T, NW, false>>(e, p, ep, st);
    B2PIV_CONFIGS(X)
#undef X
    return fail(e, B2PIV_ERR_UNSUPPORTED, "window size not compiled in");
}

static Params base_params(const b2piv_engine* e, const void* d_frames, long long frame_stride, int pitch, int n_pairs) {
    Params p;
    memset(&p, 0, sizeof(p));
    p.frames = d_frames; p.frame_stride = frame_stride; p.pitch = pitch; p.is_f32 = (e->dtype == B2PIV_F32);
    p.n_rows = e->n_rows; p.n_cols = e->n_cols; p.sy = e->wy - e->oy; p.sx = e->wx - e->ox; p.n_pairs = n_pairs;
    p.ny = e->wy; p.nx = e->wx;
    p.clip_norm = e->clip_norm; p.border_nan = e->border_nan; p.gauss_eps = e->gauss_eps;
    return p;
}

static int make_keep(b2piv_engine* e, const void* d_frames, long long frame_stride, int pitch, int n_frames, float thr,
                     cudaStream_t st) {
    const int nw = e->n_rows * e->n_cols;
    int rc = ensure(e, &e->d_keep, &e->cap_keep, (size_t)nw);
    if (rc) return rc;
    signal_keep_kernel<<<nw, 256, 0, st>>>((const unsigned char*)d_frames, frame_stride, pitch, e->dtype == B2PIV_F32, n_frames,
                                           e->n_cols, e->wy, e->wx, e->wy - e->oy, e->wx - e->ox, thr, e->d_keep);
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" {

int b2piv_version(void) { return 100; }

const char* b2piv_last_error(const b2piv_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

int b2piv_create(b2piv_engine** out, int device) {
    if (!out) { g_create_err = "out is NULL"; return B2PIV_ERR_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n <= 0) {
        g_create_err = std::string("no CUDA device available (") + cudaGetErrorString(ce) + "); b2piv has no CPU fallback";
        return B2PIV_ERR_CUDA;
    }
    if (device < 0 || device >= n) { g_create_err = "device index out of range"; return B2PIV_ERR_ARG; }
    cudaDeviceProp prop;
    if ((ce = cudaSetDevice(device)) != cudaSuccess || (ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(ce);
        return B2PIV_ERR_CUDA;
    }
    if (prop.major != 10) {
        g_create_err = "b2piv is built for sm_100a (B200) only; found sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
        return B2PIV_ERR_UNSUPPORTED;
    }
    b2piv_engine* e = new b2piv_engine();
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&e->ev_k0) != cudaSuccess || cudaEventCreate(&e->ev_k1) != cudaSuccess) {
        g_create_err = "stream/event creation failed";
        delete e;
        return B2PIV_ERR_CUDA;
    }
    *out = e;
    return B2PIV_OK;
}

void b2piv_destroy(b2piv_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (auto& kv : e->tw_cache) cudaFree(kv.second);
    cudaFree(e->d_frames); cudaFree(e->d_out); cudaFree(e->d_planes);
    cudaFree(e->d_keep); cudaFree(e->d_ens_sum); cudaFree(e->d_ens_cnt); cudaFree(e->d_pre_mean); cudaFree(e->d_pre_mm); cudaFree(e->d_mask_ws); cudaFree(e->d_mp_ws);
    cudaFree(e->d_proj_off); cudaFree(e->d_proj_src); cudaFree(e->d_planes_nat);
    for (auto ev : e->ev_chunk) cudaEventDestroy(ev);
    for (int i = 0; i < 3; ++i) { if (e->h_stage[i]) cudaFreeHost(e->h_stage[i]); if (e->ev_stage[i]) cudaEventDestroy(e->ev_stage[i]); }
    delete e->pool;
    if (e->ev_k0) cudaEventDestroy(e->ev_k0);
    if (e->ev_k1) cudaEventDestroy(e->ev_k1);
    if (e->s_copy) cudaStreamDestroy(e->s_copy);
    if (e->s_comp) cudaStreamDestroy(e->s_comp);
    delete e;
}

int b2piv_set_option(b2piv_engine* e, const char* name, double value) {
    if (!e || !name) return B2PIV_ERR_ARG;
    const std::string n(name);
    if (n == "clip_normalized") e->clip_norm = value != 0.0;
    else if (n == "border_nan") e->border_nan = value != 0.0;
    else if (n == "gauss_eps") e->gauss_eps = (float)value;
    else if (n == "copy_chunks") e->copy_chunks = value < 0 ? 0 : (int)value;
    else if (n == "stage_threads") { e->stage_threads = value < 0 ? 0 : (int)value; delete e->pool; e->pool = nullptr; }
    else if (n == "kernel_variant") e->variant = (int)value;
    else if (n == "run_len") e->run_len = value < 0 ? 0 : (int)value;
    else if (n == "groups") e->groups = value < 0 ? 0 : (int)value;
    else if (n == "rolled") e->rolled = value < 0 ? -1 : (value != 0.0);
    else return fail(e, B2PIV_ERR_ARG, "unknown option " + n);
    return B2PIV_OK;
}

int b2piv_plan(b2piv_engine* e, int height, int width, int win_y, int win_x, int ovl_y, int ovl_x, int dtype,
               int* n_rows, int* n_cols) {
    if (!e) return B2PIV_ERR_ARG;
    if (dtype != B2PIV_U8 && dtype != B2PIV_F32) return fail(e, B2PIV_ERR_ARG, "dtype must be B2PIV_U8 or B2PIV_F32");
    if (win_y <= 0 || win_x <= 0 || ovl_y < 0 || ovl_x < 0 || ovl_y >= win_y || ovl_x >= win_x)
        return fail(e, B2PIV_ERR_ARG, "need 0 <= overlap < window_size");
    if (height < win_y || width < win_x) return fail(e, B2PIV_ERR_ARG, "frame smaller than the interrogation window");
    if (!supported(win_y, win_x))
        return fail(e, B2PIV_ERR_UNSUPPORTED, "window " + std::to_string(win_y) + "x" + std::to_string(win_x) +
                                                  " not supported (any size 4..64 per axis, or 64x128 / 128x64 / 128x128)");
    CK(cudaSetDevice(e->device));
    e->H = height; e->W = width; e->wy = win_y; e->wx = win_x; e->oy = ovl_y; e->ox = ovl_x; e->dtype = dtype;
    const int old_rows = e->n_rows, old_cols = e->n_cols;
    e->n_rows = (height - win_y) / (win_y - ovl_y) + 1;
    e->n_cols = (width - win_x) / (win_x - ovl_x) + 1;
    if (e->n_rows != old_rows || e->n_cols != old_cols) e->peer = PeerOut{};   // gather buffers were sized for the old field
    // twiddle tables exp(-2 pi i j / N) for the FFT plane (= the window, or its padded power-of-two plane), in double
    int py = win_y, px = win_x;
    plane_shape(e, &py, &px);
    // (one table per transform length, built on first use and kept until the engine is destroyed: re-planning - e.g. the
    // coarse / fine grids of the two-pass scheme, twice per call - must not allocate, free or synchronise)
    for (int axis = 0; axis < 2; ++axis) {
        const int n = axis == 0 ? px : py;
        float2*& slot = e->tw_cache[n];
        if (!slot) {
            std::vector<float2> t(n);
            for (int j = 0; j < n; ++j) t[j] = make_float2((float)cos(2.0 * M_PI * j / n), (float)-sin(2.0 * M_PI * j / n));
            CK(cudaMalloc((void**)&slot, sizeof(float2) * n));
            CK(cudaMemcpy(slot, t.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
        }
        (axis == 0 ? e->d_twx : e->d_twy) = slot;
    }
    e->planned = true;
    e->ens_open = false;
    if (n_rows) *n_rows = e->n_rows;
    if (n_cols) *n_cols = e->n_cols;
    return B2PIV_OK;
}

int b2piv_pairs_device(b2piv_engine* e, const void* d_frames, long long frame_stride_bytes, int pitch_bytes, int n_frames,
                       float signal_threshold, float* d_u, float* d_v, float* d_corr_max, float* d_s2n, void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    if (!e->planned) return fail(e, B2PIV_ERR_STATE, "b2piv_plan has not been called");
    if (!d_frames || !d_u || !d_v || !d_corr_max || !d_s2n) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (n_frames < 2) return fail(e, B2PIV_ERR_ARG, "need at least 2 frames (one pair)");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Params p = base_params(e, d_frames, frame_stride_bytes, pitch_bytes, n_frames - 1);
    if (signal_threshold >= 0.f) {
        int rc = make_keep(e, d_frames, frame_stride_bytes, pitch_bytes, n_frames, signal_threshold, st);
        if (rc) return rc;
        p.keep = e->d_keep;
    }
    p.u = d_u; p.v = d_v; p.cmax = d_corr_max; p.s2n = d_s2n;
    if (e->peer.n) {
        if (e->peer.field % ((long long)e->n_rows * e->n_cols) != 0 ||
            e->peer.pair0 + (n_frames - 1) > e->peer.field / ((long long)e->n_rows * e->n_cols))
            return fail(e, B2PIV_ERR_ARG, "peer gather buffers do not hold this call's pair range (b2piv_set_peer_outputs)");
        p.peer = e->peer;
    }
    return dispatch_pairs(e, p, st);
}

int b2piv_set_peer_outputs(b2piv_engine* e, int n_peers, void* const* peer_bases, long long pairs_total, long long pair_offset) {
    if (!e) return B2PIV_ERR_ARG;
    if (n_peers == 0) { e->peer = PeerOut{}; return B2PIV_OK; }
    if (!e->planned) return fail(e, B2PIV_ERR_STATE, "b2piv_plan has not been called");
    if (n_peers < 0 || n_peers > 8 || !peer_bases) return fail(e, B2PIV_ERR_ARG, "1..8 peer buffers");
    if (pairs_total < 1 || pair_offset < 0 || pair_offset >= pairs_total) return fail(e, B2PIV_ERR_ARG, "bad pair range");
    PeerOut po = {};
    po.n = n_peers;
    for (int r = 0; r < n_peers; ++r) {
        if (!peer_bases[r]) return fail(e, B2PIV_ERR_ARG, "NULL peer buffer");
        po.base[r] = (float*)peer_bases[r];
    }
    po.field = pairs_total * (long long)e->n_rows * e->n_cols;
    po.pair0 = pair_offset;
    e->peer = po;
    return B2PIV_OK;
}

}  // extern "C"

// Row pitch of the engine's own device copy of host frames: rounded up to 16 bytes so that every frame width qualifies
// for the TMA kernels (pyorc's orthorectified frames have arbitrary widths, e.g. 371 px for the Ngwerere example)
static int host_pitch(const b2piv_engine* e) {
    const int row = e->W * (e->dtype == B2PIV_F32 ? 4 : 1);
    return (row + 15) & ~15;
}

// shared H2D pipeline: copy frames chunk-wise on s_copy, call `work(first_pair, n_pairs)` on s_comp per chunk
template <class F>
static int pipeline_host(b2piv_engine* e, const void* frames, int n_frames, bool pipelined, F&& work) {
    const size_t esz = e->dtype == B2PIV_F32 ? 4 : 1;
    const size_t row_bytes = (size_t)e->W * esz, dpitch = (size_t)host_pitch(e);
    const size_t hbytes = (size_t)e->H * row_bytes;      // frame on the host (dense)
    const size_t fbytes = (size_t)e->H * dpitch;         // frame on the device (pitched)
    int rc = ensure(e, &e->d_frames, &e->cap_frames, fbytes * n_frames);
    if (rc) return rc;
    const int n_pairs = n_frames - 1;
    // auto: ~10 MB per chunk, at most 32 chunks (measured on B200, tools/e2e_sweep.py: 100 pairs of 1080p are fastest with
    // 16-25 chunks - the tail after the last H2D is one chunk of compute, and every chunk costs one extra transform per unit)
    int chunks = 1;
    if (pipelined) {
        chunks = e->copy_chunks;
        if (chunks <= 0) {
            const size_t target = (size_t)10 << 20;
            size_t c = (fbytes * (size_t)n_frames + target - 1) / target;
            chunks = (int)(c < 1 ? 1 : (c > 32 ? 32 : c));
        }
    }
    if (chunks > n_pairs) chunks = n_pairs;
    while ((int)e->ev_chunk.size() < chunks) {
        cudaEvent_t ev;
        CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        e->ev_chunk.push_back(ev);
    }
    const int per = (n_pairs + chunks - 1) / chunks;
    // pageable source (plain numpy memory): stage through page-locked buffers with the copy threads, ~8 MB per stage chunk,
    // three buffers so that the host copy of the next stage chunk overlaps the H2D of the previous ones
    cudaPointerAttributes attr;
    bool pageable = true;
    if (cudaPointerGetAttributes(&attr, frames) == cudaSuccess) pageable = (attr.type == cudaMemoryTypeUnregistered);
    else cudaGetLastError();
    size_t stage_frames = 0;
    unsigned stage_no = 0;
    if (pageable) {
        CK(cudaStreamSynchronize(e->s_copy));   // no earlier call may still be reading the staging buffers
        stage_frames = ((size_t)8 << 20) / hbytes;
        if (stage_frames < 1) stage_frames = 1;
        if (stage_frames > (size_t)n_frames) stage_frames = (size_t)n_frames;
        const size_t need_bytes = stage_frames * hbytes;
        if (e->cap_stage < need_bytes) {
            CK(cudaStreamSynchronize(e->s_copy));
            for (int i = 0; i < 3; ++i) {
                if (e->h_stage[i]) { CK(cudaFreeHost(e->h_stage[i])); e->h_stage[i] = nullptr; }
                CK(cudaHostAlloc((void**)&e->h_stage[i], need_bytes, cudaHostAllocDefault));
                if (!e->ev_stage[i]) CK(cudaEventCreateWithFlags(&e->ev_stage[i], cudaEventDisableTiming));
            }
            e->cap_stage = need_bytes;
        }
        if (!e->pool) {
            int nt = e->stage_threads;
            if (nt <= 0) { nt = (int)std::thread::hardware_concurrency(); nt = nt > 8 ? 8 : (nt < 1 ? 1 : nt); }
            try {
                e->pool = new CopyPool(nt);
            } catch (...) {   // no threads to be had: let the driver stage the pageable copy (nothing may throw across the ABI)
                e->pool = nullptr;
                pageable = false;
            }
        }
    }
    // enqueue the H2D copy of frames [f0, f1) on s_copy
    auto h2d = [&](int f0, int f1) -> int {
        const unsigned char* src = (const unsigned char*)frames;
        if (!pageable) {
            if (dpitch == row_bytes)
                CK(cudaMemcpyAsync(e->d_frames + (size_t)f0 * fbytes, src + (size_t)f0 * hbytes, (size_t)(f1 - f0) * fbytes,
                                   cudaMemcpyHostToDevice, e->s_copy));
            else   // frames are contiguous on both sides, so a chunk is one 2-D copy of (frames * H) rows
                CK(cudaMemcpy2DAsync(e->d_frames + (size_t)f0 * fbytes, dpitch, src + (size_t)f0 * hbytes, row_bytes, row_bytes,
                                     (size_t)(f1 - f0) * e->H, cudaMemcpyHostToDevice, e->s_copy));
            return B2PIV_OK;
        }
        for (int a = f0; a < f1; a += (int)stage_frames) {
            const int b = a + (int)stage_frames < f1 ? a + (int)stage_frames : f1;
            const int slot = (int)(stage_no % 3);
            if (stage_no >= 3) CK(cudaEventSynchronize(e->ev_stage[slot]));   // the H2D that last used this buffer is done
            ++stage_no;
            e->pool->copy2d(e->h_stage[slot], row_bytes, src + (size_t)a * hbytes, row_bytes, row_bytes, (size_t)(b - a) * e->H);
            if (dpitch == row_bytes)
                CK(cudaMemcpyAsync(e->d_frames + (size_t)a * fbytes, e->h_stage[slot], (size_t)(b - a) * fbytes, cudaMemcpyHostToDevice,
                                   e->s_copy));
            else
                CK(cudaMemcpy2DAsync(e->d_frames + (size_t)a * fbytes, dpitch, e->h_stage[slot], row_bytes, row_bytes,
                                     (size_t)(b - a) * e->H, cudaMemcpyHostToDevice, e->s_copy));
            CK(cudaEventRecord(e->ev_stage[slot], e->s_copy));
        }
        return B2PIV_OK;
    };
    CK(cudaEventRecord(e->ev_k0, e->s_comp));
    int copied = 0;  // frames already enqueued for copy
    for (int c = 0; c < chunks; ++c) {
        const int p0 = c * per;
        const int p1 = (p0 + per < n_pairs) ? p0 + per : n_pairs;
        if (p0 >= p1) break;
        const int need = p1 + 1;  // frames [0, p1] must be resident
        if (need > copied) {
            rc = h2d(copied, need);
            if (rc) return rc;
            copied = need;
        }
        CK(cudaEventRecord(e->ev_chunk[c], e->s_copy));
        CK(cudaStreamWaitEvent(e->s_comp, e->ev_chunk[c], 0));
        rc = work(p0, p1 - p0);
        if (rc) return rc;
    }
    CK(cudaEventRecord(e->ev_k1, e->s_comp));
    return B2PIV_OK;
}

extern "C" {

int b2piv_pairs_host(b2piv_engine* e, const void* frames, int n_frames, float signal_threshold, float* u, float* v,
                     float* corr_max, float* s2n) {
    if (!e) return B2PIV_ERR_ARG;
    if (!e->planned) return fail(e, B2PIV_ERR_STATE, "b2piv_plan has not been called");
    if (!frames || !u || !v || !corr_max || !s2n) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (n_frames < 2) return fail(e, B2PIV_ERR_ARG, "need at least 2 frames (one pair)");
    CK(cudaSetDevice(e->device));
    const int pitch = host_pitch(e);
    const long long fstride = (long long)e->H * pitch;
    const size_t nw = (size_t)e->n_rows * e->n_cols, n_pairs = (size_t)n_frames - 1;
    const size_t field = nw * n_pairs;
    int rc = ensure(e, &e->d_out, &e->cap_out, field * 4 * sizeof(float));
    if (rc) return rc;
    const bool use_keep = signal_threshold >= 0.f;
    bool keep_done = false;
    rc = pipeline_host(e, frames, n_frames, !use_keep, [&](int p0, int np) -> int {
        if (use_keep && !keep_done) {
            int r2 = make_keep(e, e->d_frames, fstride, pitch, n_frames, signal_threshold, e->s_comp);
            if (r2) return r2;
            keep_done = true;
        }
        Params p = base_params(e, e->d_frames + (size_t)p0 * fstride, fstride, pitch, np);
        p.keep = use_keep ? e->d_keep : nullptr;
        p.u = e->d_out + 0 * field + (size_t)p0 * nw; p.v = e->d_out + 1 * field + (size_t)p0 * nw;
        p.cmax = e->d_out + 2 * field + (size_t)p0 * nw; p.s2n = e->d_out + 3 * field + (size_t)p0 * nw;
        return dispatch_pairs(e, p, e->s_comp);
    });
    if (rc) return rc;
    CK(cudaMemcpyAsync(u, e->d_out + 0 * field, field * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaMemcpyAsync(v, e->d_out + 1 * field, field * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaMemcpyAsync(corr_max, e->d_out + 2 * field, field * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaMemcpyAsync(s2n, e->d_out + 3 * field, field * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaStreamSynchronize(e->s_comp));
    CK(cudaStreamSynchronize(e->s_copy));
    CK(cudaEventElapsedTime(&e->last_kernel_ms, e->ev_k0, e->ev_k1));
    return B2PIV_OK;
}

int b2piv_corr_planes_host(b2piv_engine* e, const void* frames, int n_frames, float signal_threshold, float* corr) {
    if (!e) return B2PIV_ERR_ARG;
    if (!e->planned) return fail(e, B2PIV_ERR_STATE, "b2piv_plan has not been called");
    if (!frames || !corr) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (n_frames < 2) return fail(e, B2PIV_ERR_ARG, "need at least 2 frames (one pair)");
    CK(cudaSetDevice(e->device));
    const int pitch = host_pitch(e);
    const long long fstride = (long long)e->H * pitch;
    const size_t nw = (size_t)e->n_rows * e->n_cols, n_pairs = (size_t)n_frames - 1;
    const size_t field = nw * n_pairs, pl = field * e->wy * e->wx;
    int rc = ensure(e, &e->d_out, &e->cap_out, field * 4 * sizeof(float));
    if (rc) return rc;
    rc = ensure(e, &e->d_planes, &e->cap_planes, pl * sizeof(float));
    if (rc) return rc;
    const bool use_keep = signal_threshold >= 0.f;
    rc = pipeline_host(e, frames, n_frames, false, [&](int p0, int np) -> int {
        if (use_keep) {
            int r2 = make_keep(e, e->d_frames, fstride, pitch, n_frames, signal_threshold, e->s_comp);
            if (r2) return r2;
        }
        Params p = base_params(e, e->d_frames, fstride, pitch, np);
        p.u = e->d_out; p.v = e->d_out + field; p.cmax = e->d_out + 2 * field; p.s2n = e->d_out + 3 * field;
        p.planes = e->d_planes;
        return dispatch_pairs(e, p, e->s_comp);
    });
    if (rc) return rc;
    CK(cudaMemcpyAsync(corr, e->d_planes, pl * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaStreamSynchronize(e->s_comp));
    CK(cudaStreamSynchronize(e->s_copy));
    if (use_keep) {  // NaN planes for windows below the signal threshold, like ffpiv
        std::vector<unsigned char> keep(nw);
        CK(cudaMemcpy(keep.data(), e->d_keep, nw, cudaMemcpyDeviceToHost));
        const size_t npx = (size_t)e->wy * e->wx;
        for (size_t pr = 0; pr < n_pairs; ++pr)
            for (size_t w = 0; w < nw; ++w)
                if (!keep[w])
                    for (size_t i = 0; i < npx; ++i) corr[(pr * nw + w) * npx + i] = nanf("");
    }
    return B2PIV_OK;
}

int b2piv_ens_begin(b2piv_engine* e) {
    if (!e) return B2PIV_ERR_ARG;
    if (!e->planned) return fail(e, B2PIV_ERR_STATE, "b2piv_plan has not been called");
    CK(cudaSetDevice(e->device));
    const size_t nw = (size_t)e->n_rows * e->n_cols, npx = (size_t)e->wy * e->wx;
    const size_t need = nw * npx * sizeof(float);
    if (e->cap_ens < need || e->cap_ens_windows < nw) {
        if (e->d_ens_sum) CK(cudaFree(e->d_ens_sum));
        if (e->d_ens_cnt) CK(cudaFree(e->d_ens_cnt));
        e->d_ens_sum = e->d_ens_cnt = nullptr; e->cap_ens = 0; e->cap_ens_windows = 0;
        CK(cudaMalloc((void**)&e->d_ens_sum, need));
        CK(cudaMalloc((void**)&e->d_ens_cnt, nw * sizeof(float)));
        e->cap_ens = need;
        e->cap_ens_windows = nw;
    }
    CK(cudaMemsetAsync(e->d_ens_sum, 0, need, e->s_comp));
    CK(cudaMemsetAsync(e->d_ens_cnt, 0, nw * sizeof(float), e->s_comp));
    CK(cudaStreamSynchronize(e->s_comp));
    e->ens_open = true;
    return B2PIV_OK;
}

int b2piv_ens_add_device(b2piv_engine* e, const void* d_frames, long long frame_stride_bytes, int pitch_bytes, int n_frames,
                         float corr_min, float s2n_min, float signal_threshold, float* d_corr_max, float* d_s2n,
                         void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    if (!e->planned || !e->ens_open) return fail(e, B2PIV_ERR_STATE, "call b2piv_plan and b2piv_ens_begin first");
    if (!d_frames || !d_corr_max || !d_s2n) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (n_frames < 2) return fail(e, B2PIV_ERR_ARG, "need at least 2 frames (one pair)");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Params p = base_params(e, d_frames, frame_stride_bytes, pitch_bytes, n_frames - 1);
    if (signal_threshold >= 0.f) {
        int rc = make_keep(e, d_frames, frame_stride_bytes, pitch_bytes, n_frames, signal_threshold, st);
        if (rc) return rc;
        p.keep = e->d_keep;
    }
    p.cmax = d_corr_max; p.s2n = d_s2n;
    EnsParams ep{corr_min, s2n_min, e->d_ens_sum, e->d_ens_cnt};
    return dispatch_ens(e, p, ep, st);
}

int b2piv_ens_add_host(b2piv_engine* e, const void* frames, int n_frames, float corr_min, float s2n_min,
                       float signal_threshold, float* corr_max, float* s2n) {
    if (!e) return B2PIV_ERR_ARG;
    if (!e->planned || !e->ens_open) return fail(e, B2PIV_ERR_STATE, "call b2piv_plan and b2piv_ens_begin first");
    if (!frames || !corr_max || !s2n) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (n_frames < 2) return fail(e, B2PIV_ERR_ARG, "need at least 2 frames (one pair)");
    CK(cudaSetDevice(e->device));
    const int pitch = host_pitch(e);
    const long long fstride = (long long)e->H * pitch;
    const size_t nw = (size_t)e->n_rows * e->n_cols, field = nw * ((size_t)n_frames - 1);
    int rc = ensure(e, &e->d_out, &e->cap_out, field * 4 * sizeof(float));
    if (rc) return rc;
    rc = pipeline_host(e, frames, n_frames, false, [&](int, int) -> int {
        return b2piv_ens_add_device(e, e->d_frames, fstride, pitch, n_frames, corr_min, s2n_min, signal_threshold,
                                    e->d_out, e->d_out + field, e->s_comp);
    });
    if (rc) return rc;
    CK(cudaMemcpyAsync(corr_max, e->d_out, field * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaMemcpyAsync(s2n, e->d_out + field, field * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaStreamSynchronize(e->s_comp));
    CK(cudaStreamSynchronize(e->s_copy));
    CK(cudaEventElapsedTime(&e->last_kernel_ms, e->ev_k0, e->ev_k1));
    return B2PIV_OK;
}

int b2piv_ens_accum(b2piv_engine* e, float** d_plane_sum, float** d_count, long long* n_plane_floats, long long* n_windows) {
    if (!e) return B2PIV_ERR_ARG;
    if (!e->planned || !e->ens_open) return fail(e, B2PIV_ERR_STATE, "call b2piv_plan and b2piv_ens_begin first");
    const long long nw = (long long)e->n_rows * e->n_cols;
    if (d_plane_sum) *d_plane_sum = e->d_ens_sum;
    if (d_count) *d_count = e->d_ens_cnt;
    if (n_plane_floats) *n_plane_floats = nw * e->wy * e->wx;
    if (n_windows) *n_windows = nw;
    return B2PIV_OK;
}

int b2piv_ens_finish_host(b2piv_engine* e, float min_count, float* u, float* v, float* count) {
    if (!e) return B2PIV_ERR_ARG;
    if (!e->planned || !e->ens_open) return fail(e, B2PIV_ERR_STATE, "call b2piv_plan and b2piv_ens_begin first");
    if (!u || !v) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    CK(cudaSetDevice(e->device));
    const size_t nw = (size_t)e->n_rows * e->n_cols;
    int rc = ensure(e, &e->d_out, &e->cap_out, nw * 4 * sizeof(float));
    if (rc) return rc;
    ens_finish_kernel<<<(unsigned)nw, 256, 0, e->s_comp>>>(e->d_ens_sum, e->d_ens_cnt, e->wy, e->wx, min_count, e->border_nan,
                                                           e->gauss_eps, e->d_out, e->d_out + nw);
    CK(cudaGetLastError());
    e->launches++;
    CK(cudaMemcpyAsync(u, e->d_out, nw * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaMemcpyAsync(v, e->d_out + nw, nw * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    if (count) CK(cudaMemcpyAsync(count, e->d_ens_cnt, nw * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaStreamSynchronize(e->s_comp));
    return B2PIV_OK;
}

int b2piv_peaks_host(b2piv_engine* e, const float* corr, long long n_planes, int wy, int wx, float* u, float* v) {
    if (!e) return B2PIV_ERR_ARG;
    if (!corr || !u || !v) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (n_planes < 0 || wy < 3 || wx < 3) return fail(e, B2PIV_ERR_ARG, "need n_planes >= 0 and planes of at least 3x3");
    if (n_planes == 0) return B2PIV_OK;
    CK(cudaSetDevice(e->device));
    const size_t npx = (size_t)wy * wx;
    int rc = ensure(e, &e->d_planes, &e->cap_planes, (size_t)n_planes * npx * sizeof(float));
    if (rc) return rc;
    rc = ensure(e, &e->d_out, &e->cap_out, (size_t)n_planes * 3 * sizeof(float));
    if (rc) return rc;
    float* d_one = e->d_out + 2 * n_planes;   // per-plane divisor 1 (the kernel divides the plane by its count)
    std::vector<float> ones((size_t)n_planes, 1.0f);
    CK(cudaMemcpyAsync(e->d_planes, corr, (size_t)n_planes * npx * sizeof(float), cudaMemcpyHostToDevice, e->s_comp));
    CK(cudaMemcpyAsync(d_one, ones.data(), (size_t)n_planes * sizeof(float), cudaMemcpyHostToDevice, e->s_comp));
    ens_finish_kernel<<<(unsigned)n_planes, 256, 0, e->s_comp>>>(e->d_planes, d_one, wy, wx, 0.f, e->border_nan, e->gauss_eps, e->d_out,
                                                                  e->d_out + n_planes);
    CK(cudaGetLastError());
    e->launches++;
    CK(cudaMemcpyAsync(u, e->d_out, (size_t)n_planes * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaMemcpyAsync(v, e->d_out + n_planes, (size_t)n_planes * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaStreamSynchronize(e->s_comp));
    return B2PIV_OK;
}

// ---- frame pre-processing on the device (kernels in preproc.cuh) -------------------------------------
static int pre_grid(const b2piv_engine* e, long long n) {
    long long g = (n + 4095) / 4096, cap = (long long)e->sm_count * 8;   // 16 elements per thread and iteration
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

int b2piv_pre_normalize_device(b2piv_engine* e, const void* d_frames, int dtype, int n_frames, int height, int width, int time_interval,
                               unsigned char* d_out, void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    if (!d_frames || !d_out) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (dtype != B2PIV_U8 && dtype != B2PIV_F32) return fail(e, B2PIV_ERR_ARG, "dtype must be B2PIV_U8 or B2PIV_F32");
    if (n_frames < 1 || height < 1 || width < 1) return fail(e, B2PIV_ERR_ARG, "bad shape");
    if (time_interval < 1) return fail(e, B2PIV_ERR_ARG, "time_interval must be >= 1 (too few frames for the requested samples)");
    const int step_py = time_interval;
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const long long fe = (long long)height * width;
    int rc = ensure(e, &e->d_pre_mean, &e->cap_pre_mean, (size_t)fe * sizeof(float));
    if (rc) return rc;
    rc = ensure(e, &e->d_pre_mm, &e->cap_pre_mm, (size_t)n_frames * 2 * sizeof(unsigned));
    if (rc) return rc;
    CK(cudaMemsetAsync(e->d_pre_mm, 0xff, (size_t)n_frames * sizeof(unsigned), st));
    CK(cudaMemsetAsync(e->d_pre_mm + n_frames, 0x00, (size_t)n_frames * sizeof(unsigned), st));
    long long gx = (fe / 4 + 255) / 256;   // 4 pixels per thread and grid-stride iteration
    gx = gx < 1 ? 1 : (gx > (long long)e->sm_count * 8 ? (long long)e->sm_count * 8 : gx);
    const dim3 g1((unsigned)gx), g2((unsigned)gx, (n_frames + PRE_FPB - 1) / PRE_FPB);
    if (dtype == B2PIV_U8) {
        pre_mean_kernel<unsigned char><<<g1, 256, 0, st>>>((const unsigned char*)d_frames, fe, n_frames, step_py, e->d_pre_mean);
        pre_minmax_kernel<unsigned char><<<g2, 256, 0, st>>>((const unsigned char*)d_frames, e->d_pre_mean, fe, n_frames, e->d_pre_mm);
        pre_normalize_kernel<unsigned char><<<g2, 256, 0, st>>>((const unsigned char*)d_frames, e->d_pre_mean, e->d_pre_mm, fe, n_frames, d_out);
    } else {
        pre_mean_kernel<float><<<g1, 256, 0, st>>>((const float*)d_frames, fe, n_frames, step_py, e->d_pre_mean);
        pre_minmax_kernel<float><<<g2, 256, 0, st>>>((const float*)d_frames, e->d_pre_mean, fe, n_frames, e->d_pre_mm);
        pre_normalize_kernel<float><<<g2, 256, 0, st>>>((const float*)d_frames, e->d_pre_mean, e->d_pre_mm, fe, n_frames, d_out);
    }
    CK(cudaGetLastError());
    e->launches += 3;
    return B2PIV_OK;
}

int b2piv_pre_time_diff_device(b2piv_engine* e, const void* d_frames, int dtype, int n_frames, int height, int width, float thres,
                               int absolute, float* d_out, void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    if (!d_frames || !d_out) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (dtype != B2PIV_U8 && dtype != B2PIV_F32) return fail(e, B2PIV_ERR_ARG, "dtype must be B2PIV_U8 or B2PIV_F32");
    if (n_frames < 2) return fail(e, B2PIV_ERR_ARG, "need at least 2 frames");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const long long fe = (long long)height * width, n_out = fe * (n_frames - 1);
    if (dtype == B2PIV_U8)
        pre_time_diff_kernel<unsigned char><<<pre_grid(e, n_out), 256, 0, st>>>((const unsigned char*)d_frames, fe, n_out, thres, absolute, d_out);
    else
        pre_time_diff_kernel<float><<<pre_grid(e, n_out), 256, 0, st>>>((const float*)d_frames, fe, n_out, thres, absolute, d_out);
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}

int b2piv_pre_minmax_device(b2piv_engine* e, const void* d_in, int dtype, long long count, float lo, float hi, void* d_out,
                            void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    if (!d_in || !d_out) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (dtype != B2PIV_U8 && dtype != B2PIV_F32) return fail(e, B2PIV_ERR_ARG, "dtype must be B2PIV_U8 or B2PIV_F32");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (dtype == B2PIV_U8) {
        const float l = lo < 0.f ? 0.f : lo, h = hi > 255.f ? 255.f : hi;
        pre_clamp_kernel<unsigned char><<<pre_grid(e, count), 256, 0, st>>>((const unsigned char*)d_in, count, l, h, (unsigned char*)d_out);
    } else {
        pre_clamp_kernel<float><<<pre_grid(e, count), 256, 0, st>>>((const float*)d_in, count, lo, hi, (float*)d_out);
    }
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}

// OpenCV's getGaussianKernel(ksize, sigma <= 0, CV_32F): fixed dyadic tables up to 9 taps, exp(-x^2 / 2 sigma^2) normalised
// with sigma = 0.3 * ((ksize - 1) / 2 - 1) + 0.8 beyond (checked against cv2 in tests/test_preprocess.py)
static bool gauss_taps(int ksize, float* k) {
    if (ksize < 1 || ksize > 2 * GB_MAXR + 1 || (ksize & 1) == 0) return false;
    static const float t1[] = {1.f}, t3[] = {0.25f, 0.5f, 0.25f}, t5[] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f},
                       t7[] = {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f},
                       t9[] = {4 / 256.f, 13 / 256.f, 30 / 256.f, 51 / 256.f, 60 / 256.f, 51 / 256.f, 30 / 256.f, 13 / 256.f, 4 / 256.f};
    const float* tab = ksize == 1 ? t1 : ksize == 3 ? t3 : ksize == 5 ? t5 : ksize == 7 ? t7 : ksize == 9 ? t9 : nullptr;
    if (tab) { for (int i = 0; i < ksize; ++i) k[i] = tab[i]; return true; }
    const double sigma = 0.3 * ((ksize - 1) * 0.5 - 1.0) + 0.8;
    double sum = 0.0;
    std::vector<double> g(ksize);
    for (int i = 0; i < ksize; ++i) { const double x = i - (ksize - 1) * 0.5; g[i] = exp(-x * x / (2.0 * sigma * sigma)); sum += g[i]; }
    for (int i = 0; i < ksize; ++i) k[i] = (float)(g[i] / sum);
    return true;
}

int b2piv_pre_gauss_device(b2piv_engine* e, const void* d_frames, int dtype, int n_frames, int height, int width, int ksize1, int ksize2,
                           float* d_out, void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    if (!d_frames || !d_out) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (dtype != B2PIV_U8 && dtype != B2PIV_F32) return fail(e, B2PIV_ERR_ARG, "dtype must be B2PIV_U8 or B2PIV_F32");
    GaussTaps taps;
    memset(&taps, 0, sizeof(taps));
    if (!gauss_taps(ksize2, taps.k2)) return fail(e, B2PIV_ERR_ARG, "kernel size must be odd and between 1 and 31");
    taps.r2 = ksize2 / 2;
    taps.r1 = -1;
    if (ksize1 > 0) {
        if (!gauss_taps(ksize1, taps.k1)) return fail(e, B2PIV_ERR_ARG, "kernel size must be odd and between 1 and 31");
        taps.r1 = ksize1 / 2;
    }
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const dim3 sgrid((width + GS_BW - 1) / GS_BW, (height + GS_SH - 1) / GS_SH, n_frames);
    const int rmax = taps.r2 > taps.r1 ? taps.r2 : taps.r1;
    bool fast = height > rmax && width > rmax && sgrid.y <= 65535 && sgrid.z <= 65535;   // one reflection suffices
    if (fast) {
#define B2_GAUSS_CASE(A, B)                                                                                                      \
    if (taps.r1 == (A) && taps.r2 == (B)) {                                                                                      \
        if (dtype == B2PIV_U8) pre_gauss_strip_kernel<unsigned char, A, B><<<sgrid, GS_BW, 0, st>>>((const unsigned char*)d_frames, height, width, taps, d_out); \
        else pre_gauss_strip_kernel<float, A, B><<<sgrid, GS_BW, 0, st>>>((const float*)d_frames, height, width, taps, d_out);   \
    } else
        B2_GAUSS_CASE(-1, 1) B2_GAUSS_CASE(-1, 2) B2_GAUSS_CASE(-1, 3) B2_GAUSS_CASE(1, 2) B2_GAUSS_CASE(1, 3) B2_GAUSS_CASE(2, 4)
        fast = false;
#undef B2_GAUSS_CASE
    }
    if (!fast) {
        const int R = taps.r2 > taps.r1 ? taps.r2 : taps.r1;
        const size_t smem = ((size_t)(GB_TY + 2 * R) * (GB_TX + 2 * R) + 2 * (size_t)(GB_TY + 2 * R) * GB_TX) * sizeof(float);
        const dim3 grid((width + GB_TX - 1) / GB_TX, (height + GB_TY - 1) / GB_TY, n_frames), block(GB_TX, GB_TY);
        if (dtype == B2PIV_U8) {
            CK(cudaFuncSetAttribute(pre_gauss_kernel<unsigned char>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            pre_gauss_kernel<unsigned char><<<grid, block, smem, st>>>((const unsigned char*)d_frames, height, width, taps, d_out);
        } else {
            CK(cudaFuncSetAttribute(pre_gauss_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            pre_gauss_kernel<float><<<grid, block, smem, st>>>((const float*)d_frames, height, width, taps, d_out);
        }
    }
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}

// ---- orthoprojection with index maps (kernel in project.cuh) -----------------------------------------
// Merges the reference's two maps (nearest: out[idx_ortho[i]] = img[idx_img[i]], project.py:147-149; mean: group g =
// samples i with norm_idx[i] == g, written to out[uidx[g]], project.py:150-154) into one CSR list per target pixel.
// Later assignments win exactly as in the reference's sequential fancy-index stores.
int b2piv_project_plan(b2piv_engine* e, int height, int width, int out_height, int out_width, const long long* idx_img,
                       const long long* idx_ortho, long long n_nearest, const long long* src_idx, const long long* norm_idx,
                       long long n_samples, const long long* uidx, long long n_groups) {
    if (!e) return B2PIV_ERR_ARG;
    if (height < 1 || width < 1 || out_height < 1 || out_width < 1) return fail(e, B2PIV_ERR_ARG, "bad shape");
    if (n_nearest < 0 || n_samples < 0 || n_groups < 0) return fail(e, B2PIV_ERR_ARG, "negative count");
    if ((n_nearest && (!idx_img || !idx_ortho)) || (n_samples && (!src_idx || !norm_idx)) || (n_groups && !uidx))
        return fail(e, B2PIV_ERR_ARG, "NULL index map");
    const long long n_in = (long long)height * width, n_out = (long long)out_height * out_width;
    if (n_in >= (1ll << 31) || n_out >= (1ll << 31) || n_nearest + n_samples >= (1ll << 31))
        return fail(e, B2PIV_ERR_UNSUPPORTED, "index maps beyond 2^31 entries");
    std::vector<int> nn((size_t)n_out, -1), grp((size_t)n_out, -1);
    for (long long i = 0; i < n_nearest; ++i) {
        if (idx_img[i] < 0 || idx_img[i] >= n_in) return fail(e, B2PIV_ERR_ARG, "idx_img out of range");
        if (idx_ortho[i] < 0 || idx_ortho[i] >= n_out) return fail(e, B2PIV_ERR_ARG, "idx_ortho out of range");
        nn[(size_t)idx_ortho[i]] = (int)idx_img[i];
    }
    for (long long g = 0; g < n_groups; ++g) {
        if (uidx[g] < 0 || uidx[g] >= n_out) return fail(e, B2PIV_ERR_ARG, "uidx out of range");
        grp[(size_t)uidx[g]] = (int)g;
    }
    std::vector<int> gcount((size_t)n_groups + 1, 0);
    for (long long i = 0; i < n_samples; ++i) {
        if (src_idx[i] < 0 || src_idx[i] >= n_in) return fail(e, B2PIV_ERR_ARG, "src_idx out of range");
        if (norm_idx[i] < 0 || norm_idx[i] >= n_groups) return fail(e, B2PIV_ERR_ARG, "norm_idx out of range");
        gcount[(size_t)norm_idx[i]]++;
    }
    for (long long g = 0; g < n_groups; ++g)
        if (gcount[(size_t)g] == 0) return fail(e, B2PIV_ERR_ARG, "empty group in norm_idx (the reference would divide 0 by 0)");
    // stable counting sort of the samples by group keeps the reference's accumulation order (ascending i)
    std::vector<int> gstart((size_t)n_groups + 1, 0);
    for (long long g = 0; g < n_groups; ++g) gstart[(size_t)g + 1] = gstart[(size_t)g] + gcount[(size_t)g];
    std::vector<int> gsrc((size_t)n_samples), gpos(gstart.begin(), gstart.end());
    for (long long i = 0; i < n_samples; ++i) gsrc[(size_t)gpos[(size_t)norm_idx[i]]++] = (int)src_idx[i];
    std::vector<int> off((size_t)n_out + 1), src;
    src.reserve((size_t)(n_nearest + n_samples));
    for (long long j = 0; j < n_out; ++j) {
        off[(size_t)j] = (int)src.size();
        const int g = grp[(size_t)j];
        if (g >= 0) src.insert(src.end(), gsrc.begin() + gstart[(size_t)g], gsrc.begin() + gstart[(size_t)g + 1]);
        else if (nn[(size_t)j] >= 0) src.push_back(nn[(size_t)j]);
    }
    off[(size_t)n_out] = (int)src.size();
    CK(cudaSetDevice(e->device));
    int rc = ensure(e, &e->d_proj_off, &e->cap_proj_off, off.size() * sizeof(int));
    if (rc) return rc;
    rc = ensure(e, &e->d_proj_src, &e->cap_proj_src, (src.size() + 1) * sizeof(int));
    if (rc) return rc;
    CK(cudaMemcpy(e->d_proj_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!src.empty()) CK(cudaMemcpy(e->d_proj_src, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice));
    e->proj_h = height; e->proj_w = width; e->proj_out_h = out_height; e->proj_out_w = out_width;
    e->proj_samples = (long long)src.size();
    return B2PIV_OK;
}

}  // extern "C"
template <typename TI, typename TO>
static void launch_project(const b2piv_engine* e, const void* d_frames, int n_frames, void* d_out, cudaStream_t st) {
    constexpr int FR = 4;
    const int n_out = e->proj_out_h * e->proj_out_w;
    const dim3 grid((n_out + 255) / 256, (n_frames + FR - 1) / FR);
    proj_gather_kernel<TI, TO, FR><<<grid, 256, 0, st>>>((const TI*)d_frames, (long long)e->proj_h * e->proj_w, n_frames, e->d_proj_off,
                                                         e->d_proj_src, n_out, (TO*)d_out);
}
extern "C" {

int b2piv_project_device(b2piv_engine* e, const void* d_frames, int dtype, int n_frames, void* d_out, int out_dtype, void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    if (!e->d_proj_off) return fail(e, B2PIV_ERR_STATE, "b2piv_project_plan has not been called");
    if (!d_frames || !d_out) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (dtype != B2PIV_U8 && dtype != B2PIV_F32) return fail(e, B2PIV_ERR_ARG, "dtype must be B2PIV_U8 or B2PIV_F32");
    if (out_dtype != B2PIV_F32 && out_dtype != dtype) return fail(e, B2PIV_ERR_ARG, "out_dtype must be the input dtype or B2PIV_F32");
    if (n_frames < 1) return fail(e, B2PIV_ERR_ARG, "need at least 1 frame");
    if ((n_frames + 3) / 4 > 65535) return fail(e, B2PIV_ERR_UNSUPPORTED, "more than 262140 frames per call");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (dtype == B2PIV_U8 && out_dtype == B2PIV_U8) launch_project<unsigned char, unsigned char>(e, d_frames, n_frames, d_out, st);
    else if (dtype == B2PIV_U8) launch_project<unsigned char, float>(e, d_frames, n_frames, d_out, st);
    else launch_project<float, float>(e, d_frames, n_frames, d_out, st);
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}

// ---- velocimetry mask stack and result packing on the device (kernels in mask.cuh) -------------
// 2-D launch over (locations, time): x covers the locations (at most sm_count * 8 blocks), y strides over time so that the
// whole grid holds about sm_count * 16 blocks
static dim3 mask_grid2(const b2piv_engine* e, long long n_xy, int n_time, int block = 256) {
    long long gx = (n_xy + block - 1) / block, cap = (long long)e->sm_count * 8;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    long long gy = ((long long)e->sm_count * 16 + gx - 1) / gx;
    if (gy > n_time) gy = n_time;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    return dim3((unsigned)gx, (unsigned)gy, 1);
}
static dim3 window_grid(const b2piv_engine* e, int n_time, int ny, int nx) {
    const unsigned gx = (unsigned)((nx + 31) / 32), gy = (unsigned)((ny + 7) / 8);
    long long gz = ((long long)e->sm_count * 16 + (long long)gx * gy - 1) / ((long long)gx * gy);
    if (gz > n_time) gz = n_time;
    if (gz > 65535) gz = 65535;
    if (gz < 1) gz = 1;
    return dim3(gx, gy, (unsigned)gz);
}
static int mask_grid(const b2piv_engine* e, long long n, int block = 256) {
    long long g = (n + block - 1) / block, cap = (long long)e->sm_count * 8;
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}
#define MASK_PROLOGUE(cond_null)                                                        \
    if (!e) return B2PIV_ERR_ARG;                                                       \
    if (cond_null) return fail(e, B2PIV_ERR_ARG, "NULL pointer");                       \
    CK(cudaSetDevice(e->device));                                                       \
    cudaStream_t st = (cudaStream_t)cuda_stream;
#define MASK_EPILOGUE(k)                                                                \
    CK(cudaGetLastError());                                                             \
    e->launches += (k);                                                                 \
    return B2PIV_OK;

// time statistics: segmented kernel (loads parallel in time, sums sequential) up to 1024 time steps, else one thread per location
static void launch_time_stats(const b2piv_engine* e, const float* d_field, int n_time, long long n_xy, int* d_count, float* d_mean,
                              float* d_std, cudaStream_t st) {
    const long long n_blocks = (n_xy + 31) / 32;
    if (n_time <= 1024 && n_time >= 16) {
        constexpr int L = 32;
        const int S = (n_time + L - 1) / L;
        const long long cap = (long long)e->sm_count * (2048 / (32 * S));
        const unsigned grid = (unsigned)(n_blocks < cap ? n_blocks : cap);
        time_stats_seg_kernel<L><<<grid, dim3(32, S), 0, st>>>(d_field, n_time, n_xy, d_count, d_mean, d_std);
    } else {
        time_stats_kernel<<<mask_grid(e, n_xy, 64), 64, 0, st>>>(d_field, n_time, n_xy, d_count, d_mean, d_std);
    }
}

int b2piv_mask_elementwise(b2piv_engine* e, int op, const float* d_a, const float* d_b, long long count, float p0, float p1,
                           unsigned char* d_mask, void* cuda_stream) {
    MASK_PROLOGUE(!d_a || !d_mask || (op != B2PIV_MASK_THRESHOLD && !d_b))
    if (count < 0) return fail(e, B2PIV_ERR_ARG, "negative count");
    if (count == 0) return B2PIV_OK;
    const int g = mask_grid(e, count);
    if (op == B2PIV_MASK_MINMAX) mask_elem_kernel<0><<<g, 256, 0, st>>>(d_a, d_b, count, p0, p1, d_mask);
    else if (op == B2PIV_MASK_ANGLE) mask_elem_kernel<1><<<g, 256, 0, st>>>(d_a, d_b, count, p0, p1, d_mask);
    else if (op == B2PIV_MASK_THRESHOLD) mask_elem_kernel<2><<<g, 256, 0, st>>>(d_a, d_a, count, p0, p1, d_mask);
    else return fail(e, B2PIV_ERR_ARG, "unknown element-wise mask op");
    MASK_EPILOGUE(1)
}

int b2piv_time_stats(b2piv_engine* e, const float* d_field, int n_time, long long n_xy, int* d_count, float* d_mean, float* d_std,
                     void* cuda_stream) {
    MASK_PROLOGUE(!d_field)
    if (n_time < 1 || n_xy < 1) return fail(e, B2PIV_ERR_ARG, "empty field");
    launch_time_stats(e, d_field, n_time, n_xy, d_count, d_mean, d_std, st);
    MASK_EPILOGUE(1)
}

int b2piv_mask_count(b2piv_engine* e, const float* d_vx, int n_time, long long n_xy, double tolerance, unsigned char* d_mask_xy,
                     void* cuda_stream) {
    MASK_PROLOGUE(!d_vx || !d_mask_xy)
    if (n_time < 1 || n_xy < 1) return fail(e, B2PIV_ERR_ARG, "empty field");
    int rc = ensure(e, &e->d_mask_ws, &e->cap_mask_ws, (size_t)n_xy * sizeof(int));
    if (rc) return rc;
    int* cnt = reinterpret_cast<int*>(e->d_mask_ws);
    launch_time_stats(e, d_vx, n_time, n_xy, cnt, nullptr, nullptr, st);
    // count > tolerance * T  <=>  count >= floor(tolerance * T) + 1 (count is an integer; the product is the reference's float64 one)
    const double thr = tolerance * (double)n_time;
    const int min_count = thr < -1.0 ? 0 : (thr > 2.0e9 ? 2147483647 : (int)std::floor(thr) + 1);
    mask_count_kernel<<<mask_grid(e, n_xy), 256, 0, st>>>(cnt, n_xy, min_count, d_mask_xy);
    MASK_EPILOGUE(2)
}

// mean / std over time of both components into the engine's workspace: [xm | xs | ym | ys], n_xy floats each
static int mask_stats_xy(b2piv_engine* e, const float* d_vx, const float* d_vy, int n_time, long long n_xy, cudaStream_t st) {
    int rc = ensure(e, &e->d_mask_ws, &e->cap_mask_ws, (size_t)4 * n_xy * sizeof(float));
    if (rc) return rc;
    float* w = e->d_mask_ws;
    launch_time_stats(e, d_vx, n_time, n_xy, nullptr, w, w + n_xy, st);
    launch_time_stats(e, d_vy, n_time, n_xy, nullptr, w + 2 * n_xy, w + 3 * n_xy, st);
    return B2PIV_OK;
}

int b2piv_mask_outliers(b2piv_engine* e, const float* d_vx, const float* d_vy, int n_time, long long n_xy, float tolerance, int mode_and,
                        unsigned char* d_mask, void* cuda_stream) {
    MASK_PROLOGUE(!d_vx || !d_vy || !d_mask)
    if (n_time < 1 || n_xy < 1) return fail(e, B2PIV_ERR_ARG, "empty field");
    int rc = mask_stats_xy(e, d_vx, d_vy, n_time, n_xy, st);
    if (rc) return rc;
    const float* w = e->d_mask_ws;
    mask_outliers_kernel<<<mask_grid2(e, n_xy, n_time), 256, 0, st>>>(d_vx, d_vy, n_time, n_xy, w, w + n_xy, w + 2 * n_xy, w + 3 * n_xy,
                                                                        tolerance, mode_and, d_mask);
    MASK_EPILOGUE(3)
}

int b2piv_mask_variance(b2piv_engine* e, const float* d_vx, const float* d_vy, int n_time, long long n_xy, float tolerance, int mode_and,
                        unsigned char* d_mask_xy, void* cuda_stream) {
    MASK_PROLOGUE(!d_vx || !d_vy || !d_mask_xy)
    if (n_time < 1 || n_xy < 1) return fail(e, B2PIV_ERR_ARG, "empty field");
    int rc = mask_stats_xy(e, d_vx, d_vy, n_time, n_xy, st);
    if (rc) return rc;
    const float* w = e->d_mask_ws;
    mask_variance_kernel<<<mask_grid(e, n_xy), 256, 0, st>>>(n_xy, w, w + n_xy, w + 2 * n_xy, w + 3 * n_xy, tolerance, mode_and, d_mask_xy);
    MASK_EPILOGUE(3)
}

int b2piv_mask_rolling(b2piv_engine* e, const float* d_vx, const float* d_vy, int n_time, long long n_xy, int wdw, float tolerance,
                       unsigned char* d_mask, void* cuda_stream) {
    MASK_PROLOGUE(!d_vx || !d_vy || !d_mask)
    if (n_time < 1 || n_xy < 1) return fail(e, B2PIV_ERR_ARG, "empty field");
    if (wdw < 1) return fail(e, B2PIV_ERR_ARG, "rolling window must be >= 1");
    // time is cut into chunks of >= 8 * wdw steps (a chunk re-reads wdw - 1 steps of halo), one chunk per grid.y index
    dim3 g = mask_grid2(e, n_xy, n_time, 128);
    int t_chunk = (n_time + (int)g.y - 1) / (int)g.y;
    if (t_chunk < 8 * wdw) t_chunk = 8 * wdw;
    g.y = (unsigned)((n_time + t_chunk - 1) / t_chunk);
#define ROLL_CASE(WD) case WD: mask_rolling_kernel<WD><<<g, 128, 0, st>>>(d_vx, d_vy, n_time, n_xy, wdw, tolerance, t_chunk, d_mask); break;
    switch (wdw) {
        ROLL_CASE(1) ROLL_CASE(2) ROLL_CASE(3) ROLL_CASE(4) ROLL_CASE(5) ROLL_CASE(6) ROLL_CASE(7) ROLL_CASE(8) ROLL_CASE(9) ROLL_CASE(10)
        ROLL_CASE(11) ROLL_CASE(12) ROLL_CASE(13) ROLL_CASE(14) ROLL_CASE(15) ROLL_CASE(16)
        default: mask_rolling_kernel<0><<<g, 128, 0, st>>>(d_vx, d_vy, n_time, n_xy, wdw, tolerance, t_chunk, d_mask);
    }
#undef ROLL_CASE
    MASK_EPILOGUE(1)
}

static bool window_args(b2piv_engine* e, int n_time, int ny, int nx, const int* strides, WindowArgs* w) {
    if (n_time < 1 || ny < 1 || nx < 1 || !strides) { e->err = "empty field or NULL strides"; return false; }
    *w = WindowArgs{n_time, ny, nx, strides[0], strides[1], strides[2], strides[3]};
    if (w->wx1 < w->wx0 || w->wy1 <= w->wy0) { e->err = "window has no strides (x: [min, max], y: [min, max) like helpers.stack_window)"; return false; }
    return true;
}

int b2piv_mask_window_nan(b2piv_engine* e, const float* d_vx, int n_time, int ny, int nx, const int* strides, double tolerance,
                          unsigned char* d_mask, void* cuda_stream) {
    MASK_PROLOGUE(!d_vx || !d_mask)
    WindowArgs w;
    if (!window_args(e, n_time, ny, nx, strides, &w)) return B2PIV_ERR_ARG;
    const double n_strides = (double)(w.wx1 - w.wx0 + 1) * (double)(w.wy1 - w.wy0);
    // valid >= tolerance * n_strides  <=>  valid >= ceil(tolerance * n_strides)
    const double thr = tolerance * n_strides;
    const int min_count = thr <= 0.0 ? 0 : (thr > 2.0e9 ? 2147483647 : (int)std::ceil(thr));
    mask_window_nan_kernel<<<window_grid(e, n_time, ny, nx), dim3(32, 8), 0, st>>>(d_vx, w, min_count, d_mask);
    MASK_EPILOGUE(1)
}

int b2piv_mask_window_mean(b2piv_engine* e, const float* d_vx, const float* d_vy, int n_time, int ny, int nx, const int* strides,
                           float tolerance, int mode_and, unsigned char* d_mask, void* cuda_stream) {
    MASK_PROLOGUE(!d_vx || !d_vy || !d_mask)
    WindowArgs w;
    if (!window_args(e, n_time, ny, nx, strides, &w)) return B2PIV_ERR_ARG;
    mask_window_mean_kernel<<<window_grid(e, n_time, ny, nx), dim3(32, 8), 0, st>>>(d_vx, d_vy, w, tolerance, mode_and, d_mask);
    MASK_EPILOGUE(1)
}

int b2piv_window_replace(b2piv_engine* e, float* const* d_fields, int n_fields, int n_time, int ny, int nx, const int* strides,
                         int iterations, void* cuda_stream) {
    MASK_PROLOGUE(!d_fields)
    WindowArgs w;
    if (!window_args(e, n_time, ny, nx, strides, &w)) return B2PIV_ERR_ARG;
    if (n_fields < 1 || n_fields > 4) return fail(e, B2PIV_ERR_ARG, "1..4 fields");
    const long long n = (long long)n_time * ny * nx;
    int rc = ensure(e, &e->d_mask_ws, &e->cap_mask_ws, (size_t)n * sizeof(float));
    if (rc) return rc;
    int launches = 0;
    for (int it = 0; it < iterations; ++it) {
        for (int k = 0; k < n_fields; ++k) {
            if (!d_fields[k]) return fail(e, B2PIV_ERR_ARG, "NULL field");
            window_replace_kernel<<<window_grid(e, n_time, ny, nx), dim3(32, 8), 0, st>>>(d_fields[k], w, e->d_mask_ws);
            CK(cudaMemcpyAsync(d_fields[k], e->d_mask_ws, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st));
            ++launches;
        }
    }
    MASK_EPILOGUE(launches)
}

int b2piv_mask_apply(b2piv_engine* e, float* const* d_fields, int n_fields, int n_time, long long n_xy, const unsigned char* d_mask,
                     int mask_has_time, void* cuda_stream) {
    MASK_PROLOGUE(!d_fields || !d_mask)
    if (n_fields < 1 || n_fields > 4) return fail(e, B2PIV_ERR_ARG, "1..4 fields");
    if (n_time < 1 || n_xy < 1) return fail(e, B2PIV_ERR_ARG, "empty field");
    Fields4 fs;
    fs.n = n_fields;
    for (int k = 0; k < 4; ++k) {
        fs.f[k] = k < n_fields ? d_fields[k] : nullptr;
        if (k < n_fields && !fs.f[k]) return fail(e, B2PIV_ERR_ARG, "NULL field");
    }
    mask_apply_kernel<<<mask_grid2(e, n_xy, n_time), 256, 0, st>>>(fs, n_time, n_xy, d_mask, mask_has_time);
    MASK_EPILOGUE(1)
}

int b2piv_encode_int16(b2piv_engine* e, const float* d_field, long long count, float scale_factor, int fill_value, short* d_out,
                       void* cuda_stream) {
    MASK_PROLOGUE(!d_field || !d_out)
    if (count <= 0) return count == 0 ? B2PIV_OK : fail(e, B2PIV_ERR_ARG, "negative count");
    if (!(scale_factor > 0.f) || fill_value < -32768 || fill_value > 32767) return fail(e, B2PIV_ERR_ARG, "bad scale_factor / _FillValue");
    encode_i16_kernel<<<mask_grid(e, count), 256, 0, st>>>(d_field, count, scale_factor, fill_value, d_out);
    MASK_EPILOGUE(1)
}

int b2piv_decode_int16(b2piv_engine* e, const short* d_packed, long long count, float scale_factor, int fill_value, float* d_out,
                       void* cuda_stream) {
    MASK_PROLOGUE(!d_packed || !d_out)
    if (count <= 0) return count == 0 ? B2PIV_OK : fail(e, B2PIV_ERR_ARG, "negative count");
    decode_i16_kernel<<<mask_grid(e, count), 256, 0, st>>>(d_packed, count, scale_factor, fill_value, d_out);
    MASK_EPILOGUE(1)
}

int b2piv_rotate_uv(b2piv_engine* e, const float* d_u, const float* d_v, long long count, double theta, double* d_u2, double* d_v2,
                    void* cuda_stream) {
    MASK_PROLOGUE(!d_u || !d_v || !d_u2 || !d_v2)
    if (count <= 0) return count == 0 ? B2PIV_OK : fail(e, B2PIV_ERR_ARG, "negative count");
    rotate_uv_kernel<<<mask_grid(e, count), 256, 0, st>>>(d_u, d_v, count, std::cos(theta), std::sin(theta), d_u2, d_v2);
    MASK_EPILOGUE(1)
}

// ---- two-pass scheme (BASELINE configs[2]; kernels in multipass.cuh, definition in DESIGN.md §8) ----------
int b2piv_predictor_device(b2piv_engine* e, const float* d_u1, const float* d_v1, int n_pairs, int rows1, int cols1, int wy1, int wx1,
                           int oy1, int ox1, short* d_shift, void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    if (!e->planned) return fail(e, B2PIV_ERR_STATE, "b2piv_plan (fine grid) has not been called");
    if (!d_u1 || !d_v1 || !d_shift) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (n_pairs < 1 || rows1 < 1 || cols1 < 1 || wy1 <= oy1 || wx1 <= ox1 || oy1 < 0 || ox1 < 0)
        return fail(e, B2PIV_ERR_ARG, "bad coarse grid");
    if ((e->H - wy1) / (wy1 - oy1) + 1 != rows1 || (e->W - wx1) / (wx1 - ox1) + 1 != cols1)
        return fail(e, B2PIV_ERR_ARG, "coarse field shape does not match the planned frame size");
    if (e->H > 32767 || e->W > 32767) return fail(e, B2PIV_ERR_UNSUPPORTED, "shifts are 16-bit");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t n1 = (size_t)n_pairs * rows1 * cols1;
    int rc = ensure(e, &e->d_mp_ws, &e->cap_mp_ws, 2 * n1 * sizeof(double));
    if (rc) return rc;
    double* vu = e->d_mp_ws;
    double* vv = e->d_mp_ws + n1;
    mp_validate_kernel<<<mask_grid(e, (long long)n1, 128), 128, 0, st>>>(d_u1, d_v1, n_pairs, rows1, cols1, 0.1, 2.0, vu, vv);
    const MpGrid g1{rows1, cols1, wy1, wx1, wy1 - oy1, wx1 - ox1};
    const MpGrid g2{e->n_rows, e->n_cols, e->wy, e->wx, e->wy - e->oy, e->wx - e->ox};
    const long long n2 = (long long)n_pairs * e->n_rows * e->n_cols;
    mp_predictor_kernel<<<mask_grid(e, n2, 128), 128, 0, st>>>(vu, vv, n_pairs, g1, g2, e->H, e->W, d_shift);
    CK(cudaGetLastError());
    e->launches += 2;
    return B2PIV_OK;
}

int b2piv_pairs_shifted_device(b2piv_engine* e, const void* d_frames, long long frame_stride_bytes, int pitch_bytes, int n_frames,
                               const short* d_shift, float* d_u, float* d_v, float* d_corr_max, float* d_s2n, void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    if (!e->planned) return fail(e, B2PIV_ERR_STATE, "b2piv_plan has not been called");
    if (!d_frames || !d_shift || !d_u || !d_v || !d_corr_max || !d_s2n) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (n_frames < 2) return fail(e, B2PIV_ERR_ARG, "need at least 2 frames (one pair)");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Params p = base_params(e, d_frames, frame_stride_bytes, pitch_bytes, n_frames - 1);
    p.shift = d_shift;
    p.u = d_u; p.v = d_v; p.cmax = d_corr_max; p.s2n = d_s2n;
    return dispatch_pairs(e, p, st);
}

void* b2piv_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void b2piv_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int b2piv_last_kernel_ms(const b2piv_engine* e, float* ms) {
    if (!e || !ms) return B2PIV_ERR_ARG;
    *ms = e->last_kernel_ms;
    return B2PIV_OK;
}
long long b2piv_launch_count(const b2piv_engine* e) { return e ? e->launches : 0; }

}  // extern "C"

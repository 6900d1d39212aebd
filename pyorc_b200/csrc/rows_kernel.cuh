// rows_kernel.cuh - the row-per-thread fused PIV kernel (phases in piv_rows.cuh) and its launcher, shared by the k_rows_*.cu
// translation units (each instantiates a few variants).
#pragma once
#include "engine.h"
#include "piv_rows.cuh"

using namespace b2piv;

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
    r.dc_fix[0] = r.dc_fix[1] = 0.f;   // only the magic-number centring of uint8 windows sets it (rows_p2_pre)
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
        // box start = the 16-byte boundary below the window (bytes of uint8 frames, floats of padded float32 frames)
        constexpr int XMASK = (PAD && F32) ? ~3 : ~15;
        const int xa0 = ALIGNED ? un.x0[0] : (un.x0[0] & XMASK), xa1 = ALIGNED ? un.x0[1] : (un.x0[1] & XMASK);
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
            if constexpr (PAD && F32) {   // one un-swizzled (PFW floats x ny rows) box per window
                mbar_expect_tx(&s.mbar, 2u * (uint32_t)p.ny * (uint32_t)(R::PFW * 4));
                tma_load_3d(s.tile(), &tmap, &s.mbar, xa0, un.y0[0], frame);
                tma_load_3d(s.tile() + R::PFWIN, &tmap, &s.mbar, xa1, un.y0[1], frame);
            } else if constexpr (!F32) {
                mbar_expect_tx(&s.mbar, TILE_BYTES);
                tma_load_3d(s.tile(), &tmap, &s.mbar, xa0, un.y0[0], frame);
                tma_load_3d(s.tile() + WIN_BYTES, &tmap, &s.mbar, xa1, un.y0[1], frame);
            } else {
                mbar_expect_tx(&s.mbar, R::F_PHASES == 2 ? R::FWIN : 2 * R::FWIN);
                issue_f32(0, frame, 0);
                if (R::F_PHASES == 1) issue_f32(1, frame, R::FWIN);
            }
        };
        if (nfr > 0 && tid == 0) issue_frame_start(un.f0);
        for (int k = 0; k < maxn; ++k) {
            const bool active = k < nfr;
            const bool have_prev = k > 0;
            const int f = un.f0 + k;
            if (active) {
                while (!mbar_try_wait(&s.mbar, parity)) {}
                parity ^= 1u;
            }
            if constexpr (PAD && F32) {
                rows_f1_pad<R>(s, r, tid, p, 0, xoff0);
                rows_f1_pad<R>(s, r, tid, p, 1, xoff1);
                __syncthreads();  // A: tile (aliased on X) consumed, row sums visible
                rows_f2_pad<R>(s, r, tid, p, 0);
                rows_f2_pad<R>(s, r, tid, p, 1);
                __syncthreads();  // A3: centred second moments visible
                rows_f3_pad<R>(s, r, tid, p);
            } else if constexpr (PAD) {
                static_assert(!PAD || !ALIGNED, "padded mode uses boxes from the 16-byte boundary below the window");
                rows_p1_pad<R>(s, r, tid, p, xoff0, xoff1);
                __syncthreads();  // A
                rows_p2_pre_pad<R>(s, r, tid, p);
            } else if constexpr (!F32) {
                rows_p1<R, ALIGNED>(s, r, tid, xoff0, xoff1);
                __syncthreads();  // A: integer moments visible, tile (aliased on X) fully consumed
                rows_p2_pre<R, true>(s, r, tid, p.clip_norm);
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
                if (active && have_prev) rows_dump_planes<R, PAD>(r, tid, p, un, f - 1);
                rows_p6x<R, PAD>(s, r, tid, &p);   // peak row known from E1: column search + neighbour rows in one phase
                __syncthreads();  // F: neighbour rows dumped, peak column published
                if (active && have_prev) rows_p8<R, PAD>(s, r, tid, p, un, f - 1, true);
            }
            r.half_alpha_prev[0] = r.half_alpha_new[0];
            r.half_alpha_prev[1] = r.half_alpha_new[1];
        }
        __syncthreads();  // round boundary: the next round's first TMA overwrites X
    }
}

// ------------------------------------------------------------------------------------------------------------
// The same pipeline with the parked spectra in Tensor Memory (piv_rows.cuh, rows_p3b_tm): native uint8 windows, per-time-step
// mode.  ONE CTA per SM holds G independent groups (G * W threads); a group synchronises on its own named barrier, fetches
// its tiles on its own mbarrier and walks through its own work units, so the groups drift apart like separate CTAs would -
// but without the 33.8 KB of shared memory per group that capped an SM at four of them, and with one Tensor-Memory
// allocation (all 512 columns) for the CTA.
// ------------------------------------------------------------------------------------------------------------
template <class R, int G, bool ALIGNED>
__global__ void __launch_bounds__(R::NT* G, 1) piv_rows_tm_kernel(const __grid_constant__ CUtensorMap tmap, RParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    constexpr int W = R::W;
    constexpr int COLS = 4 * R::HS;                                   // Tensor-Memory columns per thread: (A0, A1) at ky = 0 .. W/2
    static_assert(((R::NT * G / 32 + 3) / 4) * COLS <= 512, "the warps of a lane quarter must fit into 512 columns");
    __shared__ uint32_t tm_base_s;
    const int g = threadIdx.x / R::NT;     // group within the CTA
    const int tid = threadIdx.x % R::NT;   // thread within the group (= row / column slot)
    const int warp = threadIdx.x >> 5;
    const int bar = g + 1;                 // named barrier of the group (0 is __syncthreads)
    RSmem<R>& s = *reinterpret_cast<RSmem<R>*>(base + (size_t)g * tm_group_stride<R>());
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(&s.mbar, 1);
        fence_mbar_init();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // lane quarter of this warp, column range of this warp within the quarter
    const uint32_t tm = tm_base_s + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(COLS * (warp >> 2));
    uint32_t parity = 0;
    RRegs<R> r;
    r.half_alpha_prev[0] = r.half_alpha_prev[1] = 0.f;
    r.dc_fix[0] = r.dc_fix[1] = 0.f;   // only the magic-number centring of uint8 windows sets it (rows_p2_pre)
    for (long long unit = (long long)blockIdx.x * G + g; unit < p.n_units; unit += (long long)gridDim.x * G) {
        const RUnit un = decode_unit(p, (int)unit);
        const int nfr = un.f1 - un.f0 + 1;
        constexpr int TILE_BYTES = ALIGNED ? R::TILE : R::TILE_U;
        constexpr int WIN_BYTES = TILE_BYTES / 2;
        const int xa0 = ALIGNED ? un.x0[0] : (un.x0[0] & ~15), xa1 = ALIGNED ? un.x0[1] : (un.x0[1] & ~15);
        const int xoff0 = un.x0[0] - xa0, xoff1 = un.x0[1] - xa1;
        auto issue_frame_start = [&](int frame) {
            fence_proxy_async();
            mbar_expect_tx(&s.mbar, TILE_BYTES);
            tma_load_3d(s.tile(), &tmap, &s.mbar, xa0, un.y0[0], frame);
            tma_load_3d(s.tile() + WIN_BYTES, &tmap, &s.mbar, xa1, un.y0[1], frame);
        };
        if (tid == 0 && nfr > 0) issue_frame_start(un.f0);
        for (int k = 0; k < nfr; ++k) {
            const bool have_prev = k > 0;
            const int f = un.f0 + k;
            while (!mbar_try_wait(&s.mbar, parity)) {}
            parity ^= 1u;
            rows_p1<R, ALIGNED>(s, r, tid, xoff0, xoff1);
            group_barrier<true, R::NT>(bar);  // A: integer moments visible, tile (aliased on X) fully consumed
            rows_p2_pre<R, true>(s, r, tid, p.clip_norm);
#pragma unroll 1
            for (int st = 0; st < 4; ++st) {
                fft_reg<W, 0>(r.v);
                if ((st & 1) == 0) transpose_device<R, true>(s, r, tid, st != 0, bar);
                else if (st == 1) rows_p3b_tm<R>(r, tid, tm);
            }
            const bool dead0 = (r.half_alpha_prev[0] == 0.f) || (r.half_alpha_new[0] == 0.f);
            const bool dead1 = (r.half_alpha_prev[1] == 0.f) || (r.half_alpha_new[1] == 0.f);
            rows_p5_post<R, false>(s, r, tid, dead0, dead1, &p);
            group_barrier<true, R::NT>(bar);  // E1: block max / sum / peak row; X (and the tile aliased on it) is free again
            if (tid == 0 && k + 1 < nfr) issue_frame_start(f + 1);
            if (have_prev) rows_dump_planes<R, false>(r, tid, p, un, f - 1);
            rows_p6x<R, false>(s, r, tid, &p);
            group_barrier<true, R::NT>(bar);  // F: neighbour rows dumped, peak column published
            if (have_prev) rows_p8<R, false>(s, r, tid, p, un, f - 1, true);
            r.half_alpha_prev[0] = r.half_alpha_new[0];
            r.half_alpha_prev[1] = r.half_alpha_new[1];
        }
        group_barrier<true, R::NT>(bar);  // unit boundary: the next unit's first TMA overwrites X
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_base_s) : "memory");
}

// ---- tensor map + launch ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline PFN_encodeTiled get_encode_tiled() {
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

__global__ void planes_reorder_kernel(const float* __restrict__ nat, float* __restrict__ out, long long n_planes, int W, int ny, int nx);

template <class R, int G, bool ROLLED, bool ALIGNED, bool F32, bool ENS = false, bool PAD = false>
static int launch_rows(b2piv_engine* e, const Params& gp, cudaStream_t st, const EnsParams* ep = nullptr) {
    constexpr int W = R::W;
    const int n_frames = gp.n_pairs + 1;
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {(cuuint64_t)e->W, (cuuint64_t)e->H, (cuuint64_t)n_frames};
    const cuuint64_t strides[2] = {(cuuint64_t)gp.pitch, (cuuint64_t)gp.frame_stride};
    constexpr bool FPAD = F32 && PAD;   // padded float32 mode: (PFW floats x ny rows) un-swizzled boxes
    const cuuint32_t box[3] = {(cuuint32_t)(FPAD ? R::PFW : (F32 ? 32 : (ALIGNED ? W : R::WB))), (cuuint32_t)(FPAD ? e->wy : W), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle swz = FPAD ? CU_TENSOR_MAP_SWIZZLE_NONE : F32 ? CU_TENSOR_MAP_SWIZZLE_128B
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
    p.fshift = gp.fshift; p.pair_step = gp.pair_step;
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
    if (run <= 0) run = pick_run_len(gp.n_pairs, n_wp, resident);   // engine.h
    if (run > gp.n_pairs || ENS) run = gp.n_pairs;   // ensemble: one unit owns its windows' accumulators for the whole launch
    if (gp.pair_step == 2) run = 1;                  // interleaved stack: a unit is one (a_k, b_k) pair; gp.n_pairs = 2 P - 1
    p.run_len = run;
    long long n_units = gp.pair_step == 2 ? (long long)n_wp * ((gp.n_pairs + 1) / 2) : (long long)n_wp * ((gp.n_pairs + run - 1) / run);
    long long grid = (n_units + G - 1) / G;
    if (grid > (long long)occ * e->sm_count) grid = (long long)occ * e->sm_count;
    if (G == 1 && !ENS && e->run_len <= 0 && gp.pair_step != 2) {
        // independent groups: even partition of the work instead of waves of equal units when its longest part is shorter
        // (option "unit_parts" forces a partition into that many parts - tests)
        const long long waves = (n_units + resident - 1) / resident;
        const int parts = e->force_parts > 0 ? e->force_parts : (int)resident;
        int tn = 0;
        long long tcost = 0;
        const int* tab = build_unit_table(e, n_wp, gp.n_pairs, parts, st, &tn, &tcost, e->force_parts > 0);
        if (tab && (e->force_parts > 0 || tcost < waves * (run + 1))) { p.unit_table = tab; n_units = tn; grid = parts; }
    }
    p.n_units = (int)n_units;
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

// launch of the Tensor-Memory variant (native uint8 windows, per-time-step mode)
template <class R, int G, bool ALIGNED>
static int launch_rows_tm(b2piv_engine* e, const Params& gp, cudaStream_t st) {
    constexpr int W = R::W;
    const int n_frames = gp.n_pairs + 1;
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {(cuuint64_t)e->W, (cuuint64_t)e->H, (cuuint64_t)n_frames};
    const cuuint64_t strides[2] = {(cuuint64_t)gp.pitch, (cuuint64_t)gp.frame_stride};
    const cuuint32_t box[3] = {(cuuint32_t)(ALIGNED ? W : R::WB), (cuuint32_t)W, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle swz = !ALIGNED ? CU_TENSOR_MAP_SWIZZLE_NONE : (W == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const CUresult cr = get_encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(gp.frames), dims, strides, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(e, B2PIV_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)cr));
    RParams p;
    memset(&p, 0, sizeof(p));
    p.frames = (const unsigned char*)gp.frames; p.frame_stride = gp.frame_stride; p.pitch = gp.pitch;
    p.n_rows = gp.n_rows; p.n_cols = gp.n_cols; p.sy = gp.sy; p.sx = gp.sx; p.n_pairs = gp.n_pairs;
    p.clip_norm = gp.clip_norm; p.border_nan = gp.border_nan; p.gauss_eps = gp.gauss_eps; p.keep = gp.keep;
    p.u = gp.u; p.v = gp.v; p.cmax = gp.cmax; p.s2n = gp.s2n; p.planes = gp.planes; p.peer = gp.peer;
    p.ny = p.nx = W;
    const size_t smem = tm_group_stride<R>() * G + 1024;
    auto kern = piv_rows_tm_kernel<R, G, ALIGNED>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long resident = (long long)e->sm_count * G;
    const int nw = gp.n_rows * gp.n_cols, n_wp = (nw + 1) / 2;
    int run = e->run_len;
    if (run <= 0) run = pick_run_len(gp.n_pairs, n_wp, resident);
    if (run > gp.n_pairs) run = gp.n_pairs;
    p.run_len = run;
    long long n_units = (long long)n_wp * ((gp.n_pairs + run - 1) / run);
    long long grid = (n_units + G - 1) / G;
    if (grid > e->sm_count) grid = e->sm_count;
    if (e->run_len <= 0) {
        const long long waves = (n_units + resident - 1) / resident;
        const bool forced = e->force_parts >= G;
        const int parts = forced ? (e->force_parts / G) * G : (int)resident;
        int tn = 0;
        long long tcost = 0;
        const int* tab = build_unit_table(e, n_wp, gp.n_pairs, parts, st, &tn, &tcost, forced);
        if (tab && (forced || tcost < waves * (run + 1))) { p.unit_table = tab; n_units = tn; grid = parts / G; }
    }
    p.n_units = (int)n_units;
    kern<<<(unsigned)grid, R::NT * G, smem, st>>>(tmap, p);
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}

// k_rows128.cu - 128x128 windows (and, zero-padded, even windows of 34 .. 64 px) on the polyphase row-per-thread kernel
// (piv_rows128.cuh), uint8 and float32 frames.
#include "rows_kernel.cuh"
#include "piv_rows128.cuh"

// 128 x 128 windows: four polyphase sub-groups of 64 threads run the 64 x 64 pipeline above and meet in the cross-spectrum
// phase (piv_rows128.cuh).  One CTA = one group of 256 threads = one pair of adjacent windows followed through a run of
// frames; 213 KB of shared memory (4 x transpose blocks + exchange of the new spectra), the parked spectra in Tensor Memory,
// one CTA per SM.
template <bool ENS, bool PAD, bool F32>
__global__ void __launch_bounds__(256, 1) piv_rows128_kernel(const __grid_constant__ CUtensorMap tmap, RParams p) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    R128Smem& s = *reinterpret_cast<R128Smem*>(base);
    const int tid = threadIdx.x;
    const int sub = tid >> 6;      // polyphase component (p1, p2) = (sub >> 1, sub & 1)
    const int t = tid & 63;        // thread within the sub-group (= line slot of the 64 x 64 pipeline)
    RSmem<R6>& ss = s.sub[sub];
    __shared__ uint32_t tm_base_s;
    const int warp = tid >> 5;
    static_assert(2 * R128_TM_COLS <= 512, "the two warps of a lane quarter must fit into 512 columns");
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
    // lane quarter of this warp, column range of this warp within the quarter (piv_rows_tm_kernel)
    const uint32_t tm = tm_base_s + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(R128_TM_COLS * (warp >> 2));
    uint32_t parity = 0;
    RRegs<R6> r;
    r.half_alpha_prev[0] = r.half_alpha_prev[1] = 0.f;
    for (long long unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const RUnit un = decode_unit(p, (int)unit);
        const int nfr = un.f1 - un.f0 + 1;
        constexpr int XAL = F32 ? 3 : 15;                          // padded mode: boxes from the 16-byte boundary below (4 floats / 16 bytes)
        const int xa0 = un.x0[0] & ~XAL, xa1 = un.x0[1] & ~XAL;
        const int xoff0 = un.x0[0] - xa0, xoff1 = un.x0[1] - xa1;
        auto issue_frame = [&](int frame) {
            fence_proxy_async();
            if constexpr (F32 && PAD) {
                mbar_expect_tx(&s.mbar, 2 * R128_PFWIN);
                tma_load_3d(r128_ftile(s, 0), &tmap, &s.mbar, xa0, un.y0[0], frame);
                tma_load_3d(r128_ftile(s, 1), &tmap, &s.mbar, xa1, un.y0[1], frame);
            } else if constexpr (F32) {
                mbar_expect_tx(&s.mbar, 2 * 128 * 128 * 4);
#pragma unroll
                for (int w = 0; w < 2; ++w)
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb)
#pragma unroll
                        for (int h = 0; h < 4; ++h)
                            tma_load_3d(r128_ftile(s, 2 * w + hb) + h * 8192, &tmap, &s.mbar, un.x0[w] + 32 * h, un.y0[w] + 64 * hb, frame);
            } else if constexpr (PAD) {
                mbar_expect_tx(&s.mbar, 2 * R128_PWIN);
                tma_load_3d(s.sub[0].tile(), &tmap, &s.mbar, xa0, un.y0[0], frame);
                tma_load_3d(s.sub[0].tile() + R128_PWIN, &tmap, &s.mbar, xa1, un.y0[1], frame);
            } else {
                mbar_expect_tx(&s.mbar, 2 * 128 * 128);
#pragma unroll
                for (int w = 0; w < 2; ++w)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        tma_load_3d(s.sub[j].tile() + w * 4096, &tmap, &s.mbar, un.x0[w], un.y0[w] + 32 * j, frame);
            }
        };
        if (tid == 0) issue_frame(un.f0);
        for (int k = 0; k < nfr; ++k) {
            const bool have_prev = k > 0;
            const int f = un.f0 + k;
            while (!mbar_try_wait(&s.mbar, parity)) {}
            parity ^= 1u;
            if constexpr (F32 && PAD) {
                r128_f1_pad(s, r, sub, t, p, 0, xoff0);
                r128_f1_pad(s, r, sub, t, p, 1, xoff1);
                __syncthreads();  // A: row sums visible, tile (in the spectrum blocks of sub-groups 0 / 1) fully consumed
                r128_f2_pad(s, r, sub, t, p, 0);
                r128_f2_pad(s, r, sub, t, p, 1);
                __syncthreads();  // A2: centred second moments visible
                r128_f3_pad(s, r, t, p);
            } else if constexpr (F32) {
                r128_f1(s, r, sub, t, 0);
                r128_f1(s, r, sub, t, 1);
                __syncthreads();  // A: row sums visible, tile (in the spectrum blocks) fully consumed
                r128_f2(s, r, sub, t, 0);
                r128_f2(s, r, sub, t, 1);
                __syncthreads();  // A2: centred second moments visible
                r128_f3(s, r, p.clip_norm);
            } else {
                if constexpr (PAD) r128_p1_pad(s, r, sub, t, p, xoff0, xoff1); else r128_p1(s, r, sub, t);
                __syncthreads();  // A: integer moments visible, tile (aliased on the transpose blocks) fully consumed
                if constexpr (PAD) r128_p2_pad(s, r, t, p); else r128_p2(s, r, p.clip_norm);
            }
            // forward: FFT(rows) T FFT(cols) per component; cross spectra across components; inverse: FFT(cols) T FFT(rows)
            // (one copy of the unrolled FFT: the loop body is far beyond the instruction caches, every KB counts)
#pragma unroll 1
            for (int stg = 0; stg < 4; ++stg) {
                fft_reg<64, 0>(r.v);
                if ((stg & 1) == 0) transpose_device<R6>(ss, r, t, stg != 0);
                else if (stg == 1) r128_cross<PAD>(s, r, sub, t, tm, p);
            }
            const bool dead0 = (r.half_alpha_prev[0] == 0.f) || (r.half_alpha_new[0] == 0.f);
            const bool dead1 = (r.half_alpha_prev[1] == 0.f) || (r.half_alpha_new[1] == 0.f);
            rows_p5_post<R6, PAD>(ss, r, t, dead0, dead1, &p);
            __syncthreads();  // E1: block max / sum of all components; the transpose blocks are free again
            if (!(ENS && F32) && tid == 0 && k + 1 < nfr) issue_frame(f + 1);
            if constexpr (ENS) {
                if constexpr (PAD) r128_ens_pad(s, r, sub, t, tid, p, un, f - 1, have_prev);
                else r128_ens(s, r, sub, t, tid, p, un, f - 1, have_prev);   // thresholds + accumulate; no peak search per pair
                if constexpr (F32) {   // the float32 tile lands where the planes were staged
                    __syncthreads();
                    if (tid == 0 && k + 1 < nfr) issue_frame(f + 1);
                }
            } else {
                if constexpr (PAD) r128_p6_pad(s, r, sub, t, p); else r128_p6(s, r, sub, t);
                __syncthreads();  // E2: first-argmax keys
                if (have_prev) { if constexpr (PAD) r128_dump_planes_pad(r, sub, t, p, un, f - 1); else r128_dump_planes(r, sub, t, p, un, f - 1); }
                if constexpr (PAD) r128_p7_pad(s, r, sub, t, p); else r128_p7(s, r, sub, t);
                __syncthreads();  // F: neighbour rows dumped
                if (have_prev) { if constexpr (PAD) r128_p8(s, r, tid, p, un, f - 1, 2 * p.ny, 2 * p.nx); else r128_p8(s, r, tid, p, un, f - 1); }
            }
            r.half_alpha_prev[0] = r.half_alpha_new[0];
            r.half_alpha_prev[1] = r.half_alpha_new[1];
        }
        __syncthreads();  // unit boundary: the next unit's first TMA overwrites the transpose blocks
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_base_s) : "memory");
}

int launch_rows128(b2piv_engine* e, const Params& gp, cudaStream_t st, const EnsParams* ep, bool pad) {
    const int n_frames = gp.n_pairs + 1;
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {(cuuint64_t)e->W, (cuuint64_t)e->H, (cuuint64_t)n_frames};
    const cuuint64_t strides[2] = {(cuuint64_t)gp.pitch, (cuuint64_t)gp.frame_stride};
    const bool f32 = e->dtype == B2PIV_F32;
    const cuuint32_t box[3] = {(cuuint32_t)(pad ? (f32 ? R128_PFW : R128_PWB) : (f32 ? 32 : 128)), (cuuint32_t)((pad || f32) ? 64 : 32), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = get_encode_tiled()(&tmap, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(gp.frames), dims, strides, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, pad ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(e, B2PIV_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)cr));
    RParams p;
    memset(&p, 0, sizeof(p));
    p.frames = (const unsigned char*)gp.frames; p.frame_stride = gp.frame_stride; p.pitch = gp.pitch;
    p.n_rows = gp.n_rows; p.n_cols = gp.n_cols; p.sy = gp.sy; p.sx = gp.sx; p.n_pairs = gp.n_pairs;
    p.clip_norm = gp.clip_norm; p.border_nan = gp.border_nan; p.gauss_eps = gp.gauss_eps; p.keep = gp.keep;
    p.u = gp.u; p.v = gp.v; p.cmax = gp.cmax; p.s2n = gp.s2n; p.planes = gp.planes; p.peer = gp.peer;
    p.ny = p.nx = 128;
    if (pad) {   // component size; spectrum factor of the 2 x 2 tiling and masks (piv_rows128.cuh, "Padded mode")
        p.ny = e->wy / 2; p.nx = e->wx / 2;
        p.pad_scale = (float)(1.0 / (4096.0 * e->wy * e->wx));
        const double two_pi = 6.283185307179586476925286766559;
        for (int k = 0; k <= 32; ++k) { const double th = two_pi * (double)((k * p.ny) % 64) / 64; p.pad_ty[k] = make_float2((float)(1.0 + cos(th)), (float)(-sin(th))); }
        for (int k = 0; k < 64; ++k) { const double th = two_pi * (double)((k * p.nx) % 64) / 64; p.pad_tx[k] = make_float2((float)(1.0 + cos(th)), (float)(-sin(th))); }
        for (int k = 0; k < 16; ++k) { const int left = e->wx - 4 * k; p.pad_mask[k] = left >= 4 ? 0xffffffffu : (left <= 0 ? 0u : ((1u << (8 * left)) - 1u)); }
        for (int x = 0; x < 64; ++x) p.pad_cm[x] = x < p.nx ? 1.f : 0.f;
    }
    if (ep) { p.corr_min = ep->corr_min; p.s2n_min = ep->s2n_min; p.ens_sum = ep->plane_sum; p.ens_count = ep->count; }
    auto kern = pad ? (f32 ? (ep ? piv_rows128_kernel<true, true, true> : piv_rows128_kernel<false, true, true>)
                           : (ep ? piv_rows128_kernel<true, true, false> : piv_rows128_kernel<false, true, false>))
                    : (f32 ? (ep ? piv_rows128_kernel<true, false, true> : piv_rows128_kernel<false, false, true>)
                           : (ep ? piv_rows128_kernel<true, false, false> : piv_rows128_kernel<false, false, false>));
    const size_t smem = sizeof(R128Smem) + 1024;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
    if (occ < 1) return fail(e, B2PIV_ERR_CUDA, "128x128 rows kernel does not fit on an SM");
    const long long resident = (long long)occ * e->sm_count;
    const int nw = gp.n_rows * gp.n_cols, n_wp = (nw + 1) / 2;
    int run = e->run_len;
    if (run <= 0) run = pick_run_len(gp.n_pairs, n_wp, resident);   // engine.h
    if (run > gp.n_pairs || ep) run = gp.n_pairs;   // ensemble: one unit owns its windows' accumulators for the whole launch
    p.run_len = run;
    const long long n_units = (long long)n_wp * ((gp.n_pairs + run - 1) / run);
    p.n_units = (int)n_units;
    long long grid = n_units < resident ? n_units : resident;
    kern<<<(unsigned)grid, 256, smem, st>>>(tmap, p);
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
}

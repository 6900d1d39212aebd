// k_rows_shift.cu - displaced second pass of the two-pass scheme on the row-per-thread machinery (32x32 uint8 windows).
#include "rows_kernel.cuh"

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

int launch_rows_shift(b2piv_engine* e, const Params& gp, cudaStream_t st) {
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
    if (run <= 0) run = pick_run_len(gp.n_pairs, n_wp, resident);   // engine.h
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

// abi_piv.cu - C ABI (include/b2piv.h) of the PIV path: plan, per-time-step and ensemble entry points, kernel dispatch, the
// host pipeline (chunked H2D overlapped with compute, page-locked staging of pageable frames).  sm_100a only.
//
// Replaces the ffpiv/rocket-fft CPU path that pyorc/velocimetry/ffpiv.py calls (cross_corr, u_v_displacement)
// with ONE fused kernel per frame-pair batch: window gather -> normalise -> packed complex 2-D FFT -> cross
// spectrum -> inverse FFT -> fftshift,/N,clip -> max / mean / first-argmax -> 3-point Gaussian sub-pixel fit.
#include "engine.h"

using namespace b2piv;

std::string g_create_err;

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

// px / frame -> m / s on the device, the arithmetic of `u * res_x / dt` in numpy for a float32 `u` (pyorc/velocimetry/ffpiv.py:418-419):
// float32 product with the float32-rounded resolution, float64 division by the pair's dt, one rounding to float32
__global__ void __launch_bounds__(256) units_kernel(float* __restrict__ u, float* __restrict__ v, long long n, long long nw, float res_x, float res_y,
                                                    const double* __restrict__ dt) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double d = dt[i / nw];
        u[i] = (float)((double)__fmul_rn(u[i], res_x) / d);
        v[i] = (float)((double)__fmul_rn(v[i], res_y) / d);
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

static bool supported(int wy, int wx) { return fft_config(wy, wx) || (wy >= 4 && wx >= 4 && wy <= 128 && wx <= 128); }
// a side above 64 px that is not a compiled FFT shape: large-window direct kernel (k_direct.cu)
static bool big_direct(const b2piv_engine* e) { return !fft_config(e->wy, e->wx) && (e->wy > 64 || e->wx > 64); }

static bool rows_eligible(const b2piv_engine* e, const void* d_frames, long long frame_stride, int pitch) {
    if (e->wy != e->wx || (e->wy != 64 && e->wy != 32)) return false;   // uint8 and float32 frames both qualify
    if ((pitch & 15) || (frame_stride & 15) || (((uintptr_t)d_frames) & 15)) return false;
    // every TMA box must start on a 16-byte boundary in global memory: x strides that are a multiple of 16 use exact
    // swizzled boxes, multiples of 4 (32x32 at 75 % overlap: stride 8) a 16-byte wider box read at an offset
    if ((e->wx - e->ox) & 3) return false;
    return tma_available();
}

// displaced second pass on the row-per-thread kernel: square 32x32 uint8 windows, x stride a multiple of 4
static bool rows_shift_eligible(const b2piv_engine* e, const void* d_frames, long long frame_stride, int pitch) {
    if (e->wy != 32 || e->wx != 32 || e->dtype != B2PIV_U8) return false;
    if ((pitch & 15) || (frame_stride & 15) || (((uintptr_t)d_frames) & 15)) return false;
    if ((e->wx - e->ox) & 3) return false;
    return tma_available();
}

// 128 x 128 uint8 windows on the polyphase row-per-thread kernel (piv_rows128.cuh)
static bool rows128_eligible(const b2piv_engine* e, const void* d_frames, long long frame_stride, int pitch) {
    if (e->wy != 128 || e->wx != 128) return false;   // uint8 and float32 frames both qualify
    if ((pitch & 15) || (frame_stride & 15) || (((uintptr_t)d_frames) & 15)) return false;
    if ((e->wx - e->ox) & (e->dtype == B2PIV_F32 ? 3 : 15)) return false;   // swizzled boxes start on 16-byte boundaries
    return tma_available();
}

// Padded mode of the row-per-thread kernel: any window (uint8 or float32 frames, square or not, any stride) whose larger side is
// at most 32 px, i.e. at most half of a 64 x 64 (or 32 x 32) plane.  Native 32 x 32 / 64 x 64 windows never come here.
static bool pad_eligible(const b2piv_engine* e, const void* d_frames, long long frame_stride, int pitch) {
    const int m = e->wy > e->wx ? e->wy : e->wx;
    if (2 * m > 64 || e->wy < 2 || e->wx < 2) return false;
    if ((pitch & 15) || (frame_stride & 15) || (((uintptr_t)d_frames) & 15)) return false;
    return tma_available();
}

// Padded mode of the 128-plane polyphase kernel: even windows (uint8 or float32 frames) whose larger side is 34 .. 64 px
// (piv_rows128.cuh)
static bool pad128_eligible(const b2piv_engine* e, const void* d_frames, long long frame_stride, int pitch) {
    const int m = e->wy > e->wx ? e->wy : e->wx;
    if (m <= 32 || m > 64 || (e->wy & 1) || (e->wx & 1) || e->wy < 2 || e->wx < 2) return false;
    if (fft_config(e->wy, e->wx)) return false;   // compiled FFT shapes (64x64, 64x32 ...) have their own kernels
    if ((pitch & 15) || (frame_stride & 15) || (((uintptr_t)d_frames) & 15)) return false;
    if (e->dtype == B2PIV_F32 && e->W < 68) return false;   // the float32 box is 68 floats wide (R128_PFW); narrower frames: shared-memory kernel
    return tma_available();
}

int dispatch_pairs(b2piv_engine* e, const Params& p, cudaStream_t st) {
    if (p.n_pairs <= 0) return B2PIV_OK;
    // displaced second-pass windows (multipass.cuh) break the frame-to-frame spectrum sharing of the row-per-thread kernels
    if (p.shift) {
        if (e->variant != 1 && rows_shift_eligible(e, p.frames, p.frame_stride, p.pitch)) {
            e->last_variant = 2;
            return launch_rows_shift(e, p, st);
        }
        if (e->variant == 2) return fail(e, B2PIV_ERR_UNSUPPORTED, "displaced rows kernel needs square 32x32 uint8 windows, 16-byte aligned base/pitch and an x stride that is a multiple of 4");
        e->last_variant = 1;
        return launch_generic(e, p, st);
    }
    const bool can_rows = rows_eligible(e, p.frames, p.frame_stride, p.pitch);
    if (e->variant == 2 && !can_rows && e->wy != 128)
        return fail(e, B2PIV_ERR_UNSUPPORTED, "rows kernel needs a square 32/64 window, 16-byte aligned base/pitch and an x stride that is a multiple of 4");
    if (e->variant == 2 && e->wy == 128 && !rows128_eligible(e, p.frames, p.frame_stride, p.pitch))
        return fail(e, B2PIV_ERR_UNSUPPORTED, "128x128 rows kernel needs 16-byte aligned base/pitch and an x stride of a multiple of 16 bytes");
    if (rows128_eligible(e, p.frames, p.frame_stride, p.pitch) && (e->variant == 0 || e->variant == 2)) {
        e->last_variant = 2;
        return launch_rows128(e, p, st);
    }
    if (can_rows && e->variant != 1) {
        e->last_variant = 2;
        return e->dtype == B2PIV_F32 ? launch_rows_f32(e, p, st) : launch_rows_u8(e, p, st);
    }
    // sizes that are not a compiled FFT shape: uint8 windows up to 32 px run zero-padded through the row-per-thread kernel
    // (piv_rows.cuh "Padded mode"; variant 4 forces it wherever it applies; measured 3.3x the direct kernel at 10x10 and
    // 2.6x the shared-memory kernel at 26x26).  Otherwise (float32 frames, caller-owned tensors with an odd pitch, larger
    // windows): tiny windows by direct correlation, the rest padded through the shared-memory FFT kernel; variant 3
    // forces the direct kernel
    if (big_direct(e)) {
        if (p.shift) return fail(e, B2PIV_ERR_UNSUPPORTED, "displaced windows need a window of at most 64 px per side");
        e->last_variant = 3;
        return launch_direct_big(e, p, st);
    }
    if (pad128_eligible(e, p.frames, p.frame_stride, p.pitch) && (e->variant == 0 || e->variant == 4)) {
        e->last_variant = 4;
        return launch_rows128(e, p, st, nullptr, true);
    }
    const bool tiny = !fft_config(e->wy, e->wx) && e->wy * e->wx <= 144;
    if (e->variant == 4 && !pad_eligible(e, p.frames, p.frame_stride, p.pitch))
        return fail(e, B2PIV_ERR_UNSUPPORTED, "padded rows kernels need a window of at most 32 px (uint8 frames: an even one of at most 64 px) and 16-byte aligned base/pitch");
    if (pad_eligible(e, p.frames, p.frame_stride, p.pitch) && (e->variant == 4 || (e->variant == 0 && !fft_config(e->wy, e->wx)))) {
        e->last_variant = 4;
        return e->dtype == B2PIV_F32 ? launch_rows_pad_f32(e, p, st, nullptr) : launch_rows_pad(e, p, st, nullptr);
    }
    if ((e->variant == 3 || tiny) && e->wy <= 64 && e->wx <= 64) {
        e->last_variant = 3;
        return launch_direct(e, p, st);
    }
    e->last_variant = 1;
    return launch_generic(e, p, st);
}
int dispatch_ens(b2piv_engine* e, const Params& p, const EnsParams& ep, cudaStream_t st) {
    if (p.n_pairs <= 0) return B2PIV_OK;
    const bool can_rows = rows_eligible(e, p.frames, p.frame_stride, p.pitch);
    if (e->variant == 2 && !can_rows && !rows128_eligible(e, p.frames, p.frame_stride, p.pitch))
        return fail(e, B2PIV_ERR_UNSUPPORTED, "rows kernel needs a square 32/64 window (or 128x128 uint8), 16-byte aligned base/pitch and an x stride that is a multiple of 4");
    if (can_rows && e->variant != 1 && e->variant != 3) {
        e->last_variant = 2;
        return launch_rows_ens(e, p, ep, st);
    }
    if (rows128_eligible(e, p.frames, p.frame_stride, p.pitch) && (e->variant == 0 || e->variant == 2)) {
        e->last_variant = 2;
        return launch_rows128(e, p, st, &ep);     // polyphase kernel, ensemble epilogue
    }
    if (pad128_eligible(e, p.frames, p.frame_stride, p.pitch) && (e->variant == 0 || e->variant == 4)) {
        e->last_variant = 4;
        return launch_rows128(e, p, st, &ep, true);
    }
    const bool tiny = !fft_config(e->wy, e->wx) && e->wy * e->wx <= 144;
    if (e->variant == 4 && !pad_eligible(e, p.frames, p.frame_stride, p.pitch))
        return fail(e, B2PIV_ERR_UNSUPPORTED, "padded rows kernel needs a window of at most 32 px and 16-byte aligned base/pitch");
    if (pad_eligible(e, p.frames, p.frame_stride, p.pitch) && (e->variant == 4 || (e->variant == 0 && !fft_config(e->wy, e->wx)))) {
        e->last_variant = 4;
        return e->dtype == B2PIV_F32 ? launch_rows_pad_f32(e, p, st, &ep) : launch_rows_pad(e, p, st, &ep);
    }
    e->last_variant = 1;
    if (big_direct(e)) return launch_direct_big_ens(e, p, ep, st);
    if ((e->variant == 3 || tiny) && e->wy <= 64 && e->wx <= 64) return launch_direct_ens(e, p, ep, st);
    return launch_generic_ens(e, p, ep, st);
}

Params base_params(const b2piv_engine* e, const void* d_frames, long long frame_stride, int pitch, int n_pairs) {
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
        cudaEventCreate(&e->ev_k0) != cudaSuccess || cudaEventCreate(&e->ev_k1) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_ens, cudaEventDisableTiming) != cudaSuccess) {
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
    for (auto& t : e->unit_tables) cudaFree(t.d);
    cudaFree(e->d_frames); cudaFree(e->d_out); cudaFree(e->d_planes);
    cudaFree(e->d_keep); cudaFree(e->d_ens_sum); cudaFree(e->d_ens_cnt); cudaFree(e->d_pre_mean); cudaFree(e->d_pre_mm); cudaFree(e->d_mask_ws); cudaFree(e->d_mp_ws);
    cudaFree(e->d_proj_off); cudaFree(e->d_proj_src); cudaFree(e->d_planes_nat); cudaFree(e->d_direct_ws); cudaFree(e->d_dt);
    for (auto ev : e->ev_chunk) cudaEventDestroy(ev);
    for (int i = 0; i < 3; ++i) { if (e->h_stage[i]) cudaFreeHost(e->h_stage[i]); if (e->ev_stage[i]) cudaEventDestroy(e->ev_stage[i]); }
    delete e->pool;
    delete e->stager;
    if (e->h_ring) cudaFreeHost(e->h_ring);
    for (auto ev : e->ev_ring) cudaEventDestroy(ev);
    if (e->ev_k0) cudaEventDestroy(e->ev_k0);
    if (e->ev_k1) cudaEventDestroy(e->ev_k1);
    if (e->ev_ens) cudaEventDestroy(e->ev_ens);
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
    else if (n == "stage_threads") {
        e->stage_threads = value < 0 ? 0 : (int)value;
        delete e->pool; e->pool = nullptr;
        delete e->stager; e->stager = nullptr;
    }
    else if (n == "stage_mode") e->stage_mode = value != 0.0;                           // 1: Stager (stager.h), 0: round 1's pool
    else if (n == "stage_slice_kb") e->stage_slice_kb = value < 4 ? 4 : (int)value;     // per worker and copy
    else if (n == "stage_groups") e->stage_groups = value < 2 ? 2 : (int)value;         // H2D copies in flight (ring depth)
    else if (n == "stage_nt") e->stage_nt = value != 0.0;                               // non-temporal stores into the ring
    else if (n == "kernel_variant") e->variant = (int)value;
    else if (n == "run_len") e->run_len = value < 0 ? 0 : (int)value;
    else if (n == "tmem") e->tmem = value != 0.0;
    else if (n == "unit_parts") e->force_parts = value < 0 ? 0 : (int)value;
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
                                                  " not supported (any size 4..128 per axis)");
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
    NvtxRange nvtx_range("b2piv_pairs_device");
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
    // pageable source (plain numpy memory): staged through a page-locked ring by the engine's copy threads
    cudaPointerAttributes attr;
    bool pageable = true;
    if (cudaPointerGetAttributes(&attr, frames) == cudaSuccess) pageable = (attr.type == cudaMemoryTypeUnregistered);
    else cudaGetLastError();
    if (pageable && e->stage_mode == 1) {
        // Stager (stager.h): slices of ~stage_slice_kb per worker, one H2D per group of slices, `stage_groups` groups in flight;
        // this thread issues the copies, hands landed slots back and launches a chunk's kernel once its frames are under way
        CK(cudaStreamSynchronize(e->s_copy));   // no earlier call may still be reading the ring
        if (!e->stager) {
            int nt = e->stage_threads;
            if (nt <= 0) { nt = (int)std::thread::hardware_concurrency(); nt = nt > 8 ? 8 : (nt < 1 ? 1 : nt); }
            try {
                e->stager = new Stager(nt);
            } catch (...) {   // no threads to be had: let the driver stage the pageable copy (nothing may throw across the ABI)
                e->stager = nullptr;
                pageable = false;
            }
        }
    }
    if (pageable && e->stage_mode == 1) {
        Stager::Job job;
        job.src = (const unsigned char*)frames;
        job.row_bytes = row_bytes;
        job.rows = (size_t)n_frames * e->H;
        job.slice_rows = ((size_t)e->stage_slice_kb << 10) / row_bytes;
        if (job.slice_rows < 1) job.slice_rows = 1;
        job.parts = e->stager->size();
        // a short call is not cut finer than it has to be: every worker gets one slice of the first group
        if (job.slice_rows * (size_t)job.parts > job.rows) job.slice_rows = (job.rows + job.parts - 1) / job.parts;
        job.ring_groups = e->stage_groups < 2 ? 2 : (e->stage_groups > 64 ? 64 : e->stage_groups);
        job.nt = e->stage_nt != 0;
        const size_t need_bytes = Stager::ring_bytes(job);
        if (e->cap_ring < need_bytes) {
            if (e->h_ring) { CK(cudaFreeHost(e->h_ring)); e->h_ring = nullptr; e->cap_ring = 0; }
            CK(cudaHostAlloc((void**)&e->h_ring, need_bytes, cudaHostAllocDefault));
            e->cap_ring = need_bytes;
        }
        job.ring = e->h_ring;
        while ((int)e->ev_ring.size() < job.ring_groups) {
            cudaEvent_t ev;
            CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            e->ev_ring.push_back(ev);
        }
        CK(cudaEventRecord(e->ev_k0, e->s_comp));
        int c = 0;   // next chunk to launch
        const int G = job.ring_groups;
        const int rc2 = e->stager->run(
            job,
            [&](size_t g, unsigned char* slot, size_t row0, size_t nrows) -> int {
                if (dpitch == row_bytes)
                    CK(cudaMemcpyAsync(e->d_frames + row0 * dpitch, slot, nrows * row_bytes, cudaMemcpyHostToDevice, e->s_copy));
                else   // frames are contiguous on both sides, so any run of rows is one 2-D copy
                    CK(cudaMemcpy2DAsync(e->d_frames + row0 * dpitch, dpitch, slot, row_bytes, row_bytes, nrows, cudaMemcpyHostToDevice, e->s_copy));
                CK(cudaEventRecord(e->ev_ring[g % (size_t)G], e->s_copy));
                return B2PIV_OK;
            },
            [&](size_t g) -> bool {
                const cudaError_t q = cudaEventQuery(e->ev_ring[g % (size_t)G]);
                if (q == cudaSuccess) return true;
                cudaGetLastError();                       // cudaErrorNotReady is recorded as the last error: clear it
                return q != cudaErrorNotReady;            // a real error: stop holding the slot back, the next CUDA call reports it
            },
            [&](size_t rows_issued) -> int {
                while (c < chunks) {
                    const int p0 = c * per;
                    const int p1 = (p0 + per < n_pairs) ? p0 + per : n_pairs;
                    if (p0 >= p1) { c = chunks; break; }
                    if ((size_t)(p1 + 1) * e->H > rows_issued) break;      // frames [0, p1] must be under way
                    CK(cudaEventRecord(e->ev_chunk[c], e->s_copy));
                    CK(cudaStreamWaitEvent(e->s_comp, e->ev_chunk[c], 0));
                    const int r3 = work(p0, p1 - p0);
                    if (r3) return r3;
                    ++c;
                }
                return B2PIV_OK;
            });
        if (rc2) return rc2;
        CK(cudaEventRecord(e->ev_k1, e->s_comp));
        return B2PIV_OK;
    }
    // "stage_mode" = 0, round 1's pool: ~8 MB per stage chunk, three buffers, so that the host copy of the next stage chunk
    // overlaps the H2D of the previous ones
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

static int pairs_host_impl(b2piv_engine* e, const void* frames, int n_frames, float signal_threshold, float* u, float* v, float* corr_max, float* s2n,
                           const double* dt, float res_x, float res_y) {
    if (!e) return B2PIV_ERR_ARG;
    NvtxRange nvtx_range("b2piv_pairs_host");
    if (dt && !(res_x > 0.f && res_y > 0.f)) return fail(e, B2PIV_ERR_ARG, "resolution must be positive");
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
    if (dt) {   // unit conversion while the fields are still in HBM (one tiny launch instead of four numpy passes on the host)
        rc = ensure(e, &e->d_dt, &e->cap_dt, n_pairs * sizeof(double));
        if (rc) return rc;
        CK(cudaMemcpyAsync(e->d_dt, dt, n_pairs * sizeof(double), cudaMemcpyHostToDevice, e->s_comp));
        const long long n = (long long)field;
        long long g = (n + 255) / 256;
        if (g > (long long)e->sm_count * 8) g = (long long)e->sm_count * 8;
        units_kernel<<<(unsigned)g, 256, 0, e->s_comp>>>(e->d_out, e->d_out + field, n, (long long)nw, res_x, res_y, e->d_dt);
        CK(cudaGetLastError());
        e->launches++;
    }
    CK(cudaMemcpyAsync(u, e->d_out + 0 * field, field * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaMemcpyAsync(v, e->d_out + 1 * field, field * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaMemcpyAsync(corr_max, e->d_out + 2 * field, field * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaMemcpyAsync(s2n, e->d_out + 3 * field, field * sizeof(float), cudaMemcpyDeviceToHost, e->s_comp));
    CK(cudaStreamSynchronize(e->s_comp));
    CK(cudaStreamSynchronize(e->s_copy));
    CK(cudaEventElapsedTime(&e->last_kernel_ms, e->ev_k0, e->ev_k1));
    return B2PIV_OK;
}

int b2piv_pairs_host(b2piv_engine* e, const void* frames, int n_frames, float signal_threshold, float* u, float* v,
                     float* corr_max, float* s2n) {
    return pairs_host_impl(e, frames, n_frames, signal_threshold, u, v, corr_max, s2n, nullptr, 0.f, 0.f);
}

int b2piv_pairs_host_units(b2piv_engine* e, const void* frames, int n_frames, float signal_threshold, float res_x, float res_y, const double* dt,
                           float* v_x, float* v_y, float* corr_max, float* s2n) {
    if (!dt) return e ? fail(e, B2PIV_ERR_ARG, "NULL pointer") : B2PIV_ERR_ARG;
    return pairs_host_impl(e, frames, n_frames, signal_threshold, v_x, v_y, corr_max, s2n, dt, res_x, res_y);
}

int b2piv_corr_planes_host(b2piv_engine* e, const void* frames, int n_frames, float signal_threshold, float* corr) {
    if (!e) return B2PIV_ERR_ARG;
    NvtxRange nvtx_range("b2piv_corr_planes_host");
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

// Accumulators are (re)allocated here; the zero-fill runs on `st`.  Work of an earlier ensemble that other streams may still have
// in flight on the accumulators (ens_add_device on a caller stream) is ordered first through ev_ens.
static int ens_begin_on(b2piv_engine* e, cudaStream_t st) {
    if (!e->planned) return fail(e, B2PIV_ERR_STATE, "b2piv_plan has not been called");
    CK(cudaSetDevice(e->device));
    const size_t nw = (size_t)e->n_rows * e->n_cols, npx = (size_t)e->wy * e->wx;
    const size_t need = nw * npx * sizeof(float);
    if (e->cap_ens < need || e->cap_ens_windows < nw) {
        if (e->ens_pending) CK(cudaEventSynchronize(e->ev_ens));
        if (e->d_ens_sum) CK(cudaFree(e->d_ens_sum));
        if (e->d_ens_cnt) CK(cudaFree(e->d_ens_cnt));
        e->d_ens_sum = e->d_ens_cnt = nullptr; e->cap_ens = 0; e->cap_ens_windows = 0;
        CK(cudaMalloc((void**)&e->d_ens_sum, need));
        CK(cudaMalloc((void**)&e->d_ens_cnt, nw * sizeof(float)));
        e->cap_ens = need;
        e->cap_ens_windows = nw;
    }
    if (e->ens_pending) CK(cudaStreamWaitEvent(st, e->ev_ens, 0));
    CK(cudaMemsetAsync(e->d_ens_sum, 0, need, st));
    CK(cudaMemsetAsync(e->d_ens_cnt, 0, nw * sizeof(float), st));
    CK(cudaEventRecord(e->ev_ens, st));
    e->ens_pending = true;
    e->ens_open = true;
    return B2PIV_OK;
}

int b2piv_ens_begin(b2piv_engine* e) {
    if (!e) return B2PIV_ERR_ARG;
    int rc = ens_begin_on(e, e->s_comp);
    if (rc) return rc;
    CK(cudaStreamSynchronize(e->s_comp));
    return B2PIV_OK;
}

int b2piv_ens_begin_device(b2piv_engine* e, void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    return ens_begin_on(e, (cudaStream_t)cuda_stream);
}

int b2piv_ens_add_device(b2piv_engine* e, const void* d_frames, long long frame_stride_bytes, int pitch_bytes, int n_frames,
                         float corr_min, float s2n_min, float signal_threshold, float* d_corr_max, float* d_s2n,
                         void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    NvtxRange nvtx_range("b2piv_ens_add_device");
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
    // the accumulators are shared state: whatever stream touched them last (zero-fill, an earlier add) comes first, and this
    // launch is what the next user - possibly on another stream - has to wait for
    if (e->ens_pending) CK(cudaStreamWaitEvent(st, e->ev_ens, 0));
    const int rc = dispatch_ens(e, p, ep, st);
    if (rc) return rc;
    CK(cudaEventRecord(e->ev_ens, st));
    e->ens_pending = true;
    return B2PIV_OK;
}

int b2piv_ens_add_host(b2piv_engine* e, const void* frames, int n_frames, float corr_min, float s2n_min,
                       float signal_threshold, float* corr_max, float* s2n) {
    if (!e) return B2PIV_ERR_ARG;
    NvtxRange nvtx_range("b2piv_ens_add_host");
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

int b2piv_ens_finish_device(b2piv_engine* e, float min_count, long long first_window, long long n_windows, float* d_u, float* d_v,
                            void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    NvtxRange nvtx_range("b2piv_ens_finish_device");
    if (!e->planned || !e->ens_open) return fail(e, B2PIV_ERR_STATE, "call b2piv_plan and b2piv_ens_begin first");
    if (!d_u || !d_v) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    const long long nw = (long long)e->n_rows * e->n_cols;
    if (first_window < 0 || n_windows < 0 || first_window + n_windows > nw) return fail(e, B2PIV_ERR_ARG, "window range outside the field");
    if (n_windows == 0) return B2PIV_OK;
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (e->ens_pending) CK(cudaStreamWaitEvent(st, e->ev_ens, 0));
    ens_finish_kernel<<<(unsigned)n_windows, 256, 0, st>>>(e->d_ens_sum + first_window * e->wy * e->wx, e->d_ens_cnt + first_window, e->wy, e->wx,
                                                          min_count, e->border_nan, e->gauss_eps, d_u, d_v);
    CK(cudaGetLastError());
    e->launches++;
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
    rc = b2piv_ens_finish_device(e, min_count, 0, (long long)nw, e->d_out, e->d_out + nw, e->s_comp);   // waits for ev_ens
    if (rc) return rc;
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

int b2piv_pairs_shifted_device(b2piv_engine* e, const void* d_frames, long long frame_stride_bytes, int pitch_bytes, int n_frames,
                               const short* d_shift, float* d_u, float* d_v, float* d_corr_max, float* d_s2n, void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    NvtxRange nvtx_range("b2piv_pairs_shifted_device");
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

// Pass 2 of the deformation scheme: `d_stack` is the interleaved float32 stack [2 n_pairs][H][W] of b2piv_deform_device, only the pairs
// (2k, 2k+1) are correlated; d_pred [n_pairs][n_windows][2] = (dv, du) is added to (v, u).  The plan must be float32.
int b2piv_pairs_interleaved_device(b2piv_engine* e, const float* d_stack, int n_pairs, const float* d_pred, float* d_u, float* d_v,
                                   float* d_corr_max, float* d_s2n, void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    NvtxRange nvtx_range("b2piv_pairs_interleaved_device");
    if (!e->planned) return fail(e, B2PIV_ERR_STATE, "b2piv_plan has not been called");
    if (e->dtype != B2PIV_F32) return fail(e, B2PIV_ERR_STATE, "the interleaved stack is float32: plan with B2PIV_F32");
    if (!d_stack || !d_u || !d_v || !d_corr_max || !d_s2n) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (n_pairs < 1) return fail(e, B2PIV_ERR_ARG, "need at least one pair");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int pitch = e->W * 4;
    const long long fstride = (long long)e->H * pitch;
    if (!rows_eligible(e, d_stack, fstride, pitch))
        return fail(e, B2PIV_ERR_UNSUPPORTED, "the deformation pass runs on the row-per-thread kernel: square 32 / 64 px windows, x stride a multiple of 4, "
                                              "frame width a multiple of 4 px");
    Params p = base_params(e, d_stack, fstride, pitch, 2 * n_pairs - 1);
    p.pair_step = 2;
    p.fshift = d_pred;
    p.u = d_u; p.v = d_v; p.cmax = d_corr_max; p.s2n = d_s2n;
    e->last_variant = 2;
    return launch_rows_f32(e, p, st);
}

// The gather as a separate push: copy this rank's result block [4][n_pairs][n_windows] into every peer's gather buffer
// [4][pairs_total][n_windows] at `pair_offset` with 16-byte stores.  Meant for a side stream, overlapped with the next step's
// compute: a kernel that writes peer memory pays for NVLink's acknowledgements when it ends (measured: +30 us per launch at N = 2,
// +64 us at N = 8, whatever thread issues the stores and however they are batched - tools/scale_probe.py), and in the epilogue of
// the PIV kernel that wait sits on the compute stream.
__global__ void __launch_bounds__(256) peer_push_kernel(const float* __restrict__ src, long long n_local /* floats per field */, b2piv::PeerOut po, long long nw) {
    const long long dst0 = po.pair0 * nw;
    const bool vec = ((n_local | dst0 | po.field) & 3) == 0;
    const long long n4 = vec ? n_local / 4 : 0;
    for (int f = 0; f < 4; ++f) {
        const float* s = src + f * n_local;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
            const float4 v = reinterpret_cast<const float4*>(s)[i];
            for (int r = 0; r < po.n; ++r) reinterpret_cast<float4*>(po.base[r] + f * po.field + dst0)[i] = v;
        }
        for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += (long long)gridDim.x * blockDim.x) {
            const float v = s[i];
            for (int r = 0; r < po.n; ++r) po.base[r][f * po.field + dst0 + i] = v;
        }
    }
}

int b2piv_peer_push(b2piv_engine* e, const float* d_local, int n_pairs, int n_peers, void* const* peer_bases, long long pairs_total,
                    long long pair_offset, void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    if (!e->planned) return fail(e, B2PIV_ERR_STATE, "b2piv_plan has not been called");
    if (!d_local || !peer_bases || n_peers < 1 || n_peers > 8) return fail(e, B2PIV_ERR_ARG, "1..8 peer buffers");
    if (n_pairs < 1 || pair_offset < 0 || pair_offset + n_pairs > pairs_total) return fail(e, B2PIV_ERR_ARG, "bad pair range");
    CK(cudaSetDevice(e->device));
    const long long nw = (long long)e->n_rows * e->n_cols;
    b2piv::PeerOut po = {};
    po.n = n_peers;
    for (int r = 0; r < n_peers; ++r) {
        if (!peer_bases[r]) return fail(e, B2PIV_ERR_ARG, "NULL peer buffer");
        po.base[r] = (float*)peer_bases[r];
    }
    po.field = pairs_total * nw;
    po.pair0 = pair_offset;
    const long long n_local = (long long)n_pairs * nw;
    long long g = (n_local / 4 + 255) / 256;
    g = g < 1 ? 1 : (g > 64 ? 64 : g);     // a small kernel on purpose: it shares the SMs with the PIV kernel of the next step
    peer_push_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)cuda_stream>>>(d_local, n_local, po, nw);
    CK(cudaGetLastError());
    e->launches++;
    return B2PIV_OK;
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
int b2piv_last_variant(const b2piv_engine* e) { return e ? e->last_variant : 0; }

}  // extern "C"

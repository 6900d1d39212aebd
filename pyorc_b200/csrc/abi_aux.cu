// abi_aux.cu - C ABI of the stages next to the PIV path: frame pre-processing, orthoprojection,
// the velocimetry mask stack, int16 packing, and the predictor of the two-pass scheme.  Kernels in preproc.cuh, project.cuh,
// mask.cuh, multipass.cuh.
#include "engine.h"
#include "preproc.cuh"
#include "project.cuh"
#include "mask.cuh"
#include "multipass.cuh"

using namespace b2piv;

extern "C" {

// ---- frame pre-processing on the device (kernels in preproc.cuh) -------------------------------------
static int pre_grid(const b2piv_engine* e, long long n) {
    long long g = (n + 4095) / 4096, cap = (long long)e->sm_count * 8;   // 16 elements per thread and iteration
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

int b2piv_pre_normalize_device(b2piv_engine* e, const void* d_frames, int dtype, int n_frames, int height, int width, int time_interval,
                               unsigned char* d_out, void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    NvtxRange nvtx_range("b2piv_pre_normalize_device");
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
    NvtxRange nvtx_range("b2piv_pre_gauss_device");
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
    // a projection of the previous plan may still be in flight on a caller stream (non-blocking streams do not synchronise with
    // the copies below): re-planning is rare (once per camera configuration), so simply wait for the device
    if (e->d_proj_off) CK(cudaDeviceSynchronize());
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
    NvtxRange nvtx_range("b2piv_project_device");
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
    NvtxRange nvtx_range("b2piv_predictor_device");
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

// Deformation pass of the two-pass scheme (multipass.cuh): validated pass-1 fields -> per-pixel predictor -> frame k+1 of every pair
// resampled; writes the interleaved float32 stack [2 (n_frames - 1)][H][W] = (frame k, warped frame k+1) and the predictor at the
// centres of the CURRENT plan's windows, float32 [n_pairs][n_rows * n_cols][2] = (dv, du).
int b2piv_deform_device(b2piv_engine* e, const void* d_frames, long long frame_stride_bytes, int pitch_bytes, int dtype, int n_frames,
                        const float* d_u1, const float* d_v1, int rows1, int cols1, int wy1, int wx1, int oy1, int ox1, float* d_stack,
                        float* d_pred, void* cuda_stream) {
    if (!e) return B2PIV_ERR_ARG;
    NvtxRange nvtx_range("b2piv_deform_device");
    if (!e->planned) return fail(e, B2PIV_ERR_STATE, "b2piv_plan (fine grid) has not been called");
    if (!d_frames || !d_u1 || !d_v1 || !d_stack || !d_pred) return fail(e, B2PIV_ERR_ARG, "NULL pointer");
    if (dtype != B2PIV_U8 && dtype != B2PIV_F32) return fail(e, B2PIV_ERR_ARG, "dtype must be B2PIV_U8 or B2PIV_F32");
    if (n_frames < 2 || rows1 < 1 || cols1 < 1 || wy1 <= oy1 || wx1 <= ox1 || oy1 < 0 || ox1 < 0) return fail(e, B2PIV_ERR_ARG, "bad coarse grid");
    if ((e->H - wy1) / (wy1 - oy1) + 1 != rows1 || (e->W - wx1) / (wx1 - ox1) + 1 != cols1)
        return fail(e, B2PIV_ERR_ARG, "coarse field shape does not match the planned frame size");
    const int esz = dtype == B2PIV_F32 ? 4 : 1;
    if (pitch_bytes % esz || frame_stride_bytes % esz) return fail(e, B2PIV_ERR_ARG, "pitch / frame stride must be a multiple of the element size");
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int n_pairs = n_frames - 1;
    const size_t n1 = (size_t)n_pairs * rows1 * cols1;
    int rc = ensure(e, &e->d_mp_ws, &e->cap_mp_ws, 2 * n1 * sizeof(double));
    if (rc) return rc;
    double* vu = e->d_mp_ws;
    double* vv = e->d_mp_ws + n1;
    mp_validate_kernel<<<mask_grid(e, (long long)n1, 128), 128, 0, st>>>(d_u1, d_v1, n_pairs, rows1, cols1, 0.1, 2.0, vu, vv);
    const MpGrid g1{rows1, cols1, wy1, wx1, wy1 - oy1, wx1 - ox1};
    const MpGrid g2{e->n_rows, e->n_cols, e->wy, e->wx, e->wy - e->oy, e->wx - e->ox};
    const long long npx = (long long)n_pairs * e->H * e->W;
    if (dtype == B2PIV_U8)
        mp_deform_kernel<unsigned char><<<mask_grid(e, npx), 256, 0, st>>>((const unsigned char*)d_frames, frame_stride_bytes, pitch_bytes, n_pairs, e->H,
                                                                          e->W, vu, vv, g1, d_stack);
    else
        mp_deform_kernel<float><<<mask_grid(e, npx), 256, 0, st>>>((const float*)d_frames, frame_stride_bytes / 4, pitch_bytes / 4, n_pairs, e->H, e->W, vu,
                                                                  vv, g1, d_stack);
    mp_predictor_float_kernel<<<mask_grid(e, (long long)n_pairs * e->n_rows * e->n_cols, 128), 128, 0, st>>>(vu, vv, n_pairs, g1, g2, d_pred);
    CK(cudaGetLastError());
    e->launches += 3;
    return B2PIV_OK;
}

// ---- fp32 FMA peak of this device, measured (bench.py's `fp32.peak_measured`; tools/fp32_peak.py) --------------------------------
// The fused PIV kernels are bound by fp32 issue, not by HBM, and MEASURED_PEAKS.json holds only the HBM and
// bf16 tensor peaks: this is the missing denominator.  Every thread runs 16 independent FFMA chains (enough to cover the
// 4-cycle dependent-issue latency with 8 warps per scheduler); 2 flop per FFMA.
}  // extern "C"
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = fmaf(x[k], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += x[k];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true: keeps the chains alive
}
extern "C" {

int b2piv_fp32_peak(b2piv_engine* e, int iters, double* tflops) {
    if (!e || !tflops) return B2PIV_ERR_ARG;
    if (iters < 1) return fail(e, B2PIV_ERR_ARG, "iters must be >= 1");
    CK(cudaSetDevice(e->device));
    int rc = ensure(e, &e->d_mask_ws, &e->cap_mask_ws, (size_t)e->sm_count * 8 * 256 * sizeof(float));
    if (rc) return rc;
    const int grid = e->sm_count * 8;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {   // first repetition warms up
        CK(cudaEventRecord(a, e->s_comp));
        fp32_peak_kernel<<<grid, 256, 0, e->s_comp>>>(e->d_mask_ws, iters, 0.999f, 1e-3f);
        CK(cudaEventRecord(b, e->s_comp));
        CK(cudaEventSynchronize(b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaEventDestroy(a));
    CK(cudaEventDestroy(b));
    CK(cudaGetLastError());
    e->launches += 5;
    *tflops = 2.0 * 16.0 * (double)iters * 256.0 * grid / ((double)best * 1e-3) / 1e12;
    return B2PIV_OK;
}

}  // extern "C"

/* b2piv.h - C ABI of the B200-native LSPIV cross-correlation engine (libb2piv.so).
 *
 * Drop-in boundary for the ONE hot path of localdevices/pyorc: Frames.get_piv -> get_ffpiv -> ffpiv.  Every entry
 * point cites the reference interface it replaces (paths relative to the pyorc tree @ be7d7c8, v0.9.9).
 * Plain pointers and sizes only; no torch / numpy types; never throws; every call returns a status code and
 * b2piv_last_error() explains a failure.  There is no CPU fallback anywhere behind this ABI.
 *
 * Conventions
 *   frames      : [n_frames][height][width] row-major, dtype B2PIV_U8 or B2PIV_F32 (what pyorc hands to
 *                 ffpiv.cross_corr as `frame_chunk.values`, pyorc/velocimetry/ffpiv.py:223,451)
 *   windows     : flattened row-major, index r*n_cols + c (reshape at ffpiv.py:469-470)
 *   u, v        : pixels / frame; u = column (x) shift, v = row (y, image-down) shift; no sign flip
 *                 (ffpiv.py:325-326,418-419 scale them by res/dt only)
 *   outputs     : float32, [n_frames-1][n_rows*n_cols]
 */
#ifndef B2PIV_H
#define B2PIV_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2piv_engine b2piv_engine;

enum { B2PIV_OK = 0, B2PIV_ERR_ARG = 1, B2PIV_ERR_CUDA = 2, B2PIV_ERR_UNSUPPORTED = 3, B2PIV_ERR_STATE = 4 };
enum { B2PIV_U8 = 0, B2PIV_F32 = 1 };

/* ABI version (major*100 + minor). */
int b2piv_version(void);

/* Engine bound to one CUDA device (one per thread and device; calls on one engine are serialised by the caller,
 * exactly like the reference's single-threaded chunk loop, ffpiv.py:348,399). */
int b2piv_create(b2piv_engine** out, int device);
void b2piv_destroy(b2piv_engine* e);
/* Message of the last failing call on `e` (or of the last failing b2piv_create when e == NULL). */
const char* b2piv_last_error(const b2piv_engine* e);

/* Numerical switches for the details that live in ffpiv rather than pyorc (see oracle/ffpiv_oracle.py):
 *   "clip_normalized" (0/1, default 0 = ffpiv), "border_nan" (0/1), "gauss_eps" (float), "copy_chunks" (H2D
 *   pipeline depth of the *_host calls; 0 = auto, about 10 MB of frames per chunk), "stage_threads" (threads that copy
 *   ordinary pageable host frames into the engine's page-locked staging ring; 0 = auto, min(8, hardware threads)),
 *   "stage_mode" (1: slices of "stage_slice_kb" KB per thread, one H2D per group of slices, "stage_groups" groups in flight in a
 *   ring small enough for the host's caches, plain stores or - "stage_nt" - non-temporal ones; 0: round 1's three large buffers), "kernel_variant" (0 auto, 1 shared-memory FFT,
 *   2 row-per-thread TMA - 32x32 / 64x64 and the polyphase 128x128 kernel, 3 direct, 4 row-per-thread TMA in
 *   padded mode for uint8 windows up to 32 px), "run_len". */
int b2piv_set_option(b2piv_engine* e, const char* name, double value);

/* Plan = frame geometry + window geometry.  Replaces the geometry half of ffpiv.cross_corr and
 * ffpiv.window.get_rect_coordinates (pyorc/api/frames.py:85-90): n_rows=(H-wy)/(wy-oy)+1, n_cols likewise.
 * Supported windows: any size 4..128 per axis - every even size pyorc can produce (frames.py:159-171).  32x32 / 64x64 /
 * 128x128 take the row-per-thread FFT kernels, other sizes up to 32 px - pyorc's 10, 20, 26 ... - their exact zero-padded
 * mode (uint8; 34..64 px: the padded mode of the 128-plane polyphase kernel), float32 frames of such sizes and the rectangular powers of two
 * a shared-memory FFT kernel, sizes with a side of 65..127 px a
 * direct-correlation kernel (exact, O(N^2) per window: a compatibility path).  search_area_size == window_size as pyorc
 * always passes (frames.py:168). */
int b2piv_plan(b2piv_engine* e, int height, int width, int win_y, int win_x, int ovl_y, int ovl_x, int dtype,
               int* n_rows, int* n_cols);

/* Per-time-step PIV on HOST buffers: the call behind `_get_uv_timestep` (ffpiv.py:446-474), i.e.
 * cross_corr + nanmax + nanmean + u_v_displacement fused.  Copies frames H2D in chunks overlapped with compute,
 * copies the four result fields back, synchronises.  signal_threshold < 0 means None (ffpiv.py:93-97). */
int b2piv_pairs_host(b2piv_engine* e, const void* frames, int n_frames, float signal_threshold, float* u, float* v,
                     float* corr_max, float* s2n);

/* The same call with the unit conversion of `_get_ffpiv_timestep` fused in (ffpiv.py:418-419: `u * res_x / dt`, `v * res_y / dt`,
 * float32): v_x, v_y come back in m / s - float32 product with the resolution, float64 division by the pair's time step `dt[k]`
 * (host array, n_frames - 1 values), one rounding to float32, i.e. numpy's arithmetic for a float32 `u` - computed while the
 * fields are still in HBM instead of in four numpy passes on the host. */
int b2piv_pairs_host_units(b2piv_engine* e, const void* frames, int n_frames, float signal_threshold, float res_x, float res_y, const double* dt,
                           float* v_x, float* v_y, float* corr_max, float* s2n);

/* Same on DEVICE-resident frames (stream-ordered, no synchronisation): rows `pitch_bytes` apart, frames
 * `frame_stride_bytes` apart.  Output pointers are device memory.  `cuda_stream` is a cudaStream_t (may be 0). */
int b2piv_pairs_device(b2piv_engine* e, const void* d_frames, long long frame_stride_bytes, int pitch_bytes,
                       int n_frames, float signal_threshold, float* d_u, float* d_v, float* d_corr_max, float* d_s2n,
                       void* cuda_stream);

/* Triage entry: the full correlation planes ffpiv.cross_corr returns (ffpiv.py:222-231), fftshifted, /N, clipped,
 * float32 [n_frames-1][n_windows][wy][wx] on the host.  Not part of the fast path. */
int b2piv_corr_planes_host(b2piv_engine* e, const void* frames, int n_frames, float signal_threshold, float* corr);

/* Ensemble-correlation mode (`_get_ffpiv_mean`, ffpiv.py:182-376).
 *   begin      : zero the per-window plane sums / valid counts (ffpiv.py:345)
 *   add        : process_frame_chunk + accumulation (ffpiv.py:200-243, :359-365); returns the masked per-pair
 *                corr_max and s2n [n_frames-1][n_windows] that aggregate_results averages on the host
 *   finish     : count filter, mean plane, peak fit (ffpiv.py:280-282, :324); min_count = count_min * n_chunks
 *   accum      : device pointers of the accumulators so ranks can reduce them (NCCL) before `finish`.
 * The accumulators are engine state shared by these calls, which may run on different streams (the engine's own for the
 * *_host calls, the caller's for the *_device calls): the engine orders every call after the previous one's work on
 * the accumulators (an event), so begin -> add -> add -> finish is safe on any mix of streams.  Work the CALLER enqueues on
 * the accumulator pointers (a collective) must be on the stream it then passes to b2piv_ens_finish_device.
 *   begin_device  : like begin, stream-ordered on `cuda_stream`, no synchronisation
 *   finish_device : count filter + mean plane + peak fit of windows [first_window, first_window + n_windows) into
 *                   d_u / d_v [n_windows] (device), stream-ordered - a rank that owns a window slice after a
 *                   reduce-scatter of the plane sums finishes just that slice */
int b2piv_ens_begin(b2piv_engine* e);
int b2piv_ens_begin_device(b2piv_engine* e, void* cuda_stream);
int b2piv_ens_add_host(b2piv_engine* e, const void* frames, int n_frames, float corr_min, float s2n_min,
                       float signal_threshold, float* corr_max, float* s2n);
int b2piv_ens_add_device(b2piv_engine* e, const void* d_frames, long long frame_stride_bytes, int pitch_bytes,
                         int n_frames, float corr_min, float s2n_min, float signal_threshold, float* d_corr_max,
                         float* d_s2n, void* cuda_stream);
int b2piv_ens_accum(b2piv_engine* e, float** d_plane_sum, float** d_count, long long* n_plane_floats,
                    long long* n_windows);
int b2piv_ens_finish_host(b2piv_engine* e, float min_count, float* u, float* v, float* count);
int b2piv_ens_finish_device(b2piv_engine* e, float min_count, long long first_window, long long n_windows, float* d_u, float* d_v,
                            void* cuda_stream);

/* Sub-pixel peak of arbitrary correlation planes [n_planes][wy][wx] (host): first-occurrence argmax + 3-point
 * Gaussian fit minus the plane centre - the standalone `ffpiv.u_v_displacement` (pyorc/velocimetry/ffpiv.py:324,471),
 * for callers that build their own planes (e.g. a mean plane).  u = column shift, v = row shift, [n_planes]. */
int b2piv_peaks_host(b2piv_engine* e, const float* corr, long long n_planes, int wy, int wx, float* u, float* v);

/* ---- Frame pre-processing on the device (the step before the path) --------------------------------
 * Contiguous [n_frames][height][width] device buffers, dtype B2PIV_U8 or B2PIV_F32, stream-ordered.
 *   normalize : pyorc Frames.normalize (pyorc/api/frames.py:279-306) -> uint8 [n][H][W]; the temporal mean uses every
 *               `time_interval`-th frame (= round(n_frames / samples), frames.py:297)
 *   time_diff : Frames.time_diff (frames.py:403-430) -> float32 [n-1][H][W]
 *   minmax    : Frames.minmax (frames.py:343-361), element-wise clamp, same dtype out
 *   gauss     : ksize1 == 0: Frames.smooth (frames.py:432-466, cv2.GaussianBlur(k2,k2,0));
 *               ksize1 > 0: Frames.edge_detect (frames.py:308-341): blur(ksize2) - blur(ksize1); float32 out */
int b2piv_pre_normalize_device(b2piv_engine* e, const void* d_frames, int dtype, int n_frames, int height, int width,
                               int time_interval, unsigned char* d_out, void* cuda_stream);
int b2piv_pre_time_diff_device(b2piv_engine* e, const void* d_frames, int dtype, int n_frames, int height, int width,
                               float thres, int absolute, float* d_out, void* cuda_stream);
int b2piv_pre_minmax_device(b2piv_engine* e, const void* d_in, int dtype, long long count, float lo, float hi, void* d_out,
                            void* cuda_stream);
int b2piv_pre_gauss_device(b2piv_engine* e, const void* d_frames, int dtype, int n_frames, int height, int width, int ksize1,
                           int ksize2, float* d_out, void* cuda_stream);

/* ---- Orthoprojection with index maps ----------------------------------------------------------------
 * Replaces pyorc.project.img_to_ortho + _group_average (pyorc/project.py:19-53, :123-157) behind project_numpy
 * (project.py:160-230).  The maps are the arrays CameraConfig.map_idx_img_ortho / map_mean_idx_img_ortho return
 * (pyorc/api/cameraconfig.py:739-860), as int64:
 *   nearest : out[idx_ortho[i]] = img[idx_img[i]], i < n_nearest (idx_ortho as POSITIONS, i.e. np.flatnonzero of the
 *             reference's boolean mask)
 *   mean    : out[uidx[g]] = float32 mean of img[src_idx[i]] over the samples with norm_idx[i] == g (accumulated in
 *             ascending i like the reference); n_samples / n_groups may be 0 (reducer != "mean")
 * Target pixels covered by neither map are 0.  `plan` merges the maps once per camera configuration; `device` projects
 * contiguous [n_frames][height][width] frames to [n_frames][out_height][out_width], stream-ordered.  out_dtype is the
 * input dtype (uint8: truncation, what Frames.project returns via output_dtypes=[da.dtype]) or B2PIV_F32 (the float
 * means img_to_ortho itself returns). */
int b2piv_project_plan(b2piv_engine* e, int height, int width, int out_height, int out_width, const long long* idx_img,
                       const long long* idx_ortho, long long n_nearest, const long long* src_idx, const long long* norm_idx,
                       long long n_samples, const long long* uidx, long long n_groups);
int b2piv_project_device(b2piv_engine* e, const void* d_frames, int dtype, int n_frames, void* d_out, int out_dtype,
                         void* cuda_stream);

/* ---- Velocimetry mask stack on the device ------------------------------------------------------------
 * Replaces the xarray passes of pyorc/api/mask.py:147-403 on the result fields while they are still in HBM.  Fields are
 * contiguous float32 [n_time][n_xy] (n_xy = ny * nx, row-major y, x) device buffers; masks are uint8 (1 = keep), either
 * [n_time][n_xy] or, where noted, [n_xy].  float32 arithmetic in numpy's operation order; stream-ordered.
 *   elementwise  op MINMAX   : p0 < sqrt(a^2 + b^2) < p1, a = v_x, b = v_y          (mask.py:147-161)
 *                op ANGLE    : |atan2(a, b) - p0| < p1                               (mask.py:163-186)
 *                op THRESHOLD: a > p0 (corr: mask.py:203-213, s2n: :215-225; b unused)
 *   time_stats   count of non-NaN samples, skipna mean and std (ddof 0) over time -> [n_xy] each (any output may be NULL)
 *   count        count > tolerance * n_time -> [n_xy]                                (mask.py:188-201)
 *   outliers     |(v - mean_t) / std_t| < tolerance per component, or (mode_and = 0) / and   (mask.py:227-252)
 *   variance     |std_t / max(mean_t, 1e30)| < tolerance -> [n_xy]                   (mask.py:254-285, its clamp included)
 *   rolling      s > tolerance * max of s over the centred window of wdw steps       (mask.py:287-303)
 *   window_nan / window_mean / window_replace: helpers.stack_window neighbourhoods (helpers.py:638-679), strides =
 *                {wdw_x_min, wdw_x_max, wdw_y_min, wdw_y_max} with the y maximum EXCLUSIVE as in the reference's range();
 *                window_nan: #valid >= tolerance * #strides (mask.py:305-337); window_mean: |v - mean| / mean <
 *                tolerance (mask.py:339-377); window_replace: NaNs of up to four fields <- window mean, in place,
 *                `iterations` times (mask.py:379-403)
 *   apply        field = mask ? field : NaN for up to four fields (ds[var].where(mask), mask.py:131-144) */
enum { B2PIV_MASK_MINMAX = 0, B2PIV_MASK_ANGLE = 1, B2PIV_MASK_THRESHOLD = 2 };
int b2piv_mask_elementwise(b2piv_engine* e, int op, const float* d_a, const float* d_b, long long count, float p0, float p1,
                           unsigned char* d_mask, void* cuda_stream);
int b2piv_time_stats(b2piv_engine* e, const float* d_field, int n_time, long long n_xy, int* d_count, float* d_mean, float* d_std,
                     void* cuda_stream);
int b2piv_mask_count(b2piv_engine* e, const float* d_vx, int n_time, long long n_xy, double tolerance, unsigned char* d_mask_xy,
                     void* cuda_stream);
int b2piv_mask_outliers(b2piv_engine* e, const float* d_vx, const float* d_vy, int n_time, long long n_xy, float tolerance,
                        int mode_and, unsigned char* d_mask, void* cuda_stream);
int b2piv_mask_variance(b2piv_engine* e, const float* d_vx, const float* d_vy, int n_time, long long n_xy, float tolerance,
                        int mode_and, unsigned char* d_mask_xy, void* cuda_stream);
int b2piv_mask_rolling(b2piv_engine* e, const float* d_vx, const float* d_vy, int n_time, long long n_xy, int wdw, float tolerance,
                       unsigned char* d_mask, void* cuda_stream);
int b2piv_mask_window_nan(b2piv_engine* e, const float* d_vx, int n_time, int ny, int nx, const int* strides, double tolerance,
                          unsigned char* d_mask, void* cuda_stream);
int b2piv_mask_window_mean(b2piv_engine* e, const float* d_vx, const float* d_vy, int n_time, int ny, int nx, const int* strides,
                           float tolerance, int mode_and, unsigned char* d_mask, void* cuda_stream);
int b2piv_window_replace(b2piv_engine* e, float* const* d_fields, int n_fields, int n_time, int ny, int nx, const int* strides,
                         int iterations, void* cuda_stream);
int b2piv_mask_apply(b2piv_engine* e, float* const* d_fields, int n_fields, int n_time, long long n_xy, const unsigned char* d_mask,
                     int mask_has_time, void* cuda_stream);

/* ---- Result packing ----------------------------------------------------------------------------------
 * The CF packing pyorc sets for v_x, v_y, corr, s2n (pyorc/const.py:80-83, Velocimetry.set_encoding,
 * pyorc/api/velocimetry.py:239-253): int16 = round_half_even(field / scale_factor), NaN -> fill_value (values beyond the
 * int16 range saturate); decode is the inverse xarray applies on reading.  rotate_uv: helpers.rotate_u_v
 * (pyorc/helpers.py:602-630) as used by Velocimetry.to_ugrid (api/velocimetry.py:284-289), float64 out. */
int b2piv_encode_int16(b2piv_engine* e, const float* d_field, long long count, float scale_factor, int fill_value, short* d_out,
                       void* cuda_stream);
int b2piv_decode_int16(b2piv_engine* e, const short* d_packed, long long count, float scale_factor, int fill_value, float* d_out,
                       void* cuda_stream);
int b2piv_rotate_uv(b2piv_engine* e, const float* d_u, const float* d_v, long long count, double theta, double* d_u2, double* d_v2,
                    void* cuda_stream);

/* ---- Two-pass PIV (BASELINE.json configs[2] "2-pass deform") --------------------------------
 * No reference counterpart: ffpiv is single pass (pyorc/velocimetry/ffpiv.py:446-474 calls cross_corr once), so the scheme is
 * defined in oracle/multipass_oracle.py.  predictor: pass-1 fields u1, v1 (float32 [n_pairs][rows1][cols1], px/frame, NaN
 * allowed) of the coarse grid (wy1, wx1, oy1, ox1) -> universal outlier detection on 3x3 neighbourhoods -> bilinear
 * interpolation at the window centres of the CURRENT plan (the fine grid) -> whole-pixel shifts (dy, dx), clamped so the
 * displaced window stays inside the frame: int16 [n_pairs][n_rows * n_cols][2].  pairs_shifted: like b2piv_pairs_device,
 * but frame k+1's window of every (pair, window) is displaced by its shift; u, v are shift + residual. */
int b2piv_predictor_device(b2piv_engine* e, const float* d_u1, const float* d_v1, int n_pairs, int rows1, int cols1, int wy1, int wx1,
                           int oy1, int ox1, short* d_shift, void* cuda_stream);
int b2piv_pairs_shifted_device(b2piv_engine* e, const void* d_frames, long long frame_stride_bytes, int pitch_bytes, int n_frames,
                               const short* d_shift, float* d_u, float* d_v, float* d_corr_max, float* d_s2n, void* cuda_stream);

/* Window DEFORMATION instead of a whole-pixel offset (the "deform" of configs[2]; definition: oracle/multipass_oracle.py,
 * two_pass_deform).  deform: the validated pass-1 field is interpolated to every pixel and frame k+1 of every pair is resampled at
 * (y + dv, x + du) (bilinear): writes the interleaved float32 stack [2 (n_frames - 1)][height][width] = (frame k, warped frame
 * k+1) and the un-rounded predictor at the centres of the CURRENT plan's windows, float32 [n_pairs][n_windows][2] = (dv, du).
 * pairs_interleaved: pass 2 on that stack (plan with B2PIV_F32; only the pairs (2k, 2k+1) are correlated), u, v = predictor +
 * residual. */
int b2piv_deform_device(b2piv_engine* e, const void* d_frames, long long frame_stride_bytes, int pitch_bytes, int dtype, int n_frames,
                        const float* d_u1, const float* d_v1, int rows1, int cols1, int wy1, int wx1, int oy1, int ox1, float* d_stack,
                        float* d_pred, void* cuda_stream);
int b2piv_pairs_interleaved_device(b2piv_engine* e, const float* d_stack, int n_pairs, const float* d_pred, float* d_u, float* d_v,
                                   float* d_corr_max, float* d_s2n, void* cuda_stream);

/* ---- Fused result gather over peer memory (multi-GPU) -------------------------------------------------------
 * pyorc processes chunks sequentially in one process (pyorc/velocimetry/ffpiv.py:399-440); sharded over the GPUs of a box,
 * every rank needs the [time] axis back together.  After this call b2piv_pairs_device stores every window's four results
 * not only into the caller's arrays but straight into each peer's gather buffer - float32 [4][pairs_total][n_rows * n_cols],
 * peer-mapped device memory (e.g. torch symmetric memory over NVLink) - at this rank's `pair_offset`: the all-gather is done
 * by P2P stores in the kernel epilogue (16 B per window and peer), no collective follows; the caller only needs a cross-rank
 * barrier before reading.  n_peers = 0 switches back to local-only results; so does a b2piv_plan that changes the field shape
 * (set it after b2piv_plan). */
int b2piv_set_peer_outputs(b2piv_engine* e, int n_peers, void* const* peer_bases, long long pairs_total, long long pair_offset);
/* The same gather as a separate PUSH for a side stream: copies the caller's contiguous result block float32 [4][n_pairs][n_rows *
 * n_cols] (what b2piv_pairs_device wrote) into every peer's buffer at `pair_offset` (16-byte P2P stores), stream-ordered.  A kernel
 * that writes peer memory waits for NVLink's acknowledgements when it ends; in the PIV kernel's epilogue that wait is on the compute
 * stream (+2 .. 4 % of a 1.6 ms step), here it overlaps the next step (pyorc_b200.parallel.PeerGather(mode="push")). */
int b2piv_peer_push(b2piv_engine* e, const float* d_local, int n_pairs, int n_peers, void* const* peer_bases, long long pairs_total,
                    long long pair_offset, void* cuda_stream);

/* Page-locked host memory so H2D copies run at full PCIe rate without staging. */
void* b2piv_host_alloc(size_t bytes);
void b2piv_host_free(void* p);

/* Introspection for bench.py: CUDA-event time (ms) spent in PIV kernels during the last *_host call, and the
 * number of kernels this engine has launched since creation. */
int b2piv_last_kernel_ms(const b2piv_engine* e, float* ms);
long long b2piv_launch_count(const b2piv_engine* e);
/* Kernel family that served the last PIV call: 1 shared-memory FFT kernel, 2 row-per-thread TMA kernels (32 / 64 / 128 px native),
 * 3 direct correlation, 4 row-per-thread kernels in padded mode; 0 before the first call. */
int b2piv_last_variant(const b2piv_engine* e);
/* Measured fp32 FMA throughput of the engine's device in TFLOP/s (independent FFMA chains, `iters` per thread and chain,
 * best of 4 runs): the denominator of the fp32 fraction bench.py reports for the fused kernels, which are bound by fp32
 * issue rather than HBM. */
int b2piv_fp32_peak(b2piv_engine* e, int iters, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* B2PIV_H */

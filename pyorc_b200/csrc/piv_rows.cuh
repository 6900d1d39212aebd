// piv_rows.cuh - "row-per-thread" fused LSPIV kernel for square 32x32 / 64x64 uint8 windows (sm_100a).
//
// Work unit = one PAIR of adjacent interrogation windows (w0, w1) followed through a run of consecutive frames.
// Per frame the two windows of THAT frame are packed as z = w0 + i*w1 and transformed by ONE complex 2-D FFT whose
// two spectra (A0, A1) are kept ("parked") in shared memory; the cross spectra against the previous frame's parked
// spectra, R_w = conj(A_w,prev) * A_w,new, are packed as G = R_0 + i*R_1 and ONE inverse complex FFT returns both
// correlation planes.  Every window of every frame is therefore forward-transformed once (the reference transforms
// it twice, as `b` of pair k and `a` of pair k+1): 1.0 complex FFT per window and pair instead of 1.5.
//
// A group of W threads owns the unit: thread t holds row t (then column c(t)) of the plane in REGISTERS - W complex
// values - and runs whole W-point FFTs there with compile-time twiddles (zero index arithmetic); the only shared
// memory traffic is the source tile (TMA, 64B/32B swizzle), two transposes per frame and the parked half-spectra.
//   TMA tile -> rows in regs -> row FFT -> transpose -> column FFT -> [shuffle with the -kx partner lane]
//   -> cross spectra + park -> inverse column FFT -> transpose -> inverse row FFT -> clip/max/sum -> peak fit.
//
// Replaces: ffpiv.cross_corr + nanmax/nanmean + ffpiv.u_v_displacement (pyorc/velocimetry/ffpiv.py:446-474).
// The phases are __host__ __device__ and barrier-delimited like piv_core.cuh so tests/emul can run them on the CPU.
#pragma once
#include "piv_core.cuh"
#include <cstddef>
#include "twiddle128.h"

namespace b2piv {

// ------------------------------------------------------------------------------------------------------------
// W-point in-register FFT, natural order in -> natural order out, compile-time twiddles.
// ------------------------------------------------------------------------------------------------------------
template <int W, int INV>
B2_HD void fft_reg(float2* v) {
    constexpr int A = Factor<W>::R1, B = Factor<W>::R2;  // n = B*n1 + n2, k = k1 + A*k2
#pragma unroll
    for (int n2 = 0; n2 < B; ++n2) {
        float2 t[A];
#pragma unroll
        for (int i = 0; i < A; ++i) t[i] = v[B * i + n2];
        RegDFT<A, INV>::run(t);
#pragma unroll
        for (int k1 = 0; k1 < A; ++k1) {
            const int e = (n2 * k1) * (128 / W);
            if (n2 * k1 == 0) {
                v[B * k1 + n2] = t[k1];
            } else {
                v[B * k1 + n2] = ctw<INV>(t[k1], cos128(e), sin128(e));
            }
        }
    }
#pragma unroll
    for (int k1 = 0; k1 < A; ++k1) RegDFT<B, INV>::run(v + B * k1);
    float2 o[W];
#pragma unroll
    for (int k1 = 0; k1 < A; ++k1)
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) o[k1 + A * k2] = v[B * k1 + k2];
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = o[i];
}

// ------------------------------------------------------------------------------------------------------------
// Configuration / shared memory / parameters
// ------------------------------------------------------------------------------------------------------------
template <int W_, int BP_ = 34>
struct RCfg {
    static constexpr int W = W_;
    static constexpr int NT = W;              // one thread per row
    static constexpr int NWARP = W / 32;
    static constexpr int P = W + 1;           // transpose-plane pitch in float2 (odd: conflict-free both ways)
    static constexpr int HS = W / 2 + 1;      // ky = 0 .. W/2 computed directly, the rest by Hermitian symmetry
    static constexpr int NPX = W * W;
    static constexpr int TILE = 2 * W * W;    // bytes: two windows of one frame, uint8 (16-byte aligned window starts)
    // window starts that are only 4-byte aligned (e.g. 32x32 at 75 % overlap: stride 8): the TMA box starts at the
    // 16-byte boundary below and is 16 bytes wider, rows are read at a per-window byte offset, no swizzle
    static constexpr int WB = W + 16;
    static constexpr int TILE_U = 2 * W * WB;
    // float32 frames: [W rows][128 B] boxes (32 floats, SWIZZLE_128B), W/32 boxes per window; a 64x64 window (16 KB)
    // fills the transpose buffer, so its two windows arrive one after the other, a 32x32 pair (8 KB) together
    static constexpr int FBOX = W * 128;                 // bytes per box
    static constexpr int FWIN = (W / 32) * FBOX;         // bytes per window
    static constexpr int F_PHASES = (W == 64) ? 2 : 1;   // TMA round trips per frame
    // float32 frames in padded mode (window sides <= W/2, any x stride): one un-swizzled box of PFW floats x ny rows per window
    // from the 16-byte boundary below the window start (up to 3 floats of lead-in), window 1 at the fixed offset PFWIN
    static constexpr int PFW = W / 2 + 4;                // floats per box row: 144 B (W = 64) / 80 B (W = 32), multiples of 16 B
    static constexpr int PFWIN = (W / 2) * PFW * 4;      // bytes; 4608 / 1280: multiples of 128 (TMA destination alignment)
    // Transposes move 32x32 blocks: block b (pitch BP float2) is READ by warp b.  A 64x64 transpose first moves the
    // two off-diagonal blocks (one CTA barrier), then the diagonal ones (warp-synchronous), reusing the same two
    // blocks - half a plane - so a group needs ~52 KB of shared memory and FOUR groups (8 warps, two per
    // scheduler) fit on an SM.
    // block pitch in float2.  34: rows start 16-byte aligned (272 B apart) and a thread stores its 32 values as 16 STS.128 (banks
    // 4*lane + 4*j mod 32: conflict-free); 33 (odd -> conflict-free 8-byte accesses both ways) stores them as 32 STS.64.
    // A/B on one B200 (tools/ab_lib.py).  Round 1: STS.128 gained 1.7 % at 32x32 and lost 1 % at 64x64 / 128x128.  Round 2, with
    // packed fp32 arithmetic and the parked spectra out of shared memory: +3.5 % at 64x64 (106.2 -> 110.3 M windows/s), -0.7 % for
    // the 128x128 kernel, which keeps 33 (R6 in piv_rows128.cuh).
    static constexpr int BP = BP_;
    static constexpr int XBLK = 32 * BP;          // float2 per block
};

template <class R>
struct RSmem {
    // transpose plane; the TMA tile ([window][row][W bytes], 64B/32B swizzled) is ALIASED onto its start: the tile is
    // consumed into registers before the first transpose of a frame and refilled after the last one
    alignas(1024) float2 X[R::NWARP][R::XBLK];
    float nb[2][3][R::W];                        // the three plane rows around each peak (Gaussian fit)
    unsigned red[R::NWARP][8];                   // block reductions (integer moments, float bits)
    unsigned long long redk[R::NWARP][2];
    unsigned rowc[R::NWARP][2];                  // per warp and window: first row (reference order) holding the warp's maximum
    int peak_j[2];                               // column of the peak, found by the thread that owns the peak's row
    unsigned long long mbar;                     // TMA completion barrier
    // previous frame: scaled spectra (A0, A1) at (ky <= W/2, own column), one 16-byte access.  LAST member: the kernel variant
    // that parks the spectra in Tensor Memory (rows_kernel.cuh, piv_rows_tm_kernel) packs its groups at tm_group_stride() and never
    // touches this array
    float4 park[R::HS][R::NT];
    B2_HD unsigned char* tile() { return reinterpret_cast<unsigned char*>(X); }
};
// distance between the groups of a CTA when the parked spectra live in Tensor Memory (everything but `park`)
template <class R>
constexpr size_t tm_group_stride() { return (offsetof(RSmem<R>, park) + 1023) / 1024 * 1024; }
static_assert(sizeof(float2) * RCfg<64>::NWARP * RCfg<64>::XBLK >= RCfg<64>::TILE_U, "tile must fit in the transpose blocks");
static_assert(sizeof(float2) * RCfg<64>::NWARP * RCfg<64>::XBLK >= RCfg<64>::FWIN, "float tile must fit in the transpose blocks");
static_assert(sizeof(float2) * RCfg<32>::NWARP * RCfg<32>::XBLK >= 2 * RCfg<32>::FWIN, "float tiles must fit in the transpose blocks");
static_assert(sizeof(float2) * RCfg<32>::NWARP * RCfg<32>::XBLK >= RCfg<32>::TILE_U, "tile must fit in the transpose blocks");
static_assert(sizeof(float2) * RCfg<64>::NWARP * RCfg<64>::XBLK >= 2 * RCfg<64>::PFWIN && sizeof(float2) * RCfg<32>::NWARP * RCfg<32>::XBLK >= 2 * RCfg<32>::PFWIN,
              "padded float tiles must fit in the transpose blocks");
static_assert(RCfg<64>::PFWIN % 128 == 0 && RCfg<32>::PFWIN % 128 == 0 && (RCfg<64>::PFW * 4) % 16 == 0 && (RCfg<32>::PFW * 4) % 16 == 0, "TMA box alignment");

// displaced-window variant (second pass of the two-pass scheme): dedicated tile buffers so that both tiles of the next frame
// can be prefetched while the transposes use X
template <class R>
struct RShiftSmem {
    RSmem<R> base;
    alignas(128) unsigned char tile_u[R::TILE_U];   // undisplaced windows of the frame (parked as `a` of the next pair)
    alignas(128) unsigned char tile_d[R::TILE_U];   // displaced windows of the frame (`b` of this pair)
    unsigned long long mbar_d;
};

struct RParams {
    const unsigned char* frames;   // only used by the host emulator (device reads through the tensor map)
    long long frame_stride;
    int pitch;
    int n_rows, n_cols, sy, sx;
    int n_pairs;                   // frame pairs in this launch (frames = n_pairs + 1)
    int run_len;                   // frame pairs per work unit
    int n_units;                   // = n_wpairs * ceil(n_pairs / run_len)
    // optional explicit unit list [n_units][3] = (window pair, first pair, one past the last pair), window pair < 0: no unit.
    // Kernels whose groups run independently get an even 1-D partition of the (window pair, frame pair) space this way
    // (engine.h, build_unit_table): every group the same number of frame pairs instead of whole waves of equal units.
    const int* unit_table;
    int clip_norm, border_nan;
    float gauss_eps;
    const unsigned char* keep;
    const short* shift;            // displaced second pass: [n_pairs][n_windows][2] = (dy, dx) of frame k+1's window, added to (v, u)
    const float* fshift;           // deformation pass: [n_out_pairs][n_windows][2] = (dv, du) float predictor added to (v, u)
    int pair_step;                 // 2: interleaved stack (a_0, b_0, a_1, b_1 ...), units = the pairs (2k, 2k+1), result index k; else 0 / 1
    float *u, *v, *cmax, *s2n;
    PeerOut peer;                  // optional fused gather over peer memory (piv_core.cuh)
    float* planes;
    // ensemble mode (piv_rows_kernel<..., ENS = true>): thresholds and the HBM accumulators [n_windows][W][W] / [n_windows]
    float corr_min, s2n_min;
    float* ens_sum;
    float* ens_count;
    // padded mode (PAD = true): the true window ny x nx is at most half the W x W FFT plane (see rows_p1_pad)
    int ny, nx;
    float pad_scale;               // 1 / (W*W * ny*nx)
    float2 pad_ty[33];             // Ty(ky) = 1 + exp(-2 pi i ky ny / W), ky = 0 .. W/2
    float2 pad_tx[64];             // Tx(kx) = 1 + exp(-2 pi i kx nx / W)
    unsigned pad_mask[16];         // byte mask of source word k: bytes 4k .. 4k+3 that lie inside the window (x < nx)
    float pad_cm[64];              // 1.0f for x < nx, else 0.0f (statically indexed -> constant-bank operands)
    int height;                    // host emulator only (rows below the frame read as 0, like the TMA fill)
};

// Line (row before / column after the transpose) owned by a thread, chosen so that the -kx partner sits in the SAME warp at lane^16
// (kx = 0 and W/2 are their own partners): warp w, lane l<16 -> column 16w+l ; lane 16+l -> (W-(16w+l)) % W,
// except (w=0,l=0) -> W/2.
template <int W>
B2_HD constexpr int column_of(int tid) {
    const int w = tid >> 5, l = tid & 31;
    if (l < 16) return 16 * w + l;
    const int m = 16 * w + (l - 16);
    return m == 0 ? W / 2 : (W - m) % W;
}
template <int W>
B2_HD int partner_lane_of(int tid) {
    const int c = column_of<W>(tid);
    const int l = tid & 31;
    return (c == 0 || c == W / 2) ? l : (l ^ 16);
}

// swizzled byte offset of 16-byte chunk j of row r of window w in the tile (TMA SWIZZLE_64B / SWIZZLE_32B)
template <int W>
B2_HD int tile_chunk_offset(int w, int r, int j) {
    const int x = (W == 64) ? ((r >> 1) & 3) : ((r >> 2) & 1);
    return w * W * W + r * W + ((j ^ x) << 4);
}

// Per-thread register state that lives across the phases of one frame (and, for alpha/dead, across frames).
template <class R>
struct RRegs {
    float2 v[R::W];
    unsigned px[2][R::W / 4];     // packed source row of window 0 / 1
    float half_alpha_prev[2];     // 0.5 / std of the previous frame's windows (0: dead)
    float half_alpha_new[2];
    float mean_new[2];
    float dc_fix[2];              // N * (mean - quantised mean) of the magic-number centring (rows_p2_pre<.., true>), removed from Z(0, 0)
    float rowmax[2], rowsum[2];
    bool dead[2];                 // window w has zero variance in the previous or the current frame
    int pi[2], pj[2];
    float cmaxv[2], sumv[2];
    float2 r0, r1;                // host emulator only: cross spectra of the current ky step
    float2 tx;                    // padded mode: Tx(own column)
};

struct RUnit {
    int w[2];
    int valid1;
    int f0, f1;   // frames f0..f1 inclusive; pairs f0..f1-1
    int x0[2], y0[2];
};

B2_HD RUnit decode_unit(const RParams& p, int unit) {
    const int nw = p.n_rows * p.n_cols;
    const int n_wp = (nw + 1) / 2;
    RUnit u;
    int wp, chunk = 0;
    if (p.unit_table) {
        wp = p.unit_table[3 * unit];
        u.f0 = p.unit_table[3 * unit + 1];
        u.f1 = p.unit_table[3 * unit + 2];
        if (wp < 0) { wp = 0; u.f0 = 0; u.f1 = -1; }          // no unit: f1 - f0 + 1 = 0 frames
    } else {
        wp = unit % n_wp; chunk = unit / n_wp;
        if (p.pair_step == 2) { u.f0 = 2 * chunk; u.f1 = u.f0 + 1; }
        else {
            u.f0 = chunk * p.run_len;
            u.f1 = u.f0 + p.run_len < p.n_pairs ? u.f0 + p.run_len : p.n_pairs;
        }
    }
    u.w[0] = 2 * wp;
    u.valid1 = (2 * wp + 1 < nw);
    u.w[1] = u.valid1 ? 2 * wp + 1 : 2 * wp;
    for (int k = 0; k < 2; ++k) {
        u.y0[k] = (u.w[k] / p.n_cols) * p.sy;
        u.x0[k] = (u.w[k] % p.n_cols) * p.sx;
    }
    return u;
}

// ------------------------------------------------------------------------------------------------------------
// P1: rows from the tile into registers + integer moments (exact) -> red[warp][0..3] = S0, Q0, S1, Q1
// ------------------------------------------------------------------------------------------------------------
B2_HD unsigned dp4a_u(unsigned a, unsigned b, unsigned c) {
#ifdef __CUDA_ARCH__
    return __dp4a(a, b, c);
#else
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
    return c;
#endif
}

B2_HD float bits_f32(unsigned u) { union { float f; unsigned u; } a; a.u = u; return a.f; }

// byte b of a packed word as float
B2_HD float byte_to_float(unsigned word, int b) {
    // one I2F with a byte selector: a single issue slot on the (otherwise idle) conversion pipe, cheaper here than the
    // PRMT + FADD "magic number" sequence which costs two slots of the busy FMA/ALU pipes
    return (float)((word >> (8 * b)) & 0xffu);
}

template <class R, bool ALIGNED = true>
B2_HD void rows_p1(RSmem<R>& s, RRegs<R>& r, int tid, int xoff0 = 0, int xoff1 = 0, const unsigned char* tile_override = nullptr) {
    constexpr int W = R::W;
    const unsigned char* tile_base = tile_override ? tile_override : s.tile();
    unsigned S[2] = {0, 0}, Q[2] = {0, 0};
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        if (ALIGNED) {
#pragma unroll
            for (int j = 0; j < W / 16; ++j) {
                const uint4 q = *reinterpret_cast<const uint4*>(tile_base + tile_chunk_offset<W>(w, column_of<W>(tid), j));
                r.px[w][4 * j + 0] = q.x; r.px[w][4 * j + 1] = q.y; r.px[w][4 * j + 2] = q.z; r.px[w][4 * j + 3] = q.w;
            }
        } else {
            const unsigned char* row = tile_base + (w * W + column_of<W>(tid)) * R::WB + (w == 0 ? xoff0 : xoff1);
#pragma unroll
            for (int k = 0; k < W / 4; ++k) r.px[w][k] = *reinterpret_cast<const unsigned*>(row + 4 * k);
        }
        // two accumulators per moment: the dp4a chains are the critical path of this phase (ncu: 27 % `wait` stalls)
        unsigned S2 = 0, Q2 = 0;
#pragma unroll
        for (int k = 0; k < W / 4; k += 2) {
            S[w] = dp4a_u(r.px[w][k], 0x01010101u, S[w]);
            Q[w] = dp4a_u(r.px[w][k], r.px[w][k], Q[w]);
            S2 = dp4a_u(r.px[w][k + 1], 0x01010101u, S2);
            Q2 = dp4a_u(r.px[w][k + 1], r.px[w][k + 1], Q2);
        }
        S[w] += S2; Q[w] += Q2;
    }
    unsigned vals[4] = {S[0], Q[0], S[1], Q[1]};
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int k = 0; k < 4; ++k) vals[k] = __reduce_add_sync(0xffffffffu, vals[k]);   // one REDUX each (exact: sums stay below 2^32)
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) s.red[tid >> 5][k] = vals[k];
    }
#else
    for (int k = 0; k < 4; ++k) s.red[tid >> 5][k] += vals[k];
#endif
}

// ------------------------------------------------------------------------------------------------------------
// Padded mode: window sizes that are not a power of two (pyorc: 10, 20, 26, 30 ... from the camera configuration).
// The ny x nx window of frame k is embedded ZERO-PADDED in the W x W plane (2*max(ny, nx) <= W).  The circular
// correlation of period (ny, nx) the reference computes equals the linear correlation of that plane against the window
// of frame k+1 TILED 2 x 2 - and the tiled plane is the zero-padded one convolved with four deltas, so its spectrum is the
// zero-padded spectrum times T(ky, kx) = (1 + w^(ny ky)) (1 + w^(nx kx)), w = exp(-2 pi i / W).  One forward transform
// per window and frame therefore still serves both roles (as `a` of the next pair, and times T as `b` of this one), and
// the reference's plane is the central ny x nx block of the fftshifted W x W result.  Exact (no wrap-around) like
// piv_core.cuh's phase_embed, at the cost of a native W x W window.
// P1: the TMA box starts at the 16-byte boundary below the window (any x stride); rows are funnel-shifted to the byte
// offset and masked to nx bytes, rows >= ny are zero, so the integer moments cover exactly the window.
// ------------------------------------------------------------------------------------------------------------
B2_HD unsigned funnel_r(unsigned lo, unsigned hi, int sh) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
template <class R>
B2_HD void rows_p1_pad(RSmem<R>& s, RRegs<R>& r, int tid, const RParams& p, int xoff0, int xoff1) {
    constexpr int W = R::W;
    const int row = column_of<W>(tid);
    const unsigned rowmask = row < p.ny ? 0xffffffffu : 0u;
    unsigned S[2] = {0, 0}, Q[2] = {0, 0};
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        const int xoff = w == 0 ? xoff0 : xoff1;
        const unsigned char* base = s.tile() + (w * W + row) * R::WB + (xoff & ~3);
        const int sh = (xoff & 3) * 8;
        unsigned wd[W / 4 + 1];
#pragma unroll
        for (int k = 0; k <= W / 4; ++k) wd[k] = *reinterpret_cast<const unsigned*>(base + 4 * k);
#pragma unroll
        for (int k = 0; k < W / 4; ++k) {
            r.px[w][k] = funnel_r(wd[k], wd[k + 1], sh) & p.pad_mask[k] & rowmask;
            S[w] = dp4a_u(r.px[w][k], 0x01010101u, S[w]);
            Q[w] = dp4a_u(r.px[w][k], r.px[w][k], Q[w]);
        }
    }
    unsigned vals[4] = {S[0], Q[0], S[1], Q[1]};
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int k = 0; k < 4; ++k) vals[k] = __reduce_add_sync(0xffffffffu, vals[k]);   // one REDUX each (exact: sums stay below 2^32)
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) s.red[tid >> 5][k] = vals[k];
    }
#else
    for (int k = 0; k < 4; ++k) s.red[tid >> 5][k] += vals[k];
#endif
}
// P2 (padded): moments over the ny*nx window pixels; pixels outside the window stay exactly 0 after centring
template <class R>
B2_HD void rows_p2_pre_pad(RSmem<R>& s, RRegs<R>& r, int tid, const RParams& p) {
    constexpr int W = R::W;
    const unsigned long long npx = (unsigned long long)(p.ny * p.nx);
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        unsigned long long S = 0, Q = 0;
#pragma unroll
        for (int k = 0; k < R::NWARP; ++k) { S += s.red[k][2 * w]; Q += s.red[k][2 * w + 1]; }
        const unsigned long long m2 = npx * Q - S * S;
        r.mean_new[w] = (float)S / (float)npx;
        r.half_alpha_new[w] = m2 ? 0.5f * (float)npx * (1.0f / sqrtf((float)m2)) : 0.f;
    }
    // masked bytes are 0, so a pixel outside the window must only NOT get the mean subtracted: rows >= ny use mean 0,
    // columns >= nx a 0/1 factor from the constant bank - one FFMA per value, like the FADD of the native path
    const bool rowok = column_of<W>(tid) < p.ny;
    r.dc_fix[0] = r.dc_fix[1] = 0.f;
#ifdef __CUDA_ARCH__
    if (!p.clip_norm) {
        // magic-number conversion (rows_p2_pre): byte -> 32768 + b by a byte permute, exact subtraction of 32768, then
        // byte - cm * mq with the mean rounded to 1/256; the DC offset n_px * (mean - mq) is removed from Z(0, 0) in the cross phase
        const float c0 = __fadd_rn(32768.0f, r.mean_new[0]), c1 = __fadd_rn(32768.0f, r.mean_new[1]);
        const float mq0 = c0 - 32768.0f, mq1 = c1 - 32768.0f;
        r.dc_fix[0] = (r.mean_new[0] - mq0) * (float)npx;
        r.dc_fix[1] = (r.mean_new[1] - mq1) * (float)npx;
        const float2 nm = rowok ? make_float2(-mq0, -mq1) : make_float2(0.f, 0.f);
        const float2 base = make_float2(32768.0f, 32768.0f);
#pragma unroll
        for (int k = 0; k < W / 4; ++k) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float2 mg = make_float2(__uint_as_float(__byte_perm(r.px[0][k], 0x47000000u, 0x7404u | (b << 4))),
                                              __uint_as_float(__byte_perm(r.px[1][k], 0x47000000u, 0x7404u | (b << 4))));
                r.v[4 * k + b] = pk_fma(make_float2(p.pad_cm[4 * k + b], p.pad_cm[4 * k + b]), nm, pk_sub(mg, base));
            }
        }
        r.tx = p.pad_tx[column_of<W>(tid)];
        return;
    }
#endif
    const float nm0 = rowok ? -r.mean_new[0] : 0.f, nm1 = rowok ? -r.mean_new[1] : 0.f;
#pragma unroll
    for (int k = 0; k < W / 4; ++k) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float a0 = fmaf(p.pad_cm[4 * k + b], nm0, byte_to_float(r.px[0][k], b));
            float a1 = fmaf(p.pad_cm[4 * k + b], nm1, byte_to_float(r.px[1][k], b));
            if (p.clip_norm) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
            r.v[4 * k + b] = make_float2(a0, a1);
        }
    }
    r.tx = p.pad_tx[column_of<W>(tid)];
}

// ------------------------------------------------------------------------------------------------------------
// float32 frames.  F1(w): row of window w from its tile (at byte offset `toff`) into component w of r.v + row sum
// (-> red[warp][2w]);  F2(w): mean from the block sum, centre, centred second moment (-> red[warp][2w+1]);
// F3: 0.5/std of both windows, optional clip.  Two-pass moments like numpy's float path.
// ------------------------------------------------------------------------------------------------------------
B2_HD void red_put_f32(unsigned* cell, float v, int tid) {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) *cell = __float_as_uint(v);
#else
    union { float f; unsigned u; } a;
    a.u = *cell; a.f += v; *cell = a.u;
#endif
}
template <class R>
B2_HD void rows_f1(RSmem<R>& s, RRegs<R>& r, int tid, int w, int toff) {
    constexpr int W = R::W;
    const int row = column_of<W>(tid);
    float sum = 0.f;
#pragma unroll
    for (int h = 0; h < W / 32; ++h) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 q = *reinterpret_cast<const float4*>(s.tile() + toff + h * R::FBOX + row * 128 + ((j ^ (row & 7)) << 4));
            const float vals[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int x = 32 * h + 4 * j + k;
                if (w == 0) r.v[x].x = vals[k]; else r.v[x].y = vals[k];
                sum += vals[k];
            }
        }
    }
    red_put_f32(&s.red[tid >> 5][2 * w], sum, tid);
}
template <class R>
B2_HD void rows_f2(RSmem<R>& s, RRegs<R>& r, int tid, int w) {
    constexpr int W = R::W;
    float S = 0.f;
#pragma unroll
    for (int k = 0; k < R::NWARP; ++k) S += bits_f32(s.red[k][2 * w]);
    const float mean = S * (1.0f / (float)R::NPX);
    float q = 0.f;
#pragma unroll
    for (int x = 0; x < W; ++x) {
        if (w == 0) { r.v[x].x -= mean; q = fmaf(r.v[x].x, r.v[x].x, q); }
        else        { r.v[x].y -= mean; q = fmaf(r.v[x].y, r.v[x].y, q); }
    }
    red_put_f32(&s.red[tid >> 5][2 * w + 1], q, tid);
}
template <class R>
B2_HD void rows_f3(RSmem<R>& s, RRegs<R>& r, int tid, int clip_norm) {
    constexpr int W = R::W;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        float Q = 0.f;
#pragma unroll
        for (int k = 0; k < R::NWARP; ++k) Q += bits_f32(s.red[k][2 * w + 1]);
        r.half_alpha_new[w] = Q > 0.f ? 0.5f * sqrtf((float)R::NPX) * (1.0f / sqrtf(Q)) : 0.f;   // 0.5 / sqrt(Q/N)
    }
    if (clip_norm) {
#pragma unroll
        for (int x = 0; x < W; ++x) r.v[x] = make_float2(fmaxf(r.v[x].x, 0.f), fmaxf(r.v[x].y, 0.f));
    }
}

// ------------------------------------------------------------------------------------------------------------
// float32 frames in padded mode (round 2): what pyorc's own example recipe produces - `normalize -> edge_detect -> minmax ->
// get_piv(window_size=25)` hands float32 frames and a 26 x 26 window to the engine (examples/ngwerere/ngwerere.yml,
// pyorc/api/frames.py:430,456-466).  Same embedding as the uint8 padded mode: the thread that owns row < ny reads its nx floats
// at the window's float offset `xoff` (0 .. 3) inside the box, everything else of the W x W plane is exactly 0; two-pass moments
// over the ny * nx window pixels like rows_f1 / f2 / f3.  Downstream (spectrum factor of the 2 x 2 tiling, reductions restricted
// to the window's lags) is the uint8 padded path unchanged.
// ------------------------------------------------------------------------------------------------------------
template <class R>
B2_HD void rows_f1_pad(RSmem<R>& s, RRegs<R>& r, int tid, const RParams& p, int w, int xoff) {
    constexpr int W = R::W;
    const int row = column_of<W>(tid);
    const bool rowok = row < p.ny;
    const float* src = reinterpret_cast<const float*>(s.tile() + w * R::PFWIN) + (rowok ? row : 0) * R::PFW + xoff;
    float sum = 0.f;
#pragma unroll
    for (int x = 0; x < W / 2; ++x) {
        const float val = (rowok && x < p.nx) ? src[x] : 0.f;
        if (w == 0) r.v[x].x = val; else r.v[x].y = val;
        sum += val;
    }
#pragma unroll
    for (int x = W / 2; x < W; ++x) {
        if (w == 0) r.v[x].x = 0.f; else r.v[x].y = 0.f;
    }
    red_put_f32(&s.red[tid >> 5][2 * w], sum, tid);
}
template <class R>
B2_HD void rows_f2_pad(RSmem<R>& s, RRegs<R>& r, int tid, const RParams& p, int w) {
    constexpr int W = R::W;
    float S = 0.f;
#pragma unroll
    for (int k = 0; k < R::NWARP; ++k) S += bits_f32(s.red[k][2 * w]);
    const float mean = S / (float)(p.ny * p.nx);
    const bool rowok = column_of<W>(tid) < p.ny;
    float q = 0.f;
#pragma unroll
    for (int x = 0; x < W / 2; ++x) {
        const float m = (rowok && x < p.nx) ? mean : 0.f;   // pixels outside the window stay exactly 0
        if (w == 0) { r.v[x].x -= m; q = fmaf(r.v[x].x, r.v[x].x, q); }
        else        { r.v[x].y -= m; q = fmaf(r.v[x].y, r.v[x].y, q); }
    }
    red_put_f32(&s.red[tid >> 5][2 * w + 1], q, tid);
}
template <class R>
B2_HD void rows_f3_pad(RSmem<R>& s, RRegs<R>& r, int tid, const RParams& p) {
    constexpr int W = R::W;
    const float npx = (float)(p.ny * p.nx);
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        float Q = 0.f;
#pragma unroll
        for (int k = 0; k < R::NWARP; ++k) Q += bits_f32(s.red[k][2 * w + 1]);
        r.half_alpha_new[w] = Q > 0.f ? 0.5f * sqrtf(npx) * (1.0f / sqrtf(Q)) : 0.f;   // 0.5 / sqrt(Q / (ny nx))
    }
    if (p.clip_norm) {
#pragma unroll
        for (int x = 0; x < W / 2; ++x) r.v[x] = make_float2(fmaxf(r.v[x].x, 0.f), fmaxf(r.v[x].y, 0.f));
    }
    r.dc_fix[0] = r.dc_fix[1] = 0.f;
    r.tx = p.pad_tx[column_of<W>(tid)];
}

// ------------------------------------------------------------------------------------------------------------
// P2: statistics -> mean, 0.5/std ; convert + centre (+clip) ; forward row FFT ; transposed store into X
// ------------------------------------------------------------------------------------------------------------
// MAGIC (device, un-clipped normalisation only): a byte becomes a float by ONE byte permute into the mantissa of 32768.0f
// (0x47000000 | b << 8 = 32768 + b exactly) - an ALU instruction instead of an I2F conversion, which goes through the MIO queue
// like shared memory and shuffles (A/B on one B200: +3.6 %).  The packed subtraction of c = fl(32768 + mean) then centres with the
// mean rounded to 1/256: (32768 + b) - c = b - mq EXACTLY.  The difference delta = mean - mq is known exactly (|delta| <= 2^-9, a
// multiple of 1/N) and only shifts the DC bin: FFT2(b - mq) = FFT2(b - mean) + N delta at k = 0, so the cross phase subtracts
// N (delta0 + i delta1) - a small whole number - from Z(0, 0) (RRegs::dc_fix) and everything downstream is unchanged.
template <class R, bool MAGIC = false>
B2_HD void rows_p2_pre(RSmem<R>& s, RRegs<R>& r, int tid, int clip_norm) {
    constexpr int W = R::W;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        unsigned long long S = 0, Q = 0;
#pragma unroll
        for (int k = 0; k < R::NWARP; ++k) { S += s.red[k][2 * w]; Q += s.red[k][2 * w + 1]; }
        const unsigned long long m2 = (unsigned long long)R::NPX * Q - S * S;   // N^2 * variance, exact
        r.mean_new[w] = (float)S * (1.0f / (float)R::NPX);
        // 0.5 / std = 0.5 * N / sqrt(N*Q - S^2)
        r.half_alpha_new[w] = m2 ? 0.5f * (float)R::NPX * (1.0f / sqrtf((float)m2)) : 0.f;
        r.dc_fix[w] = 0.f;
    }
#ifdef __CUDA_ARCH__
    if (MAGIC && !clip_norm) {
        const float c0 = __fadd_rn(32768.0f, r.mean_new[0]), c1 = __fadd_rn(32768.0f, r.mean_new[1]);
        r.dc_fix[0] = (r.mean_new[0] - (c0 - 32768.0f)) * (float)R::NPX;     // exact: every term is a multiple of 1/N below 2^16
        r.dc_fix[1] = (r.mean_new[1] - (c1 - 32768.0f)) * (float)R::NPX;
        const float2 c = make_float2(c0, c1);
#pragma unroll
        for (int k = 0; k < W / 4; ++k) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float2 mg = make_float2(__uint_as_float(__byte_perm(r.px[0][k], 0x47000000u, 0x7404u | (b << 4))),
                                              __uint_as_float(__byte_perm(r.px[1][k], 0x47000000u, 0x7404u | (b << 4))));
                r.v[4 * k + b] = pk_sub(mg, c);
            }
        }
        return;
    }
#endif
#pragma unroll
    for (int k = 0; k < W / 4; ++k) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float2 a = pk_sub(make_float2(byte_to_float(r.px[0][k], b), byte_to_float(r.px[1][k], b)), make_float2(r.mean_new[0], r.mean_new[1]));
            if (clip_norm) a = make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f));
            r.v[4 * k + b] = a;
        }
    }
}
// ---- forward transpose (thread = row  ->  thread = column), 32x32 blocks, see RCfg ----------------------------
// WQ = warp of this thread (compile time so that every register index stays static).
// "other": the values of my row that belong to columns owned by the OTHER warp go to block (1-WQ); after a CTA barrier I
// read, from block WQ, the other warp's rows at my column.  "own": same inside the warp (warp-synchronous).
// Thread t owns LINE sigma(t) = column_of(t) in both orientations (row sigma(t) before the transpose, column sigma(t)
// after it), so the transpose is one symmetric, in-place operation T with static register indices: thread t sends
// v[sigma(u)] to thread u and stores what u sends into the same register, v[sigma(u)] = M(sigma(u), sigma(t)) - natural
// order again.  T is its own inverse, so the forward (rows -> columns) and inverse (columns -> rows) transposes are the
// same code.  SET selects the 32 partner threads of warp SET (compile time).
template <class R, int SET>
B2_HD void tr_store_set(float2* blk, const RRegs<R>& r, int lane) {
#pragma unroll
    if constexpr (R::BP % 2 == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float2 a = r.v[column_of<R::W>(32 * SET + 2 * j)], b = r.v[column_of<R::W>(32 * SET + 2 * j + 1)];
            reinterpret_cast<float4*>(blk + lane * R::BP)[j] = make_float4(a.x, a.y, b.x, b.y);
        }
    } else {
#pragma unroll
        for (int sl = 0; sl < 32; ++sl) blk[lane * R::BP + sl] = r.v[column_of<R::W>(32 * SET + sl)];
    }
}
template <class R, int SET>
B2_HD void tr_load_set(const float2* blk, RRegs<R>& r, int lane) {
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) r.v[column_of<R::W>(32 * SET + rr)] = blk[rr * R::BP + lane];
}

#ifdef __CUDACC__
// Device-side transpose for one thread.  Block b of X is read by warp b; the off-diagonal 32x32 blocks need one CTA
// barrier, the diagonal ones are warp-synchronous.  `pre_barrier`: the other warp may still be reading its block.
// Barrier of one group: the whole CTA (one group per CTA, or lockstep groups), or - NAMED - hardware barrier `id` with the group's
// R::NT threads, so that the groups of a CTA run independently (piv_rows_tm_kernel).
template <bool NAMED, int NT>
__device__ __forceinline__ void group_barrier(int id) {
    if constexpr (NAMED) asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NT) : "memory");
    else __syncthreads();
}
template <class R, bool NAMED = false>
__device__ __forceinline__ void transpose_device(RSmem<R>& s, RRegs<R>& r, int tid, bool pre_barrier, int bar_id = 0) {
    const int lane = tid & 31, wq = tid >> 5;
    if constexpr (R::NWARP == 2) {
        if (pre_barrier) group_barrier<NAMED, R::NT>(bar_id);
        if (wq == 0) tr_store_set<R, 1>(s.X[1], r, lane); else tr_store_set<R, 0>(s.X[0], r, lane);
        group_barrier<NAMED, R::NT>(bar_id);
        if (wq == 0) {
            tr_load_set<R, 1>(s.X[0], r, lane);
            __syncwarp();
            tr_store_set<R, 0>(s.X[0], r, lane);
            __syncwarp();
            tr_load_set<R, 0>(s.X[0], r, lane);
        } else {
            tr_load_set<R, 0>(s.X[1], r, lane);
            __syncwarp();
            tr_store_set<R, 1>(s.X[1], r, lane);
            __syncwarp();
            tr_load_set<R, 1>(s.X[1], r, lane);
        }
    } else {
        __syncwarp();
        tr_store_set<R, 0>(s.X[0], r, lane);
        __syncwarp();
        tr_load_set<R, 0>(s.X[0], r, lane);
    }
}
#endif

// First frame of a run: park the scaled spectra only.
//   A0 = (z + conj zn) * (0.5/std0),  A1 = -i (z - conj zn) * (0.5/std1),  parked value = A / N^2
// `zn` = Z(-ky, -kx) comes from the partner lane (register index (W-ky)%W).
B2_HD void separate(float2 z, float2 zn, float b0, float b1, float2& a0, float2& a1) {
    a0 = pk_scale(pk_add(z, make_float2(zn.x, -zn.y)), b0);                         // ((z.x + zn.x) b0, (z.y - zn.y) b0)
    a1 = pk_scale(pk_add(make_float2(z.y, -z.x), make_float2(zn.y, zn.x)), b1);     // ((z.y + zn.y) b1, (zn.x - z.x) b1)
}

// One ky step (ky in [0, W/2]) of the cross phase, part A: needs partner's Z(kn) = `pz`.
// Computes R0, R1 at (ky, own column), parks the new spectra, writes G(ky) into v[ky].
// Scaling: the planes need 1 / N^2 (N = W * W: one 1/N for the unnormalised inverse transform, one for the mean of the
// correlation).  N is a power of two, so the native mode folds 1/N into the 0.5/std factor of BOTH roles of a spectrum (exact:
// only exponents change) and parks the separated spectra as they are; padded mode scales the parked copy by its 1 / (N ny nx).
template <class R, bool PAD = false>
B2_HD void cross_step_a(RSmem<R>& s, RRegs<R>& r, int tid, int ky, float2 pz, bool have_prev, float2& r0, float2& r1,
                        const RParams* pp = nullptr) {
    constexpr float INVN = 1.0f / (float)R::NPX;
    float2 a0, a1;
    if (PAD) separate(r.v[ky], pz, r.half_alpha_new[0], r.half_alpha_new[1], a0, a1);
    else separate(r.v[ky], pz, r.half_alpha_new[0] * INVN, r.half_alpha_new[1] * INVN, a0, a1);
    if (have_prev) {
        const float4 pk = s.park[ky][tid];
        r0 = ctw<0>(a0, pk.x, pk.y);   // conj(p0) * a0
        r1 = ctw<0>(a1, pk.z, pk.w);
        if (PAD) {   // new window in its tiled role: times T(ky, own column) = Ty(ky) Tx
            const float2 ty = pp->pad_ty[ky];
            const float2 t = ctw<1>(ty, r.tx.x, r.tx.y);
            r0 = ctw<1>(r0, t.x, t.y);
            r1 = ctw<1>(r1, t.x, t.y);
        }
    }
    if (PAD) {
        const float2 q0 = pk_scale(a0, pp->pad_scale), q1 = pk_scale(a1, pp->pad_scale);
        s.park[ky][tid] = make_float4(q0.x, q0.y, q1.x, q1.y);
    } else {
        s.park[ky][tid] = make_float4(a0.x, a0.y, a1.x, a1.y);
    }
    if (have_prev) r.v[ky] = pk_sub(make_float2(r0.x, -r0.y), make_float2(r1.y, r1.x));   // conj(G), G = R0 + i R1
}
// part B: the partner's (R0, R1) at (ky, -col) give G(-ky, col) = conj(R0) + i conj(R1), stored conjugated:
// conj(G(-ky, col)) = (R0.x + R1.y, R0.y - R1.x).  The SENDER forms that value (cross_mirror) so only one complex number
// crosses the warp per step.
B2_HD float2 cross_mirror(float2 r0, float2 r1) { return pk_add(r0, make_float2(r1.y, -r1.x)); }
template <class R>
B2_HD void cross_step_b(RRegs<R>& r, int ky, float2 q) {
    constexpr int W = R::W;
    if (ky != 0 && ky != W / 2) r.v[W - ky] = q;
}

#ifdef __CUDACC__
__device__ __forceinline__ float2 shfl2(float2 a, int src) {
    return make_float2(__shfl_sync(0xffffffffu, a.x, src), __shfl_sync(0xffffffffu, a.y, src));
}
// device: the whole cross phase for one thread
template <class R, bool PAD = false>
__device__ __forceinline__ void rows_p3b_device(RSmem<R>& s, RRegs<R>& r, int tid, bool have_prev, const RParams* pp = nullptr) {
    constexpr int W = R::W;
    const int pl = partner_lane_of<W>(tid);
    if (tid == 0) r.v[0] = pk_sub(r.v[0], make_float2(r.dc_fix[0], r.dc_fix[1]));   // Z(0, 0): thread 0 owns column 0 (rows_p2_pre / _pad)
#pragma unroll
    for (int ky = 0; ky <= W / 2; ++ky) {
        const float2 pz = shfl2(r.v[(W - ky) % W], pl);
        float2 r0 = make_float2(0.f, 0.f), r1 = make_float2(0.f, 0.f);
        cross_step_a<R, PAD>(s, r, tid, ky, pz, have_prev, r0, r1, pp);
        if (have_prev && ky != 0 && ky != W / 2) cross_step_b<R>(r, ky, shfl2(cross_mirror(r0, r1), pl));
    }
}
#endif

// ------------------------------------------------------------------------------------------------------------
// Parked spectra in Tensor Memory.  The 33.8 KB of parked spectra per 64 x 64 group are what limits an SM to four groups
// (8 warps, two per scheduler - far too few to hide the latencies of the non-FFT phases: ncu, round 2).  Tensor Memory
// (256 KB per SM, idle in a kernel without tensor-core work) holds them instead: with the .32x32b shape thread t of a warp
// owns lane 32 * (warp % 4) + t, i.e. a private row of 32-bit columns - exactly "the spectra of my column".  A thread's
// (A0, A1)(ky) are 4 columns at 4 * ky; the warps that share a lane quarter use disjoint column ranges.  Loads are issued
// one batch of two ky steps ahead (tcgen05.wait::ld waits for everything outstanding), stores are fire-and-forget until
// the end of the phase.  Verified as plain scratch memory by tools/tmem_probe.cu (profiles/r02/tmem_probe.log).
// ------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void tm_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_ld4(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tm_st4(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
// wait for the outstanding loads; the registers pass through the statement so that no use of them can be scheduled above it
__device__ __forceinline__ void tm_wait_ld(uint32_t (&v)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]) :: "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// the cross phase of rows_p3b_device with the parked spectra in Tensor Memory (native 64 x 64 / 32 x 32 windows)
template <class R>
__device__ __forceinline__ void rows_p3b_tm(RRegs<R>& r, int tid, uint32_t tm) {
    constexpr int W = R::W;
    constexpr float INVN = 1.0f / (float)R::NPX;
    constexpr int NB = (W / 2 + 1 + 1) / 2;          // batches of two ky steps; the last one holds ky = W / 2 alone
    const int pl = partner_lane_of<W>(tid);
    const float b0 = r.half_alpha_new[0] * INVN, b1 = r.half_alpha_new[1] * INVN;
    if (tid == 0) r.v[0] = pk_sub(r.v[0], make_float2(r.dc_fix[0], r.dc_fix[1]));   // Z(0, 0): thread 0 owns column 0 (rows_p2_pre)
    uint32_t pk[2][8];
    tm_ld8(tm, pk[0]);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        uint32_t (&cur)[8] = pk[b & 1];
        tm_wait_ld(cur);
        if (b + 1 < NB) {
            if (2 * (b + 1) + 1 <= W / 2) tm_ld8(tm + 8 * (b + 1), pk[(b + 1) & 1]); else tm_ld4(tm + 8 * (b + 1), pk[(b + 1) & 1]);
        }
        uint32_t out[8];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int ky = 2 * b + j;
            if (ky <= W / 2) {
                const float2 pz = shfl2(r.v[(W - ky) % W], pl);
                float2 a0, a1;
                separate(r.v[ky], pz, b0, b1, a0, a1);
                const float2 r0 = ctw<0>(a0, __uint_as_float(cur[4 * j]), __uint_as_float(cur[4 * j + 1]));       // conj(p0) * a0
                const float2 r1 = ctw<0>(a1, __uint_as_float(cur[4 * j + 2]), __uint_as_float(cur[4 * j + 3]));
                out[4 * j] = __float_as_uint(a0.x); out[4 * j + 1] = __float_as_uint(a0.y);
                out[4 * j + 2] = __float_as_uint(a1.x); out[4 * j + 3] = __float_as_uint(a1.y);
                r.v[ky] = pk_sub(make_float2(r0.x, -r0.y), make_float2(r1.y, r1.x));                                // conj(G), G = R0 + i R1
                if (ky != 0 && ky != W / 2) cross_step_b<R>(r, ky, shfl2(cross_mirror(r0, r1), pl));
            }
        }
        if (2 * b + 1 <= W / 2) tm_st8(tm + 8 * b, out); else tm_st4(tm + 8 * b, out);
    }
    tm_wait_st();
}
#endif

// ------------------------------------------------------------------------------------------------------------
// Displaced windows (second pass of the two-pass scheme, multipass.cuh): frame k+1's window of a pair is read at
// (y0 + dy, x0 + dx), any byte offset.  The TMA box starts at the 16-byte boundary below (16 bytes wider, no swizzle, like
// ALIGNED = false); rows are funnel-shifted to the byte offset.  The displaced window only serves as `b` of its pair, the
// undisplaced one as `a` of the next pair, so a frame costs two forward transforms: cross-only for the displaced tile,
// park-only for the undisplaced one.
// ------------------------------------------------------------------------------------------------------------
template <class R>
B2_HD void rows_p1_shift(RSmem<R>& s, RRegs<R>& r, int tid, const unsigned char* tile, int xoff0, int xoff1) {
    constexpr int W = R::W;
    const int row = column_of<W>(tid);
    unsigned S[2] = {0, 0}, Q[2] = {0, 0};
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        const int xoff = w == 0 ? xoff0 : xoff1;
        const unsigned char* base = tile + (w * W + row) * R::WB + (xoff & ~3);
        const int sh = (xoff & 3) * 8;
        unsigned wd[W / 4 + 1];
#pragma unroll
        for (int k = 0; k <= W / 4; ++k) wd[k] = *reinterpret_cast<const unsigned*>(base + 4 * k);
#pragma unroll
        for (int k = 0; k < W / 4; ++k) {
            r.px[w][k] = funnel_r(wd[k], wd[k + 1], sh);
            S[w] = dp4a_u(r.px[w][k], 0x01010101u, S[w]);
            Q[w] = dp4a_u(r.px[w][k], r.px[w][k], Q[w]);
        }
    }
    unsigned vals[4] = {S[0], Q[0], S[1], Q[1]};
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int k = 0; k < 4; ++k) vals[k] = __reduce_add_sync(0xffffffffu, vals[k]);   // one REDUX each (exact: sums stay below 2^32)
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) s.red[tid >> 5][k] = vals[k];
    }
#else
    for (int k = 0; k < 4; ++k) s.red[tid >> 5][k] += vals[k];
#endif
}

#ifdef __CUDACC__
// displaced tile: cross spectra against the parked (undisplaced, previous frame) spectra, nothing is parked
template <class R>
__device__ __forceinline__ void rows_cross_only_device(RSmem<R>& s, RRegs<R>& r, int tid) {
    constexpr int W = R::W;
    const int pl = partner_lane_of<W>(tid);
#pragma unroll
    for (int ky = 0; ky <= W / 2; ++ky) {
        const float2 pz = shfl2(r.v[(W - ky) % W], pl);
        float2 a0, a1;
        separate(r.v[ky], pz, r.half_alpha_new[0], r.half_alpha_new[1], a0, a1);
        const float4 pk = s.park[ky][tid];
        const float2 r0 = ctw<0>(a0, pk.x, pk.y);   // conj(p0) * a0
        const float2 r1 = ctw<0>(a1, pk.z, pk.w);
        r.v[ky] = pk_sub(make_float2(r0.x, -r0.y), make_float2(r1.y, r1.x));                   // conj(G), G = R0 + i R1
        if (ky != 0 && ky != W / 2) cross_step_b<R>(r, ky, shfl2(cross_mirror(r0, r1), pl));
    }
}
// undisplaced tile: park the scaled spectra for the next pair
template <class R>
__device__ __forceinline__ void rows_park_only_device(RSmem<R>& s, RRegs<R>& r, int tid) {
    constexpr int W = R::W;
    constexpr float INVN2 = 1.0f / ((float)R::NPX * (float)R::NPX);
    const int pl = partner_lane_of<W>(tid);
#pragma unroll
    for (int ky = 0; ky <= W / 2; ++ky) {
        const float2 pz = shfl2(r.v[(W - ky) % W], pl);
        float2 a0, a1;
        separate(r.v[ky], pz, r.half_alpha_new[0], r.half_alpha_new[1], a0, a1);
        const float2 q0 = pk_scale(a0, INVN2), q1 = pk_scale(a1, INVN2);
        s.park[ky][tid] = make_float4(q0.x, q0.y, q1.x, q1.y);
    }
}
#endif

// ------------------------------------------------------------------------------------------------------------
// P4: (forward FFT of conj G along columns, then) store column into X.
// P5: row from X, (forward FFT along the row, then) conjugate, clip, per-row max / sum.
// ------------------------------------------------------------------------------------------------------------
// r.v holds FFT2(conj G) = conj(planes): plane 0 = Re, plane 1 = -Im.
// Padded mode: the lags 0 .. n-1 of the (unshifted) W x W plane are free of wrap-around - the tiled window covers
// [0, 2n) - and by periodicity lag q >= n/2 IS the reference's negative lag q - n.  The reference's fftshifted n x n plane
// is therefore element (q + n/2) % n <- natural index q < n, per axis; everything else is set to 0, which is neutral
// for the max and the sum of the clipped (>= 0) plane.
B2_HD int shifted_index(int q, int n) { const int h = n / 2; return q < n - h ? q + h : q - (n - h); }   // (q + n/2) % n, q < n

template <class R, bool PAD = false>
B2_HD void rows_p5_post(RSmem<R>& s, RRegs<R>& r, int tid, bool dead0, bool dead1, const RParams* pp = nullptr) {
    constexpr int W = R::W;
    float m0 = 0.f, m1 = 0.f;
    float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);   // two running (window 0, window 1) sums: shorter dependency chains
#pragma unroll
    for (int x = 0; x < W; x += 2) {
        // clip to [0, 1] (inputs are uint8, no NaNs can occur, so fmin/fmax are exact here); padded mode: the upper bound
        // is 0 for the columns outside the window's lags (rows outside are dropped below and never read afterwards)
        const float h0 = PAD ? pp->pad_cm[x] : 1.f, h1 = PAD ? pp->pad_cm[x + 1] : 1.f;
        const float2 e0 = make_float2(fminf(fmaxf(r.v[x].x, 0.f), h0), fminf(fmaxf(-r.v[x].y, 0.f), h0));
        const float2 e1 = make_float2(fminf(fmaxf(r.v[x + 1].x, 0.f), h1), fminf(fmaxf(-r.v[x + 1].y, 0.f), h1));
        r.v[x] = e0; r.v[x + 1] = e1;
        m0 = fmaxf(fmaxf(e0.x, e1.x), m0); m1 = fmaxf(fmaxf(e0.y, e1.y), m1);
        sa = pk_add(sa, e0); sb = pk_add(sb, e1);
    }
    sa = pk_add(sa, sb);
    float s0 = sa.x, s1 = sa.y;
    // a window with zero variance has an exactly-zero plane in the reference (the packed inverse FFT leaves ~1e-10
    // cross-talk from its partner window): force max = sum = 0 here, first-argmax 0 in rows_p6, zeros in the dumps
    if (dead0 || (PAD && column_of<W>(tid) >= pp->ny)) { m0 = 0.f; s0 = 0.f; }
    if (dead1 || (PAD && column_of<W>(tid) >= pp->ny)) { m1 = 0.f; s1 = 0.f; }
    r.dead[0] = dead0; r.dead[1] = dead1;
    r.rowmax[0] = m0; r.rowmax[1] = m1; r.rowsum[0] = s0; r.rowsum[1] = s1;
#ifdef __CUDA_ARCH__
    // warp maxima with one REDUX each (non-negative floats order like their bit patterns), the first row (reference order)
    // that holds the warp maximum with another, sums by packed shuffle-adds
    const unsigned b0 = __float_as_uint(m0), b1 = __float_as_uint(m1);
    const unsigned wm0 = __reduce_max_sync(0xffffffffu, b0), wm1 = __reduce_max_sync(0xffffffffu, b1);
    {
        const bool rowok = !PAD || column_of<W>(tid) < pp->ny;
        const unsigned si = PAD ? (rowok ? (unsigned)shifted_index(column_of<W>(tid), pp->ny) : 0xffffu) : (unsigned)((column_of<W>(tid) + W / 2) % W);
        const unsigned c0 = (b0 == wm0 && rowok) ? si : 0xffffu, c1 = (b1 == wm1 && rowok) ? si : 0xffffu;
        const unsigned wr0 = __reduce_min_sync(0xffffffffu, c0), wr1 = __reduce_min_sync(0xffffffffu, c1);
        float2 ss = make_float2(s0, s1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss = pk_add(ss, make_float2(__shfl_xor_sync(0xffffffffu, ss.x, o), __shfl_xor_sync(0xffffffffu, ss.y, o)));
        if ((tid & 31) == 0) {
            s.red[tid >> 5][4] = wm0; s.red[tid >> 5][5] = wm1;
            s.red[tid >> 5][6] = __float_as_uint(ss.x); s.red[tid >> 5][7] = __float_as_uint(ss.y);
            s.rowc[tid >> 5][0] = wr0; s.rowc[tid >> 5][1] = wr1;
        }
    }
#else
    union { float f; unsigned u; } a;
    const int wp = tid >> 5;
    a.u = s.red[wp][4]; a.f = a.f > m0 ? a.f : m0; s.red[wp][4] = a.u;
    a.u = s.red[wp][5]; a.f = a.f > m1 ? a.f : m1; s.red[wp][5] = a.u;
    a.u = s.red[wp][6]; a.f += s0; s.red[wp][6] = a.u;
    a.u = s.red[wp][7]; a.f += s1; s.red[wp][7] = a.u;
#endif
}

B2_HD float bits_f(unsigned u) { union { float f; unsigned u; } a; a.u = u; return a.f; }

// P6: block max / sum known; rows holding the max search their FIRST matching column in fftshifted order and
// deposit key = shifted flat index (min wins).
template <class R, bool PAD = false>
B2_HD void rows_p6(RSmem<R>& s, RRegs<R>& r, int tid, const RParams* pp = nullptr) {
    constexpr int W = R::W;
    const bool rowok = !PAD || column_of<W>(tid) < pp->ny;
    // fftshifted index of this thread's row in the reference's plane
    const int si = PAD ? (rowok ? shifted_index(column_of<W>(tid), pp->ny) : 0) : (column_of<W>(tid) + W / 2) % W;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        float M = 0.f, S = 0.f;
#pragma unroll
        for (int k = 0; k < R::NWARP; ++k) { M = fmaxf(M, bits_f(s.red[k][4 + w])); S += bits_f(s.red[k][6 + w]); }
        r.cmaxv[w] = M; r.sumv[w] = S;
        unsigned long long key = ~0ull;
        if (r.rowmax[w] == M && rowok) {
            int first = W;
            if (PAD) {
                // match mask over natural columns, then first match in the reference's order: lags >= nx - nx/2 (its
                // columns 0 ..) before lags < nx - nx/2.  M == 0: every element of the plane is the maximum -> first = 0.
                unsigned mlo = 0u, mhi = 0u;
#pragma unroll
                for (int x = 0; x < W; ++x) {
                    const float val = w == 0 ? r.v[x].x : r.v[x].y;
                    if (x < 32) mlo |= (val == M) ? (1u << x) : 0u; else mhi |= (val == M) ? (1u << (x - 32)) : 0u;
                }
                const unsigned long long mm = ((unsigned long long)mhi << 32) | mlo;
                const int nh = pp->nx - pp->nx / 2;
                const unsigned long long seg_hi = mm >> nh, seg_lo = mm & ((1ull << nh) - 1ull);
                unsigned long long t = seg_hi ? seg_hi : seg_lo;
#ifdef __CUDA_ARCH__
                const int k = __ffsll((long long)t) - 1;
#else
                int k = 0;
                while (!(t & 1ull) && k < 63) { t >>= 1; ++k; }
#endif
                first = M == 0.f ? 0 : (seg_hi ? k : k + pp->nx / 2);
            } else {
                // shifted column j = (x + W/2) % W ; scan j descending so the smallest j survives
#pragma unroll
                for (int j = W - 1; j >= 0; --j) {
                    const int x = (j + W / 2) % W;
                    const float val = w == 0 ? r.v[x].x : r.v[x].y;
                    first = (val == M) ? j : first;
                }
            }
            if (r.dead[w]) first = 0;   // all-zero plane: every element is the maximum
            key = (unsigned long long)(si * W + first);
        }
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        if ((tid & 31) == 0) s.redk[tid >> 5][w] = key;
#else
        if (key < s.redk[tid >> 5][w]) s.redk[tid >> 5][w] = key;
#endif
    }
}

#ifdef __CUDACC__
// P6x (device, replaces P6 + P7 of the first version and one CTA barrier): after E1 every thread knows the block maximum AND
// the peak's row - the first row in reference order among the warps' candidates (rows_p5_post) - so the thread that owns that
// row looks for the first matching column in its registers while the rows around it are dumped for the Gaussian fit; only the
// column has to travel (peak_j).  First-occurrence argmax of the flattened plane = smallest row holding the maximum, then the
// smallest column in that row.
template <class R, bool PAD = false>
__device__ __forceinline__ void rows_p6x(RSmem<R>& s, RRegs<R>& r, int tid, const RParams* pp = nullptr) {
    constexpr int W = R::W;
    const bool rowok = !PAD || column_of<W>(tid) < pp->ny;
    const int si = PAD ? (rowok ? shifted_index(column_of<W>(tid), pp->ny) : -8) : (column_of<W>(tid) + W / 2) % W;
    float* nb = &s.nb[0][0][0];   // [w][3][W]
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        unsigned mb = 0u, row = 0xffffu;
        float S = 0.f;
#pragma unroll
        for (int k = 0; k < R::NWARP; ++k) {
            const unsigned m = s.red[k][4 + w], rc = s.rowc[k][w];
            row = m > mb ? rc : (m == mb ? min(row, rc) : row);
            mb = max(mb, m);
            S += bits_f(s.red[k][6 + w]);
        }
        const float M = __uint_as_float(mb);
        r.cmaxv[w] = M; r.sumv[w] = S;
        const int pi = r.dead[w] ? 0 : (int)row;      // all-zero plane: every element is the maximum -> flat index 0
        r.pi[w] = pi;
        if (si == pi) {
            int first = 0;
            if (!r.dead[w]) {
                if (PAD) {
                    unsigned mlo = 0u, mhi = 0u;
#pragma unroll
                    for (int x = 0; x < W; ++x) {
                        const float val = w == 0 ? r.v[x].x : r.v[x].y;
                        if (x < 32) mlo |= (val == M) ? (1u << x) : 0u; else mhi |= (val == M) ? (1u << (x - 32)) : 0u;
                    }
                    const unsigned long long mm = ((unsigned long long)mhi << 32) | mlo;
                    const int nh = pp->nx - pp->nx / 2;
                    const unsigned long long seg_hi = mm >> nh, seg_lo = mm & ((1ull << nh) - 1ull);
                    const unsigned long long t = seg_hi ? seg_hi : seg_lo;
                    const int k = __ffsll((long long)t) - 1;
                    first = M == 0.f ? 0 : (seg_hi ? k : k + pp->nx / 2);
                } else {
                    first = W;
#pragma unroll
                    for (int j = W - 1; j >= 0; --j) {   // shifted column j = (x + W/2) % W ; scan j descending so the smallest j survives
                        const int x = (j + W / 2) % W;
                        const float val = w == 0 ? r.v[x].x : r.v[x].y;
                        first = (val == M) ? j : first;
                    }
                }
            }
            s.peak_j[w] = first;
        }
        const int d = si - pi;
        if (d >= -1 && d <= 1) {
#pragma unroll
            for (int x = 0; x < W; ++x) {
                const float val = r.dead[w] ? 0.f : (w == 0 ? r.v[x].x : r.v[x].y);
                nb[(w * 3 + (d + 1)) * W + (PAD ? x : (x + W / 2) % W)] = val;   // padded: natural lag order, mapped in rows_p8
            }
        }
    }
}
#endif

// P7: peak position known; the three rows around each peak are dumped for the Gaussian fit.
template <class R, bool PAD = false>
B2_HD void rows_p7(RSmem<R>& s, RRegs<R>& r, int tid, const RParams* pp = nullptr) {
    constexpr int W = R::W;
    const bool rowok = !PAD || column_of<W>(tid) < pp->ny;
    const int si = PAD ? (rowok ? shifted_index(column_of<W>(tid), pp->ny) : -8) : (column_of<W>(tid) + W / 2) % W;
    float* nb = &s.nb[0][0][0];   // [w][3][W]
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        unsigned long long key = ~0ull;
#pragma unroll
        for (int k = 0; k < R::NWARP; ++k) key = s.redk[k][w] < key ? s.redk[k][w] : key;
        const int idx = (int)key;
        r.pi[w] = idx / W; r.pj[w] = idx % W;
        const int d = si - r.pi[w];
        if (d >= -1 && d <= 1) {
#pragma unroll
            for (int x = 0; x < W; ++x) {
                const float val = r.dead[w] ? 0.f : (w == 0 ? r.v[x].x : r.v[x].y);
                nb[(w * 3 + (d + 1)) * W + (PAD ? x : (x + W / 2) % W)] = val;   // padded: natural lag order, mapped in rows_p8
            }
        }
    }
}

// P8: Gaussian fit + outputs by thread w (pyorc/velocimetry/ffpiv.py:465-466 + ffpiv.u_v_displacement).
template <class R, bool PAD = false>
B2_HD void rows_p8(RSmem<R>& s, RRegs<R>& r, int tid, const RParams& p, const RUnit& un, int pair, bool pj_from_smem = false) {
    constexpr int W = R::W;
    if (tid >= 2) return;
    if (pj_from_smem) { r.pj[0] = s.peak_j[0]; r.pj[1] = s.peak_j[1]; }
    const int ny = PAD ? p.ny : W, nx = PAD ? p.nx : W;
    const int w = tid;
    if (w == 1 && !un.valid1) return;
    const float* nb = &s.nb[0][0][0] + w * 3 * W;
    // selects, not r.x[w]: a dynamic index would demote the whole register struct to local memory
    const int pi = w == 0 ? r.pi[0] : r.pi[1], pj = w == 0 ? r.pj[0] : r.pj[1];
    const float cmax = w == 0 ? r.cmaxv[0] : r.cmaxv[1];
    const float mean = (w == 0 ? r.sumv[0] : r.sumv[1]) / (PAD ? (float)(ny * nx) : (float)R::NPX);
    float uu, vv;
    if (pi == 0 || pi == ny - 1 || pj == 0 || pj == nx - 1) {
        if (p.border_nan) { uu = nanf(""); vv = nanf(""); }
        else { uu = (float)(pj - nx / 2); vv = (float)(pi - ny / 2); }
    } else {
        const float eps = p.gauss_eps;
        const float lc = logf(cmax + eps);
        // padded mode keeps the rows in natural lag order: reference column j <- lag (j + nx - nx/2) % nx
        const int h = nx / 2;
        const int q0 = PAD ? (pj < h ? pj + nx - h : pj - h) : pj;
        const int qm = PAD ? (pj - 1 < h ? pj - 1 + nx - h : pj - 1 - h) : pj - 1;
        const int qp = PAD ? (pj + 1 < h ? pj + 1 + nx - h : pj + 1 - h) : pj + 1;
        const float ll = logf(nb[0 * W + q0] + eps), lr = logf(nb[2 * W + q0] + eps);
        const float ld = logf(nb[1 * W + qm] + eps), lu = logf(nb[1 * W + qp] + eps);
        vv = ((float)pi + (ll - lr) / (2.f * ll - 4.f * lc + 2.f * lr)) - (float)(ny / 2);
        uu = ((float)pj + (ld - lu) / (2.f * ld - 4.f * lc + 2.f * lu)) - (float)(nx / 2);
    }
    float oc = cmax, os = cmax / mean;
    const int widx = w == 0 ? un.w[0] : un.w[1];
    if (p.keep && !p.keep[widx]) { uu = vv = oc = os = nanf(""); }
    const int opair = p.pair_step == 2 ? pair / 2 : pair;
    const long long o = (long long)opair * p.n_rows * p.n_cols + widx;
    if (p.shift) { vv += (float)p.shift[2 * o]; uu += (float)p.shift[2 * o + 1]; }
    if (p.fshift) { vv += p.fshift[2 * o]; uu += p.fshift[2 * o + 1]; }
    p.u[o] = uu; p.v[o] = vv; p.cmax[o] = oc; p.s2n[o] = os;
    if (p.peer.n) peer_store(p.peer, opair, (long long)p.n_rows * p.n_cols, widx, uu, vv, oc, os);
}

// optional triage dump of the full planes (fftshifted, clipped) - every thread writes its row
template <class R, bool PAD = false>
B2_HD void rows_dump_planes(RRegs<R>& r, int tid, const RParams& p, const RUnit& un, int pair) {
    constexpr int W = R::W;
    if (!p.planes) return;
    // padded mode dumps the whole W x W plane in NATURAL lag order (planes_reorder_kernel crops / shifts it afterwards):
    // no per-element conditions in this kernel's instruction stream
    const int si = PAD ? column_of<W>(tid) : (column_of<W>(tid) + W / 2) % W;
    const long long nw = (long long)p.n_rows * p.n_cols;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        if (w == 1 && !un.valid1) continue;
        float* dst = p.planes + (((long long)pair * nw + un.w[w]) * W + si) * W;
#pragma unroll
        for (int x = 0; x < W; ++x) dst[PAD ? x : (x + W / 2) % W] = r.dead[w] ? 0.f : (w == 0 ? r.v[x].x : r.v[x].y);
    }
}

// Ensemble mode (pyorc/velocimetry/ffpiv.py:200-243 thresholds, :359-365 accumulation): instead of locating the peak of
// every pair, a plane that passes corr_min / s2n_min is ADDED to the window's accumulator plane in HBM (fftshifted
// coordinates, the layout ens_finish_kernel reads).  A work unit owns its two windows for every frame of the launch
// (run_len = n_pairs): all additions to one accumulator element come from one thread in frame order, like the
// reference's np.sum(corr, axis=0); thread t updates row sigma(t) with 16-byte reductions.  The accumulators of a 1080p grid are
// 31 MB, i.e. they live in L2 between frame pairs.
template <class R, bool PAD = false>
B2_HD void rows_ens(RSmem<R>& s, RRegs<R>& r, int tid, const RParams& p, const RUnit& un, int pair, bool store) {
    constexpr int W = R::W;
    const int ny = PAD ? p.ny : W, nx = PAD ? p.nx : W;
    const bool rowok = !PAD || column_of<W>(tid) < ny;
    const int si = PAD ? (rowok ? shifted_index(column_of<W>(tid), ny) : 0) : (column_of<W>(tid) + W / 2) % W;
    const long long nw = (long long)p.n_rows * p.n_cols;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        float M = 0.f, S = 0.f;
#pragma unroll
        for (int k = 0; k < R::NWARP; ++k) { M = fmaxf(M, bits_f(s.red[k][4 + w])); S += bits_f(s.red[k][6 + w]); }
        const float ratio = M / (S / (PAD ? (float)(ny * nx) : (float)R::NPX));
        const int widx = w == 0 ? un.w[0] : un.w[1];
        bool ok = (M >= p.corr_min) && (ratio >= p.s2n_min) && !r.dead[w];   // dead: 0 / 0 = NaN fails the test in the reference
        if (p.keep && !p.keep[widx]) ok = false;                             // NaN plane in the reference -> masked out
        if (!store || (w == 1 && !un.valid1)) continue;
        if (ok && PAD) {
            // accumulator planes are [n_windows][ny][nx]: rows are not 16-byte aligned, scalar reductions on the block
            if (rowok) {
                // lag x goes to column x + nx/2 (x < nx - nx/2) or x - (nx - nx/2): two base pointers, static offsets
                float* row = p.ens_sum + ((long long)widx * ny + si) * nx;
                float* base_lo = row + nx / 2;
                float* base_hi = row - (nx - nx / 2);
                const int nh = nx - nx / 2;
#pragma unroll
                for (int x = 0; x < W; ++x) {
                    if (x < nx) {
                        float* dst = (x < nh ? base_lo : base_hi) + x;
#ifdef __CUDA_ARCH__
                        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst), "f"(w == 0 ? r.v[x].x : r.v[x].y) : "memory");
#else
                        *dst += w == 0 ? r.v[x].x : r.v[x].y;
#endif
                    }
                }
            }
        } else if (ok) {
            float* dst = p.ens_sum + ((long long)widx * W + si) * W;
#pragma unroll
            for (int c = 0; c < W / 4; ++c) {
                const int x = 4 * c, j = (x + W / 2) % W;   // shifted column of element x; W/2 is a multiple of 4
#ifdef __CUDA_ARCH__
                // fire-and-forget 16-byte reduction at the L2 (sm_90+): no load latency to hide, no registers for the old
                // values; the additions to one address all come from this thread in program order, so the sum order is
                // still the frame order.  (A load-add-store version cost 45 % more time per frame.)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j),
                             "f"(w == 0 ? r.v[x].x : r.v[x].y), "f"(w == 0 ? r.v[x + 1].x : r.v[x + 1].y),
                             "f"(w == 0 ? r.v[x + 2].x : r.v[x + 2].y), "f"(w == 0 ? r.v[x + 3].x : r.v[x + 3].y)
                             : "memory");
#else
                for (int q = 0; q < 4; ++q) dst[j + q] += w == 0 ? r.v[x + q].x : r.v[x + q].y;
#endif
            }
        }
        if (tid == 0) {
            const long long o = (long long)pair * nw + widx;
            p.cmax[o] = ok ? M : 0.f;
            p.s2n[o] = ok ? ratio : 0.f;
            if (ok && M > 1e-6f) p.ens_count[widx] += 1.f;
        }
    }
}

}  // namespace b2piv

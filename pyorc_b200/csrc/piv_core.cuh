// piv_core.cuh - barrier-delimited phases of the fused LSPIV interrogation kernel (sm_100a).
//
// One work item = TWO interrogation windows (w0, w1) of one frame pair (frame k, frame k+1).  Each window pair
// (a = window in frame k, b = window in frame k+1) is packed as z = a + i*b into ONE complex plane, so a single
// complex 2-D FFT yields both spectra; the cross spectra R_w = conj(A_w) * B_w of the two windows are then packed
// as G = R_0 + i*R_1 and ONE complex inverse FFT yields both real correlation planes (Re -> w0, Im -> w1).
// That is 1.5 complex 64x64 FFTs per window, all in shared memory / registers; nothing but the 8 KB of source
// pixels and 16 B of results per window ever touches HBM.
//
// Replaces (reference, CPU): ffpiv.cross_corr + np.nanmax/np.nanmean + ffpiv.u_v_displacement as called from
// pyorc/velocimetry/ffpiv.py:446-474 (per-time-step) and :200-243 (ensemble, see ens kernels in b2piv.cu).
//
// Every function here is __host__ __device__: tests/emul compiles the same phases for the CPU and runs them
// "thread by thread, phase by phase" to check the index mathematics without a GPU.  The emulator is test-only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define B2_HD __host__ __device__ __forceinline__

namespace b2piv {

// ------------------------------------------------------------------------------------------------------------
// complex helpers
// ------------------------------------------------------------------------------------------------------------
// Blackwell packed fp32 arithmetic (PTX add / sub / mul / fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2): ONE instruction works on
// a register pair, i.e. on a whole complex number.  Measured on B200 (tools/fp32x2_bench.cu, profiles/r02/fp32x2_bench.log):
// the same 128 values / clk / SM as the scalar forms - no extra flops - but HALF the issue slots and half the code bytes, and
// ptxas folds component swaps and sign changes (x -> (x.y, -x.x), the +-i rotations of a butterfly) into operand selectors
// (`R14.F32x2.LO_HI.NP`) and broadcasts scalar factors (`UR6.F32`), so a complex add is 1 instruction instead of 2 and a twiddle
// multiplication 2 instead of 4.  The row-per-thread kernels were limited by instruction issue and fetch (ncu, round 1: 60 %
// issue-active, the FFT passes stalled on `no_instruction` only), which is exactly what this relieves.  The host build (tests/emul)
// keeps the scalar forms; both are round-to-nearest IEEE operations, the fused multiply-adds being chosen explicitly here.
#ifdef __CUDA_ARCH__
#define B2_PK_IN(a) "l"(pk_bits(a))
__device__ __forceinline__ unsigned long long pk_bits(float2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 pk_float2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ float2 pk_add(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : B2_PK_IN(a), B2_PK_IN(b));
    return pk_float2(d);
}
__device__ __forceinline__ float2 pk_sub(float2 a, float2 b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : B2_PK_IN(a), B2_PK_IN(b));
    return pk_float2(d);
}
__device__ __forceinline__ float2 pk_mul(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : B2_PK_IN(a), B2_PK_IN(b));
    return pk_float2(d);
}
__device__ __forceinline__ float2 pk_fma(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : B2_PK_IN(a), B2_PK_IN(b), B2_PK_IN(c));
    return pk_float2(d);
}
#else
inline float2 pk_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 pk_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
inline float2 pk_mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 pk_fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
B2_HD float2 pk_scale(float2 a, float s) { return pk_mul(a, make_float2(s, s)); }

B2_HD float2 cadd(float2 a, float2 b) { return pk_add(a, b); }
B2_HD float2 csub(float2 a, float2 b) { return pk_sub(a, b); }
// a * (c - i s)  (INV = 0)   or   a * (c + i s)  (INV = 1):  c * (a.x, a.y) + s * (+-a.y, -+a.x)
template <int INV>
B2_HD float2 ctw(float2 a, float c, float s) {
    const float2 r = INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
    return pk_fma(r, make_float2(s, s), pk_mul(a, make_float2(c, c)));
}
// a * w  (INV=0)   or   a * conj(w)  (INV=1)
template <int INV>
B2_HD float2 cmulw(float2 a, float2 w) {
    return INV == 0 ? ctw<1>(a, w.x, w.y) : ctw<0>(a, w.x, w.y);
}

// cos/sin(2*pi*j/16), j = 0..7, as compile-time literals (fold to immediates after unrolling)
B2_HD constexpr float cos16(int j) {
    return j == 0 ? 1.0f : j == 1 ? 0.92387953251128674f : j == 2 ? 0.70710678118654752f : j == 3 ? 0.38268343236508977f
         : j == 4 ? 0.0f : j == 5 ? -0.38268343236508977f : j == 6 ? -0.70710678118654752f : -0.92387953251128674f;
}
B2_HD constexpr float sin16(int j) {
    return j == 0 ? 0.0f : j == 1 ? 0.38268343236508977f : j == 2 ? 0.70710678118654752f : j == 3 ? 0.92387953251128674f
         : j == 4 ? 1.0f : j == 5 ? 0.92387953251128674f : j == 6 ? 0.70710678118654752f : 0.38268343236508977f;
}

// ------------------------------------------------------------------------------------------------------------
// In-register DFT of R in {1,2,4,8,16} points, natural order in -> natural order out.
// Forward kernel exp(-2*pi*i*n*k/R) for INV=0, conjugate for INV=1.  Unnormalised.
// ------------------------------------------------------------------------------------------------------------
template <int R, int INV>
struct RegDFT {
    static B2_HD void run(float2* v) {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        RegDFT<R / 2, INV>::run(e);
        RegDFT<R / 2, INV>::run(o);
#pragma unroll
        for (int k = 0; k < R / 2; ++k) {
            // t = o[k] * w_R^k  (forward: w = exp(-2 pi i k / R))
            float2 t;
            const int j = k * (16 / R);  // index into the 16th roots table
            if (k == 0) {
                t = o[k];
            } else if (4 * k == R) {  // w = -i (fwd) / +i (inv)
                t = INV ? make_float2(-o[k].y, o[k].x) : make_float2(o[k].y, -o[k].x);
            } else {
                const float c = cos16(j), s = sin16(j);  // w = c - i s (fwd), c + i s (inv)
                t = ctw<INV>(o[k], c, s);
            }
            v[k] = cadd(e[k], t);
            v[k + R / 2] = csub(e[k], t);
        }
    }
};
template <int INV>
struct RegDFT<1, INV> {
    static B2_HD void run(float2*) {}
};

// ------------------------------------------------------------------------------------------------------------
// 1-D length factorisation N = R1 * R2.  Forward: strided R1-point DFTs (n = R2*n1 + n2), twiddle
// w_N^(n2*k1), then contiguous R2-point DFTs; everything in place, spectrum left in "digit-swapped" order:
// frequency k = k1 + R1*k2 lives at position p = k1*R2 + k2.  The inverse runs the mirror image
// (contiguous R2, conj twiddle, strided R1) and returns natural order, so no reordering pass ever runs.
// ------------------------------------------------------------------------------------------------------------
template <int N> struct Factor;
template <> struct Factor<16>  { static constexpr int R1 = 4,  R2 = 4; };
template <> struct Factor<32>  { static constexpr int R1 = 8,  R2 = 4; };
template <> struct Factor<64>  { static constexpr int R1 = 8,  R2 = 8; };
template <> struct Factor<128> { static constexpr int R1 = 16, R2 = 8; };

template <int N>
B2_HD int freq_of_pos(int p) { return p / Factor<N>::R2 + Factor<N>::R1 * (p % Factor<N>::R2); }
template <int N>
B2_HD int pos_of_freq(int k) { return (k % Factor<N>::R1) * Factor<N>::R2 + k / Factor<N>::R1; }
template <int N>
B2_HD int negpos(int p) { return pos_of_freq<N>((N - freq_of_pos<N>(p)) % N); }

// Kernel-wide compile-time configuration.
template <int WY_, int WX_, int NT_, int NWIN_, bool PADDED_ = false>
struct Cfg {
    static constexpr int WY = WY_, WX = WX_, NT = NT_, NWIN = NWIN_;  // NWIN windows per work item (1 or 2)
    // PADDED: the true window (Params::ny, nx - run time) is at most half the FFT plane; otherwise it IS the plane and
    // every loop bound / divisor below is a compile-time constant
    static constexpr bool PADDED = PADDED_;
    static constexpr int NPX = WY * WX;
    static constexpr int P = WX + 1;          // plane pitch in float2 (odd -> conflict-free row & column walks)
    static constexpr int PLANE = WY * P;      // float2 per plane
    static constexpr int NWARP = NT / 32;
    static constexpr int R1X = Factor<WX>::R1, R2X = Factor<WX>::R2;
    static constexpr int R1Y = Factor<WY>::R1, R2Y = Factor<WY>::R2;
    static constexpr int NRED = 8;            // reduction slots per warp
};

struct Params;

// Shared-memory image of one CTA.  (On the host emulator this is a plain heap object.)
template <class C>
struct Smem {
    float2 plane[C::NWIN][C::PLANE];   // z = a + i b per window; plane[0] later holds G and the result planes
    float2 twx[C::WX];                 // exp(-2 pi i j / WX)
    float2 twy[C::WY];                 // exp(-2 pi i j / WY)
    unsigned long long red[C::NWARP][C::NRED];  // cross-warp reduction scratch
    float scale[2];                    // 1 / (N^2 std_a std_b) per window (0 when a std is 0)
    float mean[4];                     // mean of a0, b0, a1, b1
    float isum[4];                     // u8 input: exact integer pixel sums of a0, b0, a1, b1 (as float, < 2^24)
    float stat[2][4];                  // per window: corr_max, sum, peak index (as float bits), unused
};

// Fused result gather over peer memory (multi-GPU): when n > 0 every window's four results are ALSO stored into each peer's
// gather buffer [4][pairs_total][n_windows] (peer-mapped device memory, e.g. torch symmetric memory over NVLink) at this
// rank's pair offset - the all-gather of pyorc_b200.parallel happens as P2P stores from the kernel epilogue.
struct PeerOut {
    int n;
    float* base[8];
    long long field;     // pairs_total * n_windows: distance between the four fields of a gather buffer
    long long pair0;     // global index of this launch's first pair
};
B2_HD void peer_store(const PeerOut& po, long long local_pair, long long n_windows, int widx, float uu, float vv, float oc, float os) {
    const long long go = (po.pair0 + local_pair) * n_windows + widx;
    for (int r = 0; r < po.n; ++r) {
        float* b = po.base[r];
        b[go] = uu; b[go + po.field] = vv; b[go + 2 * po.field] = oc; b[go + 3 * po.field] = os;
    }
}

// Per-launch parameters (plain data, passed by value).
struct Params {
    const void* frames;       // [n_frames][H][pitch] u8 or f32
    long long frame_stride;   // bytes between frames
    int pitch;                // bytes between rows
    int is_f32;               // 0: u8, 1: f32
    int n_rows, n_cols;       // PIV field shape
    int sy, sx;               // window stride (w - overlap) in pixels
    int n_pairs;              // frame pairs in this launch
    int ny, nx;               // true window size; equal to the FFT plane (WY, WX) or, for sizes that are not a power
                              // of two, at most half of it ("padded mode", see phase_embed)
    const short* shift;       // optional [n_pairs][n_windows][2] = (dy, dx): whole-pixel displacement of frame k+1's window
                              // (second pass of the two-pass scheme, multipass.cuh); the result is shift + residual
    int clip_norm;            // 1: clip normalised windows at 0 (OpenPIV normalize_intensity)
    int border_nan;           // 1: border peak -> NaN displacement, 0: integer peak
    float gauss_eps;          // epsilon added before logs
    const unsigned char* keep;  // optional [n_windows] keep mask (signal_threshold); nullptr = keep all
    float* u; float* v; float* cmax; float* s2n;   // [n_pairs][n_rows*n_cols]
    float* planes;            // optional debug dump [n_pairs][n_windows][WY][WX] (fftshifted, clipped), or nullptr
    const float* fshift;      // optional [n_out_pairs][n_windows][2] = (dv, du) float predictor added to (v, u) (deformation pass)
    int pair_step;            // 0 / 1: consecutive frame pairs (k, k+1); 2: the frames are an interleaved stack (a_0, b_0, a_1, b_1 ...) and
                              // only the pairs (2k, 2k+1) are computed, result index k (deformation pass, multipass.cuh)
    float* scratch;           // large-window direct kernel (k_direct.cu): one correlation plane per CTA, [grid][wy * wx]
    PeerOut peer;             // optional fused gather (n = 0: off)
};

template <class C> B2_HD int win_ny(const Params& p) { return C::PADDED ? p.ny : C::WY; }
template <class C> B2_HD int win_nx(const Params& p) { return C::PADDED ? p.nx : C::WX; }

// ------------------------------------------------------------------------------------------------------------
// Block reduction helpers.  deposit(): each thread contributes v to slot; after a barrier total() sums warps.
// Device: warp shuffle then lane 0 stores.  Host emulator: sequential accumulate into the warp's cell.
// ------------------------------------------------------------------------------------------------------------
template <class C>
B2_HD void red_zero(Smem<C>& s, int tid) {
    for (int i = tid; i < C::NWARP * C::NRED; i += C::NT) s.red[i / C::NRED][i % C::NRED] = 0ull;
}
template <class C>
B2_HD void deposit_sum_u64(Smem<C>& s, int tid, int slot, unsigned long long v) {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) s.red[tid >> 5][slot] = v;
#else
    s.red[tid >> 5][slot] += v;
#endif
}
template <class C>
B2_HD void deposit_max_u64(Smem<C>& s, int tid, int slot, unsigned long long v) {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    if ((tid & 31) == 0) s.red[tid >> 5][slot] = v;
#else
    if (v > s.red[tid >> 5][slot]) s.red[tid >> 5][slot] = v;
#endif
}
template <class C>
B2_HD void deposit_sum_f32(Smem<C>& s, int tid, int slot, float v) {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) s.red[tid >> 5][slot] = (unsigned long long)__float_as_uint(v);
#else
    union { float f; unsigned u; } a, b;
    a.u = (unsigned)s.red[tid >> 5][slot];
    b.f = a.f + v;
    s.red[tid >> 5][slot] = (unsigned long long)b.u;
#endif
}
template <class C>
B2_HD unsigned long long total_sum_u64(const Smem<C>& s, int slot) {
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < C::NWARP; ++w) t += s.red[w][slot];
    return t;
}
template <class C>
B2_HD unsigned long long total_max_u64(const Smem<C>& s, int slot) {
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < C::NWARP; ++w) t = s.red[w][slot] > t ? s.red[w][slot] : t;
    return t;
}
template <class C>
B2_HD float total_sum_f32(const Smem<C>& s, int slot) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < C::NWARP; ++w) {
        union { float f; unsigned u; } a;
        a.u = (unsigned)s.red[w][slot];
        t += a.f;
    }
    return t;
}

// Work-item decoding: item -> (pair, window indices).  Items enumerate window PAIRS (2j, 2j+1) per frame pair;
// an odd trailing window is paired with itself (its Im result is discarded).
struct Item {
    int pair;      // frame pair index
    int w[2];      // flattened window index (r*n_cols + c) for slot 0 / 1
    int valid1;    // slot 1 is a real (distinct) window
};
template <class C>
B2_HD Item decode_item(const Params& p, int item) {
    const int nw = p.n_rows * p.n_cols;
    const int per_pair = (C::NWIN == 2) ? (nw + 1) / 2 : nw;
    Item it;
    it.pair = item / per_pair;
    const int j = item % per_pair;
    if (C::NWIN == 2) {
        it.w[0] = 2 * j;
        it.w[1] = (2 * j + 1 < nw) ? 2 * j + 1 : 2 * j;
        it.valid1 = (2 * j + 1 < nw);
    } else {
        it.w[0] = it.w[1] = j;
        it.valid1 = 0;
    }
    return it;
}

// ------------------------------------------------------------------------------------------------------------
// PHASE 0: twiddle tables (once per CTA) + zero reduction scratch
// ------------------------------------------------------------------------------------------------------------
template <class C>
B2_HD void phase_init(Smem<C>& s, int tid, const float2* twx_g, const float2* twy_g) {
    for (int i = tid; i < C::WX; i += C::NT) s.twx[i] = twx_g[i];
    for (int i = tid; i < C::WY; i += C::NT) s.twy[i] = twy_g[i];
}

// ------------------------------------------------------------------------------------------------------------
// PHASE 1: load the windows (frame k -> .x, frame k+1 -> .y) and accumulate sums.
//   u8 input: exact integer sum / sum of squares.  f32 input: float sum here, centred second moment in phase 2.
//   Lanes walk along x (coalesced 1-byte / 4-byte loads; the 50 % overlap re-reads hit L2).
// ------------------------------------------------------------------------------------------------------------
template <class C>
B2_HD void phase_load(Smem<C>& s, int tid, const Params& p, const Item& it) {
    const unsigned char* base = (const unsigned char*)p.frames + (long long)it.pair * p.frame_stride;
#pragma unroll
    for (int w = 0; w < C::NWIN; ++w) {
        const int r = it.w[w] / p.n_cols, c = it.w[w] % p.n_cols;
        const long long off = (long long)(r * p.sy) * p.pitch;
        const int x0 = c * p.sx;
        // frame k+1's window, optionally displaced by whole pixels (the caller keeps it inside the frame)
        long long boff = p.frame_stride;
        if (p.shift) {
            const short* sh = p.shift + 2 * ((long long)it.pair * p.n_rows * p.n_cols + it.w[w]);
            boff += (long long)sh[0] * p.pitch + (long long)sh[1] * (p.is_f32 ? 4 : 1);
        }
        unsigned long long sa = 0, sb = 0, qa = 0, qb = 0;
        float fa = 0.f, fb = 0.f;
        for (int e = tid; e < win_ny<C>(p) * win_nx<C>(p); e += C::NT) {
            const int y = e / win_nx<C>(p), x = e % win_nx<C>(p);
            float a, b;
            if (!p.is_f32) {
                const unsigned char* ra = base + off + (long long)y * p.pitch + x0 + x;
                const unsigned ua = ra[0], ub = ra[boff];
                sa += ua; sb += ub; qa += ua * ua; qb += ub * ub;
                a = (float)ua; b = (float)ub;
            } else {
                const float* ra = (const float*)(base + off + (long long)y * p.pitch) + x0 + x;
                a = ra[0];
                b = *(const float*)((const unsigned char*)ra + boff);
                fa += a; fb += b;
            }
            s.plane[w][y * C::P + x] = make_float2(a, b);
        }
        if (!p.is_f32) {
            // pack sum (<= 2^22+) and sum of squares (<= 2^30+) of a and b into separate slots
            deposit_sum_u64<C>(s, tid, 4 * w + 0, sa);
            deposit_sum_u64<C>(s, tid, 4 * w + 1, qa);
            deposit_sum_u64<C>(s, tid, 4 * w + 2, sb);
            deposit_sum_u64<C>(s, tid, 4 * w + 3, qb);
        } else {
            deposit_sum_f32<C>(s, tid, 4 * w + 0, fa);
            deposit_sum_f32<C>(s, tid, 4 * w + 2, fb);
        }
    }
}

// PHASE 2 (u8): means + scale from exact integer moments.  (f32: means only, then phase 2b/2c.)
template <class C>
B2_HD void phase_stats(Smem<C>& s, int tid, const Params& p) {
    if (tid < C::NWIN) {
        const int w = tid;
        if (!p.is_f32) {
            const double n = (double)(win_ny<C>(p) * win_nx<C>(p));
            const double sa = (double)total_sum_u64<C>(s, 4 * w + 0), qa = (double)total_sum_u64<C>(s, 4 * w + 1);
            const double sb = (double)total_sum_u64<C>(s, 4 * w + 2), qb = (double)total_sum_u64<C>(s, 4 * w + 3);
            const double va = (qa - sa * sa / n) / n, vb = (qb - sb * sb / n) / n;  // population variance
            s.mean[2 * w + 0] = (float)(sa / n);
            s.mean[2 * w + 1] = (float)(sb / n);
            s.isum[2 * w + 0] = (float)sa;
            s.isum[2 * w + 1] = (float)sb;
            // unnormalised inverse FFT of size NPX returns NPX * sum(a b); the reference divides the sum by n
            s.scale[w] = (va > 0.0 && vb > 0.0) ? (float)(1.0 / ((double)C::NPX * n * sqrt(va) * sqrt(vb))) : 0.f;
        } else {
            s.mean[2 * w + 0] = total_sum_f32<C>(s, 4 * w + 0) / (float)(win_ny<C>(p) * win_nx<C>(p));
            s.mean[2 * w + 1] = total_sum_f32<C>(s, 4 * w + 2) / (float)(win_ny<C>(p) * win_nx<C>(p));
        }
    }
}

// PHASE 3: centre (and optionally clip at 0) in place.  f32 input also accumulates the centred 2nd moment.
template <class C>
B2_HD void phase_center(Smem<C>& s, int tid, const Params& p) {
#pragma unroll
    for (int w = 0; w < C::NWIN; ++w) {
        const float ma = s.mean[2 * w], mb = s.mean[2 * w + 1];
        const float ia = s.isum[2 * w], ib = s.isum[2 * w + 1];
        float qa = 0.f, qb = 0.f;
        const bool exact_n = !C::PADDED;
        for (int e = tid; e < win_ny<C>(p) * win_nx<C>(p); e += C::NT) {
            const int y = e / win_nx<C>(p), x = e % win_nx<C>(p);
            float2 z = s.plane[w][y * C::P + x];
            if (!p.is_f32 && exact_n) {  // exact: (N*x - S) is an integer below 2^24, 1/N is a power of two
                z.x = (z.x * (float)C::NPX - ia) * (1.0f / (float)C::NPX);
                z.y = (z.y * (float)C::NPX - ib) * (1.0f / (float)C::NPX);
            } else { z.x -= ma; z.y -= mb; }
            qa += z.x * z.x; qb += z.y * z.y;
            if (p.clip_norm) { z.x = z.x < 0.f ? 0.f : z.x; z.y = z.y < 0.f ? 0.f : z.y; }
            s.plane[w][y * C::P + x] = z;
        }
        if (p.is_f32) {
            deposit_sum_f32<C>(s, tid, 4 * w + 1, qa);
            deposit_sum_f32<C>(s, tid, 4 * w + 3, qb);
        }
    }
}
// PHASE 3c (padded mode only, ny*nx < NPX): windows whose size is not a power of two run through the power-of-two
// FFT as an EXACT circular correlation of period (ny, nx): frame k's window is zero-padded, frame k+1's window is
// tiled periodically over [0, 2ny) x [0, 2nx) (zero beyond), so  sum_x a'(x) b'(x + s)  for 0 <= s < n never wraps in
// the (>= 2n)-point plane.  Elements inside the window keep their value and are only READ here (by the threads that fill
// their tiled copies), elements outside are only written: no barrier needed inside the phase.
template <class C>
B2_HD void phase_embed(Smem<C>& s, int tid, const Params& p) {
    if (C::PADDED) {
#pragma unroll
        for (int w = 0; w < C::NWIN; ++w) {
            for (int e = tid; e < C::NPX; e += C::NT) {
                const int y = e / C::WX, x = e % C::WX;
                if (y < p.ny && x < p.nx) continue;
                const float b = (y < 2 * p.ny && x < 2 * p.nx) ? s.plane[w][(y % p.ny) * C::P + (x % p.nx)].y : 0.f;
                s.plane[w][y * C::P + x] = make_float2(0.f, b);
            }
        }
    }
}

// PHASE 3b (f32 only): scale from centred second moments.
template <class C>
B2_HD void phase_stats_f32(Smem<C>& s, int tid, const Params& p) {
    if (p.is_f32 && tid < C::NWIN) {
        const int w = tid;
        const double n = (double)(win_ny<C>(p) * win_nx<C>(p));
        const double va = (double)total_sum_f32<C>(s, 4 * w + 1) / n, vb = (double)total_sum_f32<C>(s, 4 * w + 3) / n;
        s.scale[w] = (va > 0.0 && vb > 0.0) ? (float)(1.0 / ((double)C::NPX * n * sqrt(va) * sqrt(vb))) : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------------------
// FFT passes.  "line" = a row (ES = 1) or a column (ES = P).  Row passes put lanes along y, column passes along
// x: with the odd pitch both walks are bank-conflict free.
//   fwd_strided : for sub = n2: load (R2*n1+n2), DFT_R1 -> k1, * w_N^(n2*k1), store at (R2*k1+n2)
//   fwd_contig  : for sub = k1: load (R2*k1+n2), DFT_R2 -> k2, store at (R2*k1+k2)
//   inv_contig  : for sub = k1: load (R2*k1+k2), IDFT_R2 -> n2, * conj w_N^(n2*k1), store at (R2*k1+n2)
//   inv_strided : for sub = n2: load (R2*k1+n2), IDFT_R1 -> n1, store at (R2*n1+n2)
// ------------------------------------------------------------------------------------------------------------
template <int N, int ES, int INV>
B2_HD void line_strided(float2* line, const float2* tw, int n2) {
    constexpr int R1 = Factor<N>::R1, R2 = Factor<N>::R2;
    float2 v[R1];
#pragma unroll
    for (int i = 0; i < R1; ++i) v[i] = line[(R2 * i + n2) * ES];
    RegDFT<R1, INV>::run(v);
    if (INV == 0) {
#pragma unroll
        for (int k1 = 1; k1 < R1; ++k1) v[k1] = cmulw<0>(v[k1], tw[n2 * k1]);
    }
#pragma unroll
    for (int i = 0; i < R1; ++i) line[(R2 * i + n2) * ES] = v[i];
}
template <int N, int ES, int INV>
B2_HD void line_contig(float2* line, const float2* tw, int k1) {
    constexpr int R2 = Factor<N>::R2;
    float2 v[R2];
#pragma unroll
    for (int i = 0; i < R2; ++i) v[i] = line[(R2 * k1 + i) * ES];
    RegDFT<R2, INV>::run(v);
    if (INV == 1) {
#pragma unroll
        for (int n2 = 1; n2 < R2; ++n2) v[n2] = cmulw<1>(v[n2], tw[n2 * k1]);
    }
#pragma unroll
    for (int i = 0; i < R2; ++i) line[(R2 * k1 + i) * ES] = v[i];
}

// One pass over NPL planes.  DIM: 0 = along x (rows), 1 = along y (columns).  KIND: 0 strided, 1 contiguous.
template <class C, int NPL, int DIM, int KIND, int INV>
B2_HD void fft_pass(Smem<C>& s, int tid) {
    constexpr int N = DIM == 0 ? C::WX : C::WY;            // transform length
    constexpr int L = DIM == 0 ? C::WY : C::WX;            // number of lines
    constexpr int R1 = Factor<N>::R1, R2 = Factor<N>::R2;
    constexpr int NSUB = KIND == 0 ? R2 : R1;
    constexpr int ES = DIM == 0 ? 1 : C::P;
    const float2* tw = DIM == 0 ? s.twx : s.twy;
    for (int t = tid; t < NPL * L * NSUB; t += C::NT) {
        const int pl = t / (L * NSUB);
        const int r = t % (L * NSUB);
        const int line = r % L, sub = r / L;               // lanes along the line index
        float2* base = s.plane[pl] + (DIM == 0 ? line * C::P : line);
        if (KIND == 0) line_strided<N, ES, INV>(base, tw, sub);
        else           line_contig<N, ES, INV>(base, tw, sub);
    }
}

// ------------------------------------------------------------------------------------------------------------
// PHASE X: cross spectra.  For every frequency pair {k, -k} (positions in digit-swapped order):
//   A = (z + conj(zn))/2, B = (z - conj(zn))/(2i),  R = conj(A) B * scale,  R(-k) = conj(R(k))
//   G(k) = R0 + i R1 ,  G(-k) = conj(R0) + i conj(R1)       -> written to plane[0] in place.
// ------------------------------------------------------------------------------------------------------------
B2_HD float2 cross_spec(float2 z, float2 zn, float sc) {
    const float ax = 0.5f * (z.x + zn.x), ay = 0.5f * (z.y - zn.y);
    const float bx = 0.5f * (z.y + zn.y), by = -0.5f * (z.x - zn.x);
    return make_float2((ax * bx + ay * by) * sc, (ax * by - ay * bx) * sc);
}

template <class C>
B2_HD void phase_cross(Smem<C>& s, int tid) {
    constexpr int WY = C::WY, WX = C::WX;
    constexpr int HALF = C::R2Y / 2;  // k2y < HALF <=> 0 <= ky < WY/2
    const float sc0 = s.scale[0], sc1 = (C::NWIN == 2) ? s.scale[1] : 0.f;
    for (int t = tid; t < (WY / 2 + 1) * WX; t += C::NT) {
        const int slot = t / WX, px = t % WX;
        int py;
        if (slot < WY / 2) py = (slot / HALF) * C::R2Y + (slot % HALF);
        else               py = HALF;  // k1y = 0, k2y = R2Y/2  -> ky = WY/2
        const int ky = freq_of_pos<WY>(py), kx = freq_of_pos<WX>(px);
        const bool selfrow = (ky == 0) || (ky == WY / 2);
        if (selfrow && kx > WX / 2) continue;
        const int pyn = negpos<WY>(py), pxn = negpos<WX>(px);
        const int i = py * C::P + px, in = pyn * C::P + pxn;
        const float2 r0 = cross_spec(s.plane[0][i], s.plane[0][in], sc0);
        float2 r1 = make_float2(0.f, 0.f);
        if (C::NWIN == 2) r1 = cross_spec(s.plane[1][i], s.plane[1][in], sc1);
        s.plane[0][i] = make_float2(r0.x - r1.y, r0.y + r1.x);
        if (in != i) s.plane[0][in] = make_float2(r0.x + r1.y, -r0.y + r1.x);
    }
}

// ------------------------------------------------------------------------------------------------------------
// PHASE R: reductions over the fftshifted, clipped planes: max (first occurrence) and sum.
//   key = (value bits << 32) | (~index)  : values are in [0,1] so their bit patterns order like the floats.
// ------------------------------------------------------------------------------------------------------------
B2_HD float clip01(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }

template <class C>
B2_HD float shifted_value(const Smem<C>& s, int w, int i, int j, int ny = C::WY, int nx = C::WX) {  // fftshifted coordinates
    const int sy = (i + ny - ny / 2) % ny, sx = (j + nx - nx / 2) % nx;
    const float2 z = s.plane[0][sy * C::P + sx];
    return clip01(w == 0 ? z.x : z.y);
}

template <class C>
B2_HD void phase_reduce(Smem<C>& s, int tid, const Params& p, const Item& it) {
#pragma unroll
    for (int w = 0; w < C::NWIN; ++w) {
        unsigned long long best = 0ull;
        float sum = 0.f;
        // a window with zero variance in either frame has an exactly-zero plane in the reference; the packed
        // inverse FFT would otherwise leave ~1e-10 rounding cross-talk from its partner window there
        const bool dead = (s.scale[w] == 0.f);
        const int ny = win_ny<C>(p), nx = win_nx<C>(p);
        for (int e = tid; e < ny * nx; e += C::NT) {
            const int i = e / nx, j = e % nx;
            const float v = dead ? 0.f : shifted_value<C>(s, w, i, j, ny, nx);
            sum += v;
            union { float f; unsigned u; } cv; cv.f = v;
            const unsigned long long key = ((unsigned long long)cv.u << 32) | (unsigned long long)(0xffffffffu - (unsigned)e);
            best = key > best ? key : best;
            if (p.planes && (w == 0 || it.valid1)) {
                const long long nw = (long long)p.n_rows * p.n_cols;
                p.planes[(((long long)it.pair * nw + it.w[w]) * ny + i) * nx + j] = v;
            }
        }
        deposit_max_u64<C>(s, tid, 2 * w + 0, best);
        deposit_sum_f32<C>(s, tid, 2 * w + 1, sum);
    }
}

// ------------------------------------------------------------------------------------------------------------
// PHASE P: sub-pixel Gaussian peak + outputs (one thread per window).
// Mirrors ffpiv.u_v_displacement + pyorc/velocimetry/ffpiv.py:465-466.
// ------------------------------------------------------------------------------------------------------------
template <class C>
B2_HD void phase_peak(Smem<C>& s, int tid, const Params& p, const Item& it) {
    if (tid >= C::NWIN) return;
    const int w = tid;
    if (w == 1 && !it.valid1) return;
    const unsigned long long key = total_max_u64<C>(s, 2 * w + 0);
    const float sum = total_sum_f32<C>(s, 2 * w + 1);
    union { float f; unsigned u; } cv; cv.u = (unsigned)(key >> 32);
    const float cmax = cv.f;
    const int idx = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
    const int ny = win_ny<C>(p), nx = win_nx<C>(p);
    const int pi = idx / nx, pj = idx % nx;
    const float mean = sum / (float)(ny * nx);
    float uu, vv;
    const bool border = (pi == 0) || (pi == ny - 1) || (pj == 0) || (pj == nx - 1);
    if (border) {
        if (p.border_nan) { uu = nanf(""); vv = nanf(""); }
        else { uu = (float)(pj - nx / 2); vv = (float)(pi - ny / 2); }
    } else {
        const float eps = p.gauss_eps;
        const float lc = logf(cmax + eps);
        const float ll = logf(shifted_value<C>(s, w, pi - 1, pj, ny, nx) + eps), lr = logf(shifted_value<C>(s, w, pi + 1, pj, ny, nx) + eps);
        const float ld = logf(shifted_value<C>(s, w, pi, pj - 1, ny, nx) + eps), lu = logf(shifted_value<C>(s, w, pi, pj + 1, ny, nx) + eps);
        const float di = (ll - lr) / (2.f * ll - 4.f * lc + 2.f * lr);
        const float dj = (ld - lu) / (2.f * ld - 4.f * lc + 2.f * lu);
        vv = ((float)pi + di) - (float)(ny / 2);
        uu = ((float)pj + dj) - (float)(nx / 2);
    }
    float o_c = cmax, o_s = cmax / mean;
    if (p.keep && !p.keep[it.w[w]]) { uu = vv = o_c = o_s = nanf(""); }
    const long long o = (long long)it.pair * p.n_rows * p.n_cols + it.w[w];
    if (p.shift) { vv += (float)p.shift[2 * o]; uu += (float)p.shift[2 * o + 1]; }
    p.u[o] = uu; p.v[o] = vv; p.cmax[o] = o_c; p.s2n[o] = o_s;
    if (p.peer.n) peer_store(p.peer, it.pair, (long long)p.n_rows * p.n_cols, it.w[w], uu, vv, o_c, o_s);
}

}  // namespace b2piv

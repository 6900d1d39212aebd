// piv_rows128.cuh - 128x128 windows on the row-per-thread machinery of piv_rows.cuh (sm_100a, uint8 frames).
//
// A thread cannot hold a 128-point line in registers, so the 128 x 128 window is split into its four POLYPHASE
// components a_p[m1, m2] = a[2 m1 + p1, 2 m2 + p2], p = (p1, p2) in {0,1}^2: four 64 x 64 images.  The circular
// cross-correlation c[n] = sum_m a[m] b[m + n] (period 128) of the reference,
//     irfft2(conj(rfft2 a) * rfft2 b)                                  (ffpiv.cross_corr, pyorc/velocimetry/ffpiv.py:450-459)
// has the polyphase components (n = 2 n' + q, p + q = 2 s + r per axis, i.e. r = p xor q, s = p and q)
//     c_q[n'] = sum_p corr64(a_p, b_r)[n' + s]
// and therefore, in the 64 x 64 frequency domain,
//     C_q[k] = sum_p conj(A_p[k]) B_{p xor q}[k] * exp(+2 pi i (k1 s1 + k2 s2) / 64).
// No 128-point transform is needed at all: FOUR sub-groups of 64 threads (one per component) each run the unchanged
// 64 x 64 pipeline of piv_rows.cuh - TMA tile -> rows in registers -> FFT -> transpose -> FFT -> Hermitian separation of the
// two packed windows - and only the cross-spectrum step couples them: sub-group q gathers the parked spectra A_p of the
// previous frame and the new spectra B_r of all four components (through shared memory, in batches of 11 ky rows), forms
// C_q with the two phase factors, and inverse-transforms it with the unchanged code.  The result rows a thread ends up
// with are rows of the polyphase component c_q, i.e. every other element of every other row of the reference's plane;
// max / sum / first-occurrence argmax / the three rows around the peak only need the index map
//     reference (fftshifted) row = (2 m1 + q1 + 64) % 128,  column = (2 m2 + q2 + 64) % 128.
// Work per window pair: 1.0 complex 64x64 FFT per component and frame, like the native sizes (the shared-memory kernel
// this replaces for 128 x 128 does 2.0 complex 128x128 FFTs per window and pair).
#pragma once
#include "piv_rows.cuh"

namespace b2piv {

using R6 = RCfg<64, 33>;   // odd transpose pitch: measured 0.7 % faster here (piv_rows.cuh, RCfg::BP)

struct R128Smem {
    RSmem<R6> sub[4];              // per polyphase component: transpose blocks (+ TMA tile / exchange aliases) and parked spectra
    float nb[2][3][128];           // the three plane rows around each peak (reference order)
    unsigned long long mbar;       // TMA completion barrier of the whole group
};
constexpr int R128_BATCH = 11;     // ky rows per exchange batch (3 batches cover ky = 0 .. 32)
// Padded mode (PAD = true): even windows of 34 .. 64 px per side (the larger one; the other any even size).  The period-n
// circular correlation of the reference needs a 2n-point plane (piv_rows.cuh, "Padded mode": window `a` zero-padded, window `b`
// tiled 2 x 2 through the spectrum factor T), i.e. the 128 x 128 plane of this kernel: each polyphase component holds
// ny/2 x nx/2 samples of the window, the tiling shift n = 2 (n/2) stays inside a component, so T(k) = (1 + e^{-2 pi i k1 (ny/2)/64})
// (1 + e^{-2 pi i k2 (nx/2)/64}) multiplies every C_q alike.  RParams::ny / nx hold the COMPONENT size (ny/2, nx/2).
// Tile: one 80-byte x 64-row unswizzled box per window from the 16-byte boundary below its start, in sub[0]'s transpose blocks.
constexpr int R128_PWB = 80;                    // padded tile: bytes per row
constexpr int R128_PWIN = 64 * R128_PWB;        // and per window
static_assert(sizeof(float2) * R6::NWARP * R6::XBLK >= 2 * R128_PWIN, "padded tile must fit in a sub-group's transpose blocks");
// Padded mode on float32 frames: one 68-float x 64-row unswizzled box per window from the 16-byte boundary below its start (the
// window's 64 floats at a float offset of 0 .. 3), window w in the spectrum block of sub-group w (idle between the cross phases,
// like the native float32 tile).
constexpr int R128_PFW = 68;                    // padded float32 tile: floats per row
constexpr int R128_PFWIN = 64 * R128_PFW * 4;   // bytes per window
constexpr int R128_NPX = 128 * 128;
// tile: window w rows [32 j, 32 j + 32) -> sub[j].tile() + w * 4096 ([32 rows][128 B], SWIZZLE_128B)
static_assert(sizeof(float2) * R6::NWARP * R6::XBLK >= 2 * 4096, "tile quarter must fit in a sub-group's transpose blocks");
static_assert(sizeof(float2) * R6::NWARP * R6::XBLK >= sizeof(float4) * R128_BATCH * 64, "a batch of returned cross spectra must fit in a sub-group's transpose blocks");
constexpr int R128_TM_COLS = 4 * 9 * 4;   // Tensor-Memory columns per thread: 4 components x 9 own bins x (A0, A1)
static_assert(sizeof(RSmem<R6>::park) == sizeof(float4) * 33 * 64, "parked spectra are [33][64] float4");

#ifdef __CUDACC__
// P1: row 2 sigma(t) + p1 of both windows from the tile, bytes of column parity p2 packed (64 bytes per window), exact
// integer moments of the sub-image -> sub[own].red[warp][0..3]
__device__ __forceinline__ void r128_p1(R128Smem& s, RRegs<R6>& r, int sub, int t) {
    const int p1 = sub >> 1, p2 = sub & 1;
    const int row = 2 * column_of<64>(t) + p1;
    const int j = row >> 5, rr = row & 31;
    const unsigned sel = p2 ? 0x7531u : 0x6420u;
    unsigned S[2] = {0, 0}, Q[2] = {0, 0};
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        const unsigned char* base = s.sub[j].tile() + w * 4096 + rr * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 q = *reinterpret_cast<const uint4*>(base + ((c ^ (rr & 7)) << 4));
            r.px[w][2 * c + 0] = __byte_perm(q.x, q.y, sel);
            r.px[w][2 * c + 1] = __byte_perm(q.z, q.w, sel);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            S[w] = __dp4a(r.px[w][k], 0x01010101u, S[w]);
            Q[w] = __dp4a(r.px[w][k], r.px[w][k], Q[w]);
        }
    }
    unsigned vals[4] = {S[0], Q[0], S[1], Q[1]};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vals[k] += __shfl_xor_sync(0xffffffffu, vals[k], o);
    }
    if ((t & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) s.sub[sub].red[t >> 5][k] = vals[k];
    }
}

// P2: moments of the whole 128 x 128 windows (all four components) -> mean, 0.5 / std; convert + centre
__device__ __forceinline__ void r128_p2(R128Smem& s, RRegs<R6>& r, int clip_norm) {
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        unsigned long long S = 0, Q = 0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int k = 0; k < 2; ++k) { S += s.sub[g].red[k][2 * w]; Q += s.sub[g].red[k][2 * w + 1]; }
        }
        const unsigned long long m2 = (unsigned long long)R128_NPX * Q - S * S;   // N^2 * variance, exact
        r.mean_new[w] = (float)S * (1.0f / (float)R128_NPX);
        r.half_alpha_new[w] = m2 ? 0.5f * (float)R128_NPX * (1.0f / sqrtf((float)m2)) : 0.f;
    }
    r.dc_fix[0] = r.dc_fix[1] = 0.f;
    if (!clip_norm) {
        // magic-number conversion + centring with the mean rounded to 1/256 (rows_p2_pre in piv_rows.cuh): the offset
        // delta = mean - mq shifts the DC bin of EVERY polyphase component by 4096 delta, removed in r128_cross
        const float c0 = __fadd_rn(32768.0f, r.mean_new[0]), c1 = __fadd_rn(32768.0f, r.mean_new[1]);
        r.dc_fix[0] = (r.mean_new[0] - (c0 - 32768.0f)) * 4096.0f;
        r.dc_fix[1] = (r.mean_new[1] - (c1 - 32768.0f)) * 4096.0f;
        const float2 c = make_float2(c0, c1);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float2 mg = make_float2(__uint_as_float(__byte_perm(r.px[0][k], 0x47000000u, 0x7404u | (b << 4))),
                                              __uint_as_float(__byte_perm(r.px[1][k], 0x47000000u, 0x7404u | (b << 4))));
                r.v[4 * k + b] = pk_sub(mg, c);
            }
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float2 a = pk_sub(make_float2(byte_to_float(r.px[0][k], b), byte_to_float(r.px[1][k], b)), make_float2(r.mean_new[0], r.mean_new[1]));
            if (clip_norm) a = make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f));
            r.v[4 * k + b] = a;
        }
    }
}

// P1 (padded): row 2 sigma(t) + p1 of both windows at their byte offset in the unswizzled tile, masked to the window, bytes of
// column parity p2 packed (at most 32 per window) -> r.px[w][0..7]; exact integer moments of the sub-image
__device__ __forceinline__ void r128_p1_pad(R128Smem& s, RRegs<R6>& r, int sub, int t, const RParams& p, int xoff0, int xoff1) {
    const int p1 = sub >> 1, p2 = sub & 1;
    const int row = 2 * column_of<64>(t) + p1;
    const unsigned rowmask = row < 2 * p.ny ? 0xffffffffu : 0u;
    const int rr = row < 64 ? row : 0;           // rows past the tile are masked anyway
    const unsigned sel = p2 ? 0x7531u : 0x6420u;
    unsigned S[2] = {0, 0}, Q[2] = {0, 0};
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        const int xoff = w == 0 ? xoff0 : xoff1;
        const unsigned char* base = s.sub[0].tile() + w * R128_PWIN + rr * R128_PWB + (xoff & ~3);
        const int sh = (xoff & 3) * 8;
        unsigned wd[17];
#pragma unroll
        for (int k = 0; k <= 16; ++k) wd[k] = *reinterpret_cast<const unsigned*>(base + 4 * k);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned a = funnel_r(wd[2 * k], wd[2 * k + 1], sh) & p.pad_mask[2 * k] & rowmask;
            const unsigned b = funnel_r(wd[2 * k + 1], wd[2 * k + 2], sh) & p.pad_mask[2 * k + 1] & rowmask;
            r.px[w][k] = __byte_perm(a, b, sel);
            r.px[w][8 + k] = 0u;
            S[w] = __dp4a(r.px[w][k], 0x01010101u, S[w]);
            Q[w] = __dp4a(r.px[w][k], r.px[w][k], Q[w]);
        }
    }
    unsigned vals[4] = {S[0], Q[0], S[1], Q[1]};
#pragma unroll
    for (int k = 0; k < 4; ++k) vals[k] = __reduce_add_sync(0xffffffffu, vals[k]);
    if ((t & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) s.sub[sub].red[t >> 5][k] = vals[k];
    }
}

// P2 (padded): moments over the ny * nx window pixels (all four components); pixels outside the window stay exactly 0
__device__ __forceinline__ void r128_p2_pad(R128Smem& s, RRegs<R6>& r, int t, const RParams& p) {
    const unsigned long long npx = 4ull * (unsigned long long)(p.ny * p.nx);
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        unsigned long long S = 0, Q = 0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int k = 0; k < 2; ++k) { S += s.sub[g].red[k][2 * w]; Q += s.sub[g].red[k][2 * w + 1]; }
        }
        const unsigned long long m2 = npx * Q - S * S;
        r.mean_new[w] = (float)S / (float)npx;
        r.half_alpha_new[w] = m2 ? 0.5f * (float)npx * (1.0f / sqrtf((float)m2)) : 0.f;
    }
    const bool rowok = column_of<64>(t) < p.ny;
    r.dc_fix[0] = r.dc_fix[1] = 0.f;
    if (!p.clip_norm) {
        // magic-number conversion (rows_p2_pre_pad): byte - cm * mq with the mean rounded to 1/256; a component holds ny * nx
        // samples (component size), so its DC bin is off by ny * nx * (mean - mq), removed in r128_cross
        const float c0 = __fadd_rn(32768.0f, r.mean_new[0]), c1 = __fadd_rn(32768.0f, r.mean_new[1]);
        const float mq0 = c0 - 32768.0f, mq1 = c1 - 32768.0f;
        r.dc_fix[0] = (r.mean_new[0] - mq0) * (float)(p.ny * p.nx);
        r.dc_fix[1] = (r.mean_new[1] - mq1) * (float)(p.ny * p.nx);
        const float2 nm = rowok ? make_float2(-mq0, -mq1) : make_float2(0.f, 0.f);
        const float2 base = make_float2(32768.0f, 32768.0f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float2 mg = make_float2(__uint_as_float(__byte_perm(r.px[0][k], 0x47000000u, 0x7404u | (b << 4))),
                                              __uint_as_float(__byte_perm(r.px[1][k], 0x47000000u, 0x7404u | (b << 4))));
                r.v[4 * k + b] = pk_fma(make_float2(p.pad_cm[4 * k + b], p.pad_cm[4 * k + b]), nm, pk_sub(mg, base));
            }
        }
    } else {
        const float nm0 = rowok ? -r.mean_new[0] : 0.f, nm1 = rowok ? -r.mean_new[1] : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const float a0 = fmaf(p.pad_cm[4 * k + b], nm0, byte_to_float(r.px[0][k], b));
                const float a1 = fmaf(p.pad_cm[4 * k + b], nm1, byte_to_float(r.px[1][k], b));
                r.v[4 * k + b] = make_float2(fmaxf(a0, 0.f), fmaxf(a1, 0.f));
            }
        }
    }
#pragma unroll
    for (int x = 32; x < 64; ++x) r.v[x] = make_float2(0.f, 0.f);
    r.tx = p.pad_tx[column_of<64>(t)];
}

// ---- float32 frames (pyorc's time_diff / smooth / edge_detect output) ------------------------------------------------------
// Tile: window w, rows [64 hb, 64 hb + 64) in the spectrum block of sub-group 2 w + hb (idle between the cross phases):
// [4 column blocks][64 rows][128 B] from 32-float x 64-row TMA boxes, SWIZZLE_128B.  F1(w): the thread's row of window w, the
// floats of its column parity -> component w of r.v, row sum; F2(w): mean over all four components, centre, centred second
// moment; F3: 0.5 / std of both windows, optional clip.  Two-pass moments like numpy's float path (rows_f1 .. f3 of piv_rows.cuh).
__device__ __forceinline__ unsigned char* r128_ftile(R128Smem& s, int g) {
    unsigned char* b = reinterpret_cast<unsigned char*>(&s.sub[g].park[0][0]);
    return b + ((1024u - (smem_u32(b) & 1023u)) & 1023u);
}
static_assert(sizeof(RSmem<R6>::park) >= 4 * 64 * 128 + 1024, "half a float32 window (plus alignment slack) must fit in a spectrum block");
__device__ __forceinline__ void r128_f1(R128Smem& s, RRegs<R6>& r, int sub, int t, int w) {
    const int p1 = sub >> 1, p2 = sub & 1;
    const int row = 2 * column_of<64>(t) + p1;
    const int rr = row & 63;
    const unsigned char* tile = r128_ftile(s, 2 * w + (row >> 6));
    float sum = 0.f;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 q = *reinterpret_cast<const float4*>(tile + h * 8192 + rr * 128 + ((j ^ (rr & 7)) << 4));
            const float a = p2 ? q.y : q.x, b = p2 ? q.w : q.z;
            const int x = 16 * h + 2 * j;
            if (w == 0) { r.v[x].x = a; r.v[x + 1].x = b; } else { r.v[x].y = a; r.v[x + 1].y = b; }
            sum += a + b;
        }
    }
    red_put_f32(&s.sub[sub].red[t >> 5][2 * w], sum, t);
}
__device__ __forceinline__ void r128_f2(R128Smem& s, RRegs<R6>& r, int sub, int t, int w) {
    float S = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int k = 0; k < 2; ++k) S += bits_f32(s.sub[g].red[k][2 * w]);
    }
    const float mean = S * (1.0f / (float)R128_NPX);
    float q = 0.f;
#pragma unroll
    for (int x = 0; x < 64; ++x) {
        if (w == 0) { r.v[x].x -= mean; q = fmaf(r.v[x].x, r.v[x].x, q); }
        else        { r.v[x].y -= mean; q = fmaf(r.v[x].y, r.v[x].y, q); }
    }
    red_put_f32(&s.sub[sub].red[t >> 5][2 * w + 1], q, t);
}
__device__ __forceinline__ void r128_f3(R128Smem& s, RRegs<R6>& r, int clip_norm) {
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        float Q = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int k = 0; k < 2; ++k) Q += bits_f32(s.sub[g].red[k][2 * w + 1]);
        }
        r.half_alpha_new[w] = Q > 0.f ? 0.5f * 128.0f * (1.0f / sqrtf(Q)) : 0.f;   // 0.5 / sqrt(Q / N), N = 128 * 128
    }
    r.dc_fix[0] = r.dc_fix[1] = 0.f;
    if (clip_norm) {
#pragma unroll
        for (int x = 0; x < 64; ++x) r.v[x] = make_float2(fmaxf(r.v[x].x, 0.f), fmaxf(r.v[x].y, 0.f));
    }
}

// float32 frames in padded mode (even windows of 34 .. 64 px of pyorc's float32 filters, e.g. a 4K user's 50 px window behind
// edge_detect): the embedding of r128_p1_pad / r128_p2_pad with the two-pass float moments of r128_f1 .. f3.  F1(w): the thread's row
// 2 sigma(t) + p1 of window w (if it is a window row), the floats of its column parity at the window's float offset `xoff` in the
// box -> component w of r.v[0 .. nx), everything else exactly 0, row sum; F2(w): mean over the (2 ny)(2 nx) window pixels, centre
// the window's pixels only, centred second moment; F3: 0.5 / std, optional clip, spectrum factor of the own column.  p.ny / p.nx
// are the COMPONENT size.  From the row transform on it is the uint8 padded path (r128_cross<PAD>, r128_p6_pad ...).
static_assert(sizeof(RSmem<R6>::park) >= R128_PFWIN + 1024, "a padded float32 window (plus alignment slack) must fit in a spectrum block");
__device__ __forceinline__ void r128_f1_pad(R128Smem& s, RRegs<R6>& r, int sub, int t, const RParams& p, int w, int xoff) {
    const int p1 = sub >> 1, p2 = sub & 1;
    const int row = 2 * column_of<64>(t) + p1;
    const bool rowok = row < 2 * p.ny;
    const float* src = reinterpret_cast<const float*>(r128_ftile(s, w)) + (rowok ? row : 0) * R128_PFW + xoff + p2;
    float sum = 0.f;
#pragma unroll
    for (int x = 0; x < 32; ++x) {
        const float val = (rowok && x < p.nx) ? src[2 * x] : 0.f;
        if (w == 0) r.v[x].x = val; else r.v[x].y = val;
        sum += val;
    }
#pragma unroll
    for (int x = 32; x < 64; ++x) {
        if (w == 0) r.v[x].x = 0.f; else r.v[x].y = 0.f;
    }
    red_put_f32(&s.sub[sub].red[t >> 5][2 * w], sum, t);
}
__device__ __forceinline__ void r128_f2_pad(R128Smem& s, RRegs<R6>& r, int sub, int t, const RParams& p, int w) {
    float S = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int k = 0; k < 2; ++k) S += bits_f32(s.sub[g].red[k][2 * w]);
    }
    const float mean = S / (float)(4 * p.ny * p.nx);
    const bool rowok = column_of<64>(t) < p.ny;
    float q = 0.f;
#pragma unroll
    for (int x = 0; x < 32; ++x) {
        const float m = (rowok && x < p.nx) ? mean : 0.f;   // pixels outside the window stay exactly 0
        if (w == 0) { r.v[x].x -= m; q = fmaf(r.v[x].x, r.v[x].x, q); }
        else        { r.v[x].y -= m; q = fmaf(r.v[x].y, r.v[x].y, q); }
    }
    red_put_f32(&s.sub[sub].red[t >> 5][2 * w + 1], q, t);
}
__device__ __forceinline__ void r128_f3_pad(R128Smem& s, RRegs<R6>& r, int t, const RParams& p) {
    const float npx = (float)(4 * p.ny * p.nx);
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        float Q = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int k = 0; k < 2; ++k) Q += bits_f32(s.sub[g].red[k][2 * w + 1]);
        }
        r.half_alpha_new[w] = Q > 0.f ? 0.5f * sqrtf(npx) * (1.0f / sqrtf(Q)) : 0.f;   // 0.5 / sqrt(Q / ((2 ny)(2 nx)))
    }
    r.dc_fix[0] = r.dc_fix[1] = 0.f;
    if (p.clip_norm) {
#pragma unroll
        for (int x = 0; x < 32; ++x) r.v[x] = make_float2(fmaxf(r.v[x].x, 0.f), fmaxf(r.v[x].y, 0.f));
    }
    r.tx = p.pad_tx[column_of<64>(t)];
}

// packed fp32 forms (piv_core.cuh): two instructions each
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return ctw<1>(a, b.x, b.y); }
__device__ __forceinline__ float2 cmulc(float2 p, float2 a) { return ctw<0>(a, p.x, p.y); }   // conj(p) * a
__device__ __forceinline__ float2 cfma(float2 m, float2 b, float2 a) {   // a + m * b = a + m.x (b.x, b.y) + m.y (-b.y, b.x)
    return pk_fma(make_float2(-b.y, b.x), make_float2(m.y, m.y), pk_fma(b, make_float2(m.x, m.x), a));
}

// Both windows of a spectrum bin travel together as one float4 (A0.x, A0.y, A1.x, A1.y).  The NEW separated spectra of all four
// components are published in the memory of RSmem::park ([33][64] float4 per component); the cross spectra travel back to
// their owners through the transpose blocks ([11][64] float4 per component and batch).  The PARKED spectra of the previous
// frame live in Tensor Memory, private to the thread that needs them (r128_cross).
__device__ __forceinline__ float4* r128_pub(R128Smem& s, int g, int ky, int t) {
    return &s.sub[g].park[ky][t];
}
__device__ __forceinline__ float4* r128_ret(R128Smem& s, int q, int sl, int t) {
    return reinterpret_cast<float4*>(&s.sub[q].X[0][0]) + sl * 64 + t;
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                   "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
// wait for the outstanding loads; the registers pass through the statement so that no use of them can be scheduled above it
__device__ __forceinline__ void tm_wait_ld16(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) :: "memory");
}

// exp(+2 pi i ky / 64), ky = 0 .. 32: phase of a carry along y (indexed at run time by the rolled batch loop below)
__device__ __constant__ float2 R128_TWY[33] = {{1.0000000000e+00f, 0.0000000000e+00f}, {9.9518472667e-01f, 9.8017140330e-02f}, {9.8078528040e-01f, 1.9509032202e-01f}, {9.5694033573e-01f, 2.9028467725e-01f}, {9.2387953251e-01f, 3.8268343237e-01f}, {8.8192126435e-01f, 4.7139673683e-01f}, {8.3146961230e-01f, 5.5557023302e-01f}, {7.7301045336e-01f, 6.3439328416e-01f}, {7.0710678119e-01f, 7.0710678119e-01f}, {6.3439328416e-01f, 7.7301045336e-01f}, {5.5557023302e-01f, 8.3146961230e-01f}, {4.7139673683e-01f, 8.8192126435e-01f}, {3.8268343237e-01f, 9.2387953251e-01f}, {2.9028467725e-01f, 9.5694033573e-01f}, {1.9509032202e-01f, 9.8078528040e-01f}, {9.8017140330e-02f, 9.9518472667e-01f}, {0.0f, 1.0f}, {-9.8017140330e-02f, 9.9518472667e-01f}, {-1.9509032202e-01f, 9.8078528040e-01f}, {-2.9028467725e-01f, 9.5694033573e-01f}, {-3.8268343237e-01f, 9.2387953251e-01f}, {-4.7139673683e-01f, 8.8192126435e-01f}, {-5.5557023302e-01f, 8.3146961230e-01f}, {-6.3439328416e-01f, 7.7301045336e-01f}, {-7.0710678119e-01f, 7.0710678119e-01f}, {-7.7301045336e-01f, 6.3439328416e-01f}, {-8.3146961230e-01f, 5.5557023302e-01f}, {-8.8192126435e-01f, 4.7139673683e-01f}, {-9.2387953251e-01f, 3.8268343237e-01f}, {-9.5694033573e-01f, 2.9028467725e-01f}, {-9.8078528040e-01f, 1.9509032202e-01f}, {-9.9518472667e-01f, 9.8017140330e-02f}, {-1.0f, 0.0f}};

// The cross phase.  On entry r.v holds Z_q(ky, own column) = FFT of (window 0 + i window 1) of the thread's component q = sub;
// on return r.v holds conj(G_q), G_q = C_q(window 0) + i C_q(window 1), ready for the inverse pass.
//
// The sixteen products conj(A_p) B_r of a spectrum bin feed the four C_q, each product exactly one of them.  Round 1 let every
// sub-group q gather all four parked A_p and all four new B_r of every bin from shared memory - each spectrum value was read
// four times, 1.1 MB of shared-memory reads per frame and CTA, a third of the frame time.  Now the BINS are dealt out
// instead of the components: thread (sub, t) takes the ky rows {b0 + sub, b0 + sub + 4, b0 + sub + 8} of each batch of 11 at
// its own column, for all four components: it reads the four new B_r once (published by their owners), forms the sixteen
// products and the four C_q with their phase factors, hands (R0, R1)_q back to the owner of component q through the transpose
// blocks, and keeps the scaled B_p as the next frame's parked spectra - in TENSOR MEMORY, since nobody else ever needs them
// again: 4 components x 9 bins x 4 floats = 144 columns per thread (.32x32b shape: thread t of warp w owns lane
// 32 (w % 4) + t; the two warps of a lane quarter use columns [0, 144) and [144, 288)).  Shared-memory traffic of the phase:
// 0.54 MB per frame instead of 1.5 MB.
//
// Code size matters more than instruction count here (the frame loop is far beyond the instruction caches: a fully
// unrolled version of this phase ran at HALF the speed), so the three batches share ONE copy of the code: the results are
// collected in lo[ky] = conj G(ky) and hi[ky] = conj G(-ky) (ky = 0 .. 32), a batch always writes lo[0..10] / hi[0..10], and both
// arrays are rotated by 11 between batches (66 register moves per batch) - after three batches they are in natural order.
template <bool PAD>
__device__ __forceinline__ void r128_cross(R128Smem& s, RRegs<R6>& r, int sub, int t, uint32_t tm, const RParams& p) {
    const float SCALE = PAD ? p.pad_scale : 1.0f / (4096.0f * (float)R128_NPX);   // 1/4096 of the 64x64 inverse, 1/N of the coefficient
    constexpr int B = R128_BATCH;
    const int pl = partner_lane_of<64>(t);
    // phase factor of a carry along x: exp(+2 pi i c / 64) for the own column c
    const int c = column_of<64>(t);
    float sn, cs;
    sincospif((float)c * (1.0f / 32.0f), &sn, &cs);
    const float2 mx = make_float2(cs, sn);
    if (t == 0) r.v[0] = pk_sub(r.v[0], make_float2(r.dc_fix[0], r.dc_fix[1]));   // Z_q(0, 0): thread 0 of a sub-group owns column 0 (r128_p2)
    // -- publish the separated new spectra B0, B1 of the own component, all ky
#pragma unroll
    for (int ky = 0; ky <= 32; ++ky) {
        const float2 pz = shfl2(r.v[(64 - ky) % 64], pl);
        float2 a0, a1;
        separate(r.v[ky], pz, r.half_alpha_new[0], r.half_alpha_new[1], a0, a1);
        *r128_pub(s, sub, ky, t) = make_float4(a0.x, a0.y, a1.x, a1.y);
    }
    uint32_t cur[16];
    tm_ld16(tm, cur);            // parked spectra of the first own bin (batch 0, j = 0); in flight across the barrier
    __syncthreads();
    float2 lo[33], hi[33];
    int n_bin = 0;               // running index of the own bins: 3 * batch + j
#pragma unroll 1
    for (int b0 = 0; b0 < 33; b0 += B) {
        // -- own bins of this batch: ky = b0 + sub + 4 j
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int sl = sub + 4 * j;
            if (sl < B) {        // warp-uniform (sub-group 3 has two bins per batch)
                tm_wait_ld16(cur);
                const int ky = b0 + sl;
                const float2 my = R128_TWY[ky];
                float2 tile_f = make_float2(1.f, 0.f);   // padded mode: the new window in its tiled role, T(ky, own column) = Ty(ky) Tx
                if (PAD) tile_f = ctw<1>(p.pad_ty[ky], r.tx.x, r.tx.y);
                float4 nw[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) nw[p] = *r128_pub(s, p, ky, t);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float2 R[2];
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        float2 term[4];
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const float2 a = make_float2(__uint_as_float(cur[4 * p + 2 * w]), __uint_as_float(cur[4 * p + 2 * w + 1]));
                            const float4 n4 = nw[p ^ q];
                            term[p] = cmulc(a, w == 0 ? make_float2(n4.x, n4.y) : make_float2(n4.z, n4.w));
                        }
                        // p = (p1, p2) = (p >> 1, p & 1), carry s = p and q: R = t0 + mx^q2 t1 + my^q1 (t2 + mx^q2 t3)
                        const float2 l = (q & 1) ? cfma(mx, term[1], term[0]) : pk_add(term[0], term[1]);
                        const float2 h = (q & 1) ? cfma(mx, term[3], term[2]) : pk_add(term[2], term[3]);
                        R[w] = (q & 2) ? cfma(my, h, l) : pk_add(l, h);
                        if (PAD) R[w] = ctw<1>(R[w], tile_f.x, tile_f.y);
                    }
                    *r128_ret(s, q, sl, t) = make_float4(R[0].x, R[0].y, R[1].x, R[1].y);
                }
                // the parked spectra of the next own bin (of this batch, else of the next one) travel while the results are stored
                if (j < 2 && sub + 4 * (j + 1) < B) tm_ld16(tm + 16 * (n_bin + j + 1), cur);
                else if (b0 + B < 33) tm_ld16(tm + 16 * (n_bin + 3), cur);
                // the new spectra become the parked ones (scaled once, here)
                uint32_t out[16];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float2 s0 = pk_scale(make_float2(nw[p].x, nw[p].y), SCALE), s1 = pk_scale(make_float2(nw[p].z, nw[p].w), SCALE);
                    out[4 * p] = __float_as_uint(s0.x); out[4 * p + 1] = __float_as_uint(s0.y);
                    out[4 * p + 2] = __float_as_uint(s1.x); out[4 * p + 3] = __float_as_uint(s1.y);
                }
                tm_st16(tm + 16 * (n_bin + j), out);
            }
        }
        n_bin += 3;
        __syncthreads();
        // -- the cross spectra of the own component come back
#pragma unroll
        for (int sl = 0; sl < B; ++sl) {
            const float4 R = *r128_ret(s, sub, sl, t);
            const float2 R0 = make_float2(R.x, R.y), R1 = make_float2(R.z, R.w);
            lo[sl] = pk_sub(make_float2(R0.x, -R0.y), make_float2(R1.y, R1.x));   // conj(G), G = R0 + i R1
            hi[sl] = shfl2(cross_mirror(R0, R1), pl);                             // conj(G(-ky)); unused for ky = 0, 32
        }
        // -- rotate both register arrays by one batch
        float2 tl[B], th[B];
#pragma unroll
        for (int k = 0; k < B; ++k) { tl[k] = lo[k]; th[k] = hi[k]; }
#pragma unroll
        for (int k = 0; k + B < 33; ++k) { lo[k] = lo[k + B]; hi[k] = hi[k + B]; }
#pragma unroll
        for (int k = 0; k < B; ++k) { lo[33 - B + k] = tl[k]; hi[33 - B + k] = th[k]; }
        __syncthreads();         // the next batch (or the transposes of the inverse pass) overwrite the blocks
    }
    tm_wait_st();
#pragma unroll
    for (int k = 0; k <= 32; ++k) r.v[k] = lo[k];
#pragma unroll
    for (int k = 1; k < 32; ++k) r.v[64 - k] = hi[k];
}

// P6: rows holding the block maximum look for their first matching column in the reference's (fftshifted) order
__device__ __forceinline__ void r128_p6(R128Smem& s, RRegs<R6>& r, int sub, int t) {
    const int q1 = sub >> 1, q2 = sub & 1;
    const int si = (2 * column_of<64>(t) + q1 + 64) & 127;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        float M = 0.f, S = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int k = 0; k < 2; ++k) { M = fmaxf(M, bits_f(s.sub[g].red[k][4 + w])); S += bits_f(s.sub[g].red[k][6 + w]); }
        }
        r.cmaxv[w] = M; r.sumv[w] = S;
        unsigned long long key = ~0ull;
        if (r.rowmax[w] == M) {
            int first = 64;
            // reference column j = (2 x + q2 + 64) % 128 ascends with pos = (x + 32) % 64: scan pos descending
#pragma unroll
            for (int pos = 63; pos >= 0; --pos) {
                const int x = (pos + 32) % 64;
                const float val = w == 0 ? r.v[x].x : r.v[x].y;
                first = (val == M) ? pos : first;
            }
            int j = 2 * first + q2;
            int i = si;
            if (r.dead[w]) { i = 0; j = 0; }   // all-zero plane: every element is the maximum -> flat index 0
            key = (unsigned long long)(i * 128 + j);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        if ((t & 31) == 0) s.sub[sub].redk[t >> 5][w] = key;
    }
}

// P7: the three reference rows around each peak; a reference row is shared by the two components with the same q1
__device__ __forceinline__ void r128_p7(R128Smem& s, RRegs<R6>& r, int sub, int t) {
    const int q1 = sub >> 1, q2 = sub & 1;
    const int si = (2 * column_of<64>(t) + q1 + 64) & 127;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        unsigned long long key = ~0ull;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int k = 0; k < 2; ++k) key = s.sub[g].redk[k][w] < key ? s.sub[g].redk[k][w] : key;
        }
        const int idx = (int)key;
        r.pi[w] = idx >> 7; r.pj[w] = idx & 127;
        const int d = si - r.pi[w];
        if (d >= -1 && d <= 1) {
            float* row = &s.nb[w][d + 1][0];
#pragma unroll
            for (int x = 0; x < 64; ++x) {
                const float val = r.dead[w] ? 0.f : (w == 0 ? r.v[x].x : r.v[x].y);
                row[(2 * x + q2 + 64) & 127] = val;
            }
        }
    }
}

// P6 / P7 (padded): the thread's row holds the lags (2 sigma(t) + q1, 2 x + q2), x < nx (component size); the reference plane
// is the fftshifted (2 ny) x (2 nx) plane of the window's own size: row / column (lag + n/2) % n, the lags >= n/2 first.
__device__ __forceinline__ void r128_p6_pad(R128Smem& s, RRegs<R6>& r, int sub, int t, const RParams& p) {
    const int q1 = sub >> 1, q2 = sub & 1;
    const int hy = p.ny, hx = p.nx;
    const bool rowok = column_of<64>(t) < hy;
    const int si = rowok ? shifted_index(2 * column_of<64>(t) + q1, 2 * hy) : 0;
    const int xs = (hx - q2 + 1) >> 1;            // first x whose lag 2 x + q2 is >= nx / 2 (= hx)
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        float M = 0.f, S = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int k = 0; k < 2; ++k) { M = fmaxf(M, bits_f(s.sub[g].red[k][4 + w])); S += bits_f(s.sub[g].red[k][6 + w]); }
        }
        r.cmaxv[w] = M; r.sumv[w] = S;
        unsigned long long key = ~0ull;
        if (r.rowmax[w] == M && rowok) {
            unsigned mm = 0u;
#pragma unroll
            for (int x = 0; x < 32; ++x) {
                const float val = w == 0 ? r.v[x].x : r.v[x].y;
                mm |= (val == M) ? (1u << x) : 0u;
            }
            mm &= hx >= 32 ? 0xffffffffu : ((1u << hx) - 1u);
            const unsigned seg_hi = xs >= 32 ? 0u : (mm >> xs);
            int j;
            if (seg_hi) j = 2 * (xs + __ffs((int)seg_hi) - 1) + q2 - hx;
            else j = 2 * (__ffs((int)mm) - 1) + q2 + hx;
            key = (unsigned long long)(si * 2 * hx + j);
            if (M == 0.f || r.dead[w]) key = 0ull;   // all-zero plane: every element is the maximum -> flat index 0
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        if ((t & 31) == 0) s.sub[sub].redk[t >> 5][w] = key;
    }
}

__device__ __forceinline__ void r128_p7_pad(R128Smem& s, RRegs<R6>& r, int sub, int t, const RParams& p) {
    const int q1 = sub >> 1, q2 = sub & 1;
    const int hy = p.ny, hx = p.nx;
    const bool rowok = column_of<64>(t) < hy;
    const int si = rowok ? shifted_index(2 * column_of<64>(t) + q1, 2 * hy) : -8;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        unsigned long long key = ~0ull;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int k = 0; k < 2; ++k) key = s.sub[g].redk[k][w] < key ? s.sub[g].redk[k][w] : key;
        }
        const int idx = (int)key;
        r.pi[w] = idx / (2 * hx); r.pj[w] = idx - r.pi[w] * (2 * hx);
        const int d = si - r.pi[w];
        if (d >= -1 && d <= 1) {
            // lag l = 2 x + q2 goes to column l + hx (l < hx) or l - hx: two base pointers, static offsets
            float* row = &s.nb[w][d + 1][0];
            float* base_lo = row + hx + q2;
            float* base_hi = row - hx + q2;
#pragma unroll
            for (int x = 0; x < 32; ++x) {
                if (x < hx) {
                    const float val = r.dead[w] ? 0.f : (w == 0 ? r.v[x].x : r.v[x].y);
                    (2 * x + q2 < hx ? base_lo : base_hi)[2 * x] = val;
                }
            }
        }
    }
}

// P8: Gaussian fit + outputs by threads 0 / 1 of the group (pyorc/velocimetry/ffpiv.py:465-466 + ffpiv.u_v_displacement)
__device__ __forceinline__ void r128_p8(R128Smem& s, RRegs<R6>& r, int tid, const RParams& p, const RUnit& un, int pair, int ny = 128, int nx = 128) {
    if (tid >= 2) return;
    const int w = tid;
    if (w == 1 && !un.valid1) return;
    const float* nb = &s.nb[w][0][0];
    const int pi = w == 0 ? r.pi[0] : r.pi[1], pj = w == 0 ? r.pj[0] : r.pj[1];
    const float cmax = w == 0 ? r.cmaxv[0] : r.cmaxv[1];
    const float mean = (w == 0 ? r.sumv[0] : r.sumv[1]) / (float)(ny * nx);
    float uu, vv;
    if (pi == 0 || pi == ny - 1 || pj == 0 || pj == nx - 1) {
        if (p.border_nan) { uu = nanf(""); vv = nanf(""); }
        else { uu = (float)(pj - nx / 2); vv = (float)(pi - ny / 2); }
    } else {
        const float eps = p.gauss_eps;
        const float lc = logf(cmax + eps);
        const float ll = logf(nb[0 * 128 + pj] + eps), lr = logf(nb[2 * 128 + pj] + eps);
        const float ld = logf(nb[1 * 128 + pj - 1] + eps), lu = logf(nb[1 * 128 + pj + 1] + eps);
        vv = ((float)pi + (ll - lr) / (2.f * ll - 4.f * lc + 2.f * lr)) - (float)(ny / 2);
        uu = ((float)pj + (ld - lu) / (2.f * ld - 4.f * lc + 2.f * lu)) - (float)(nx / 2);
    }
    float oc = cmax, os = cmax / mean;
    const int widx = w == 0 ? un.w[0] : un.w[1];
    if (p.keep && !p.keep[widx]) { uu = vv = oc = os = nanf(""); }
    const long long o = (long long)pair * p.n_rows * p.n_cols + widx;
    p.u[o] = uu; p.v[o] = vv; p.cmax[o] = oc; p.s2n[o] = os;
    if (p.peer.n) peer_store(p.peer, pair, (long long)p.n_rows * p.n_cols, widx, uu, vv, oc, os);
}

// Ensemble mode (rows_ens of piv_rows.cuh for the polyphase layout): thresholds on the pair's max / mean, then the plane is added
// to the window's accumulator in HBM with fire-and-forget reductions at the L2.  A thread holds every other element of one
// reference row and the lanes of a warp hold different rows, so reductions straight from the registers touch 32 sectors per
// instruction (measured 7.7 M windows/s, bound by the L2's atomic rate).  The planes are therefore put in reference order in
// shared memory first - in the memory of the published spectra, idle after the cross phase; row pitch 129 floats: the stores of a
// warp (32 rows, one column) and the loads (one row, 32 columns) are both conflict-free - and added by all 256 threads with
// consecutive lanes on consecutive floats: 4 sectors per instruction.  Every element is added by the same thread in every frame,
// in frame order = the reference's np.sum(corr, axis=0) (pyorc/velocimetry/ffpiv.py:345-376).
__device__ __forceinline__ float* r128_stage(R128Smem& s, int w, int row) {     // window w, reference row: 64 rows per spectrum block
    return reinterpret_cast<float*>(&s.sub[2 * w + (row >> 6)].park[0][0]) + (row & 63) * 129;
}
static_assert(sizeof(RSmem<R6>::park) >= 64 * 129 * sizeof(float), "half a staged plane must fit in a spectrum block");
__device__ __forceinline__ void r128_ens(R128Smem& s, RRegs<R6>& r, int sub, int t, int tid, const RParams& p, const RUnit& un, int pair, bool store) {
    const int q1 = sub >> 1, q2 = sub & 1;
    const int si = (2 * column_of<64>(t) + q1 + 64) & 127;
    const long long nw = (long long)p.n_rows * p.n_cols;
    bool okw[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        float M = 0.f, S = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int k = 0; k < 2; ++k) { M = fmaxf(M, bits_f(s.sub[g].red[k][4 + w])); S += bits_f(s.sub[g].red[k][6 + w]); }
        }
        const float ratio = M / (S / (float)R128_NPX);
        const int widx = w == 0 ? un.w[0] : un.w[1];
        bool ok = (M >= p.corr_min) && (ratio >= p.s2n_min) && !r.dead[w];   // dead: 0 / 0 = NaN fails the test in the reference
        if (p.keep && !p.keep[widx]) ok = false;                             // NaN plane in the reference -> masked out
        const bool wr = store && !(w == 1 && !un.valid1);
        okw[w] = ok && wr;                                                   // the same for every thread of the CTA
        if (okw[w]) {
            float* row = r128_stage(s, w, si);
#pragma unroll
            for (int x = 0; x < 64; ++x) row[(2 * x + q2 + 64) & 127] = w == 0 ? r.v[x].x : r.v[x].y;
        }
        if (wr && tid == 0) {
            const long long o = (long long)pair * nw + widx;
            p.cmax[o] = ok ? M : 0.f;
            p.s2n[o] = ok ? ratio : 0.f;
            if (ok && M > 1e-6f) p.ens_count[widx] += 1.f;
        }
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        if (!okw[w]) continue;
        float* dst = p.ens_sum + (long long)(w == 0 ? un.w[0] : un.w[1]) * R128_NPX;
        const int col = tid & 127;
#pragma unroll 8
        for (int i = 0; i < 64; ++i) {
            const int row = 2 * i + (tid >> 7);
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + row * 128 + col), "f"(r128_stage(s, w, row)[col]) : "memory");
        }
    }
}

// Ensemble mode of the padded kernel: the same staging, for a plane of the window's own size [2 ny][2 nx] (reference order, the
// layout of the accumulators).  A staged plane (pitch 2 nx + 1 floats, at most 64 x 65) fits in one spectrum block.
__device__ __forceinline__ void r128_ens_pad(R128Smem& s, RRegs<R6>& r, int sub, int t, int tid, const RParams& p, const RUnit& un, int pair, bool store) {
    const int q1 = sub >> 1, q2 = sub & 1;
    const int hy = p.ny, hx = p.nx, ny = 2 * p.ny, nx = 2 * p.nx, pitch = nx + 1;
    const bool rowok = column_of<64>(t) < hy;
    const int si = rowok ? shifted_index(2 * column_of<64>(t) + q1, ny) : 0;
    const long long nw = (long long)p.n_rows * p.n_cols;
    bool okw[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        float M = 0.f, S = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
            for (int k = 0; k < 2; ++k) { M = fmaxf(M, bits_f(s.sub[g].red[k][4 + w])); S += bits_f(s.sub[g].red[k][6 + w]); }
        }
        const float ratio = M / (S / (float)(ny * nx));
        const int widx = w == 0 ? un.w[0] : un.w[1];
        bool ok = (M >= p.corr_min) && (ratio >= p.s2n_min) && !r.dead[w];   // dead: 0 / 0 = NaN fails the test in the reference
        if (p.keep && !p.keep[widx]) ok = false;                             // NaN plane in the reference -> masked out
        const bool wr = store && !(w == 1 && !un.valid1);
        okw[w] = ok && wr;                                                   // the same for every thread of the CTA
        if (okw[w] && rowok) {
            // lag l = 2 x + q2 goes to column l + hx (l < hx) or l - hx: two base pointers, static offsets
            float* row = reinterpret_cast<float*>(&s.sub[w].park[0][0]) + si * pitch;
            float* base_lo = row + hx + q2;
            float* base_hi = row - hx + q2;
#pragma unroll
            for (int x = 0; x < 32; ++x) {
                if (x < hx) (2 * x + q2 < hx ? base_lo : base_hi)[2 * x] = w == 0 ? r.v[x].x : r.v[x].y;
            }
        }
        if (wr && tid == 0) {
            const long long o = (long long)pair * nw + widx;
            p.cmax[o] = ok ? M : 0.f;
            p.s2n[o] = ok ? ratio : 0.f;
            if (ok && M > 1e-6f) p.ens_count[widx] += 1.f;
        }
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        if (!okw[w]) continue;
        float* dst = p.ens_sum + (long long)(w == 0 ? un.w[0] : un.w[1]) * (ny * nx);
        const float* src = reinterpret_cast<const float*>(&s.sub[w].park[0][0]);
        for (int e = tid; e < ny * nx; e += 256) {
            const int row = e / nx, col = e - row * nx;
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + e), "f"(src[row * pitch + col]) : "memory");
        }
    }
}
static_assert(sizeof(RSmem<R6>::park) >= 64 * 65 * sizeof(float), "a staged padded plane must fit in a spectrum block");

// optional triage dump of the full planes (fftshifted, clipped): every thread writes its 64 elements of one reference row
__device__ __forceinline__ void r128_dump_planes(RRegs<R6>& r, int sub, int t, const RParams& p, const RUnit& un, int pair) {
    if (!p.planes) return;
    const int q1 = sub >> 1, q2 = sub & 1;
    const int si = (2 * column_of<64>(t) + q1 + 64) & 127;
    const long long nw = (long long)p.n_rows * p.n_cols;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        if (w == 1 && !un.valid1) continue;
        float* dst = p.planes + (((long long)pair * nw + un.w[w]) * 128 + si) * 128;
#pragma unroll
        for (int x = 0; x < 64; ++x) dst[(2 * x + q2 + 64) & 127] = r.dead[w] ? 0.f : (w == 0 ? r.v[x].x : r.v[x].y);
    }
}
// padded mode: the planes are [2 ny][2 nx] (the window's own size), reference order
__device__ __forceinline__ void r128_dump_planes_pad(RRegs<R6>& r, int sub, int t, const RParams& p, const RUnit& un, int pair) {
    if (!p.planes) return;
    const int q1 = sub >> 1, q2 = sub & 1;
    const int hy = p.ny, hx = p.nx;
    if (column_of<64>(t) >= hy) return;
    const int si = shifted_index(2 * column_of<64>(t) + q1, 2 * hy);
    const long long nw = (long long)p.n_rows * p.n_cols;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        if (w == 1 && !un.valid1) continue;
        float* dst = p.planes + (((long long)pair * nw + un.w[w]) * (2 * hy) + si) * (2 * hx);
#pragma unroll
        for (int x = 0; x < 32; ++x) {
            if (x < hx) dst[shifted_index(2 * x + q2, 2 * hx)] = r.dead[w] ? 0.f : (w == 0 ? r.v[x].x : r.v[x].y);
        }
    }
}
#endif  // __CUDACC__

}  // namespace b2piv

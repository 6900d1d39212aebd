// piv_direct.cuh - any-size interrogation windows (up to 64x64, e.g. pyorc's 10, 20, 26, 50) by DIRECT circular
// cross-correlation in shared memory: c(s) = 1/N * sum_x a(x) b(x+s), exact fp32 products, no FFT.
//
// O(N^2) per window, so this is the compatibility path (pyorc takes its window size from the camera configuration:
// 25 -> 26 by round_to_even, pyorc/api/frames.py:159-167; its only golden test uses 10, tests/test_frames.py:142);
// power-of-two windows take the FFT kernels.  One CTA per (frame pair, window); same outputs and the same
// reduction / peak-fit semantics as piv_core.cuh.  Phases are __host__ __device__ for tests/emul.
#pragma once
#include "piv_core.cuh"

namespace b2piv {

constexpr int DNT = 256;   // threads per CTA

struct DView {             // views into the CTA's dynamic shared memory
    float* a;              // [wy][wx] centred window of frame k
    float* b2;             // [2wy][2wx] centred window of frame k+1, tiled 2x2 so that (y+sy, x+sx) never wraps
    float* plane;          // [wy][wx] correlation plane, fftshifted + clipped
    unsigned long long* red;   // [8 warps][8 slots]
    float* scal;           // [4]: mean_a, mean_b, scale, unused
    int wy, wx;
};
B2_HD size_t direct_smem_bytes(int wy, int wx) { return (size_t)(6 * wy * wx) * sizeof(float) + 64 * sizeof(unsigned long long) + 16; }
B2_HD DView direct_view(unsigned char* base, int wy, int wx) {
    DView v;
    v.red = reinterpret_cast<unsigned long long*>(base);
    v.scal = reinterpret_cast<float*>(base + 64 * sizeof(unsigned long long));
    v.a = v.scal + 4;
    v.b2 = v.a + wy * wx;
    v.plane = v.b2 + 4 * wy * wx;
    v.wy = wy; v.wx = wx;
    return v;
}

// reduction helpers on the raw scratch (same protocol as piv_core.cuh: deposit per warp, total after a barrier)
B2_HD void d_dep_sum(DView& s, int tid, int slot, float v) {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) s.red[(tid >> 5) * 8 + slot] = (unsigned long long)__float_as_uint(v);
#else
    union { float f; unsigned u; } a, b;
    a.u = (unsigned)s.red[(tid >> 5) * 8 + slot];
    b.f = a.f + v;
    s.red[(tid >> 5) * 8 + slot] = (unsigned long long)b.u;
#endif
}
B2_HD void d_dep_max(DView& s, int tid, int slot, unsigned long long v) {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    if ((tid & 31) == 0) s.red[(tid >> 5) * 8 + slot] = v;
#else
    if (v > s.red[(tid >> 5) * 8 + slot]) s.red[(tid >> 5) * 8 + slot] = v;
#endif
}
B2_HD float d_tot_sum(const DView& s, int slot) {
    float t = 0.f;
    for (int w = 0; w < DNT / 32; ++w) { union { float f; unsigned u; } a; a.u = (unsigned)s.red[w * 8 + slot]; t += a.f; }
    return t;
}
B2_HD unsigned long long d_tot_max(const DView& s, int slot) {
    unsigned long long t = 0;
    for (int w = 0; w < DNT / 32; ++w) t = s.red[w * 8 + slot] > t ? s.red[w * 8 + slot] : t;
    return t;
}

// D1: load both windows, partial sums
B2_HD void direct_load(DView& s, int tid, const Params& p, int pair, int widx) {
    const unsigned char* base = (const unsigned char*)p.frames + (long long)pair * p.frame_stride;
    const int r = widx / p.n_cols, c = widx % p.n_cols;
    const long long off = (long long)(r * p.sy) * p.pitch;
    const int x0 = c * p.sx;
    float fa = 0.f, fb = 0.f;
    for (int e = tid; e < s.wy * s.wx; e += DNT) {
        const int y = e / s.wx, x = e % s.wx;
        float a, b;
        if (!p.is_f32) {
            const unsigned char* ra = base + off + (long long)y * p.pitch + x0 + x;
            a = (float)ra[0]; b = (float)ra[p.frame_stride];
        } else {
            const float* ra = (const float*)(base + off + (long long)y * p.pitch) + x0 + x;
            a = ra[0];
            b = *(const float*)((const unsigned char*)ra + p.frame_stride);
        }
        s.a[e] = a;
        s.b2[y * 2 * s.wx + x] = b;
        fa += a; fb += b;
    }
    d_dep_sum(s, tid, 0, fa);
    d_dep_sum(s, tid, 1, fb);
}
// D2: centre (+clip), centred second moments
B2_HD void direct_center(DView& s, int tid, const Params& p) {
    const float n = (float)(s.wy * s.wx);
    const float ma = d_tot_sum(s, 0) / n, mb = d_tot_sum(s, 1) / n;
    float qa = 0.f, qb = 0.f;
    for (int e = tid; e < s.wy * s.wx; e += DNT) {
        const int y = e / s.wx, x = e % s.wx;
        float a = s.a[e] - ma, b = s.b2[y * 2 * s.wx + x] - mb;
        qa += a * a; qb += b * b;
        if (p.clip_norm) { a = a < 0.f ? 0.f : a; b = b < 0.f ? 0.f : b; }
        s.a[e] = a;
        // replicate b into the 2x2 tiling
        s.b2[y * 2 * s.wx + x] = b;
        s.b2[y * 2 * s.wx + x + s.wx] = b;
        s.b2[(y + s.wy) * 2 * s.wx + x] = b;
        s.b2[(y + s.wy) * 2 * s.wx + x + s.wx] = b;
    }
    d_dep_sum(s, tid, 2, qa);
    d_dep_sum(s, tid, 3, qb);
}
// D3: correlate, clip, reduce.  Output index e is in fftshifted coordinates; lag = (i - wy/2, j - wx/2) mod size.
B2_HD void direct_correlate(DView& s, int tid, const Params& p, int pair, int widx) {
    const double n = (double)(s.wy * s.wx);
    const double va = (double)d_tot_sum(s, 2) / n, vb = (double)d_tot_sum(s, 3) / n;
    const float scale = (va > 0.0 && vb > 0.0) ? (float)(1.0 / (n * sqrt(va) * sqrt(vb))) : 0.f;
    unsigned long long best = 0ull;
    float sum = 0.f;
    for (int e = tid; e < s.wy * s.wx; e += DNT) {
        const int i = e / s.wx, j = e % s.wx;
        const int sy = (i + s.wy - s.wy / 2) % s.wy, sx = (j + s.wx - s.wx / 2) % s.wx;
        float acc = 0.f;
        for (int y = 0; y < s.wy; ++y) {
            const float* ar = s.a + y * s.wx;
            const float* br = s.b2 + (y + sy) * 2 * s.wx + sx;
            for (int x = 0; x < s.wx; ++x) acc = fmaf(ar[x], br[x], acc);
        }
        const float v = scale == 0.f ? 0.f : clip01(acc * scale);
        s.plane[e] = v;
        sum += v;
        union { float f; unsigned u; } cv; cv.f = v;
        const unsigned long long key = ((unsigned long long)cv.u << 32) | (unsigned long long)(0xffffffffu - (unsigned)e);
        best = key > best ? key : best;
        if (p.planes) p.planes[((long long)pair * p.n_rows * p.n_cols + widx) * (s.wy * s.wx) + e] = v;
    }
    d_dep_max(s, tid, 4, best);
    d_dep_sum(s, tid, 5, sum);
}
// D4: peak fit + outputs (thread 0)
B2_HD void direct_peak(DView& s, int tid, const Params& p, int pair, int widx) {
    if (tid != 0) return;
    const unsigned long long key = d_tot_max(s, 4);
    union { float f; unsigned u; } cv; cv.u = (unsigned)(key >> 32);
    const float cmax = cv.f;
    const int idx = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
    const int pi = idx / s.wx, pj = idx % s.wx;
    const float mean = d_tot_sum(s, 5) / (float)(s.wy * s.wx);
    float uu, vv;
    if (pi == 0 || pi == s.wy - 1 || pj == 0 || pj == s.wx - 1) {
        if (p.border_nan) { uu = nanf(""); vv = nanf(""); }
        else { uu = (float)(pj - s.wx / 2); vv = (float)(pi - s.wy / 2); }
    } else {
        const float eps = p.gauss_eps;
        const float lc = logf(cmax + eps);
        const float ll = logf(s.plane[(pi - 1) * s.wx + pj] + eps), lr = logf(s.plane[(pi + 1) * s.wx + pj] + eps);
        const float ld = logf(s.plane[pi * s.wx + pj - 1] + eps), lu = logf(s.plane[pi * s.wx + pj + 1] + eps);
        vv = ((float)pi + (ll - lr) / (2.f * ll - 4.f * lc + 2.f * lr)) - (float)(s.wy / 2);
        uu = ((float)pj + (ld - lu) / (2.f * ld - 4.f * lc + 2.f * lu)) - (float)(s.wx / 2);
    }
    float oc = cmax, os = cmax / mean;
    if (p.keep && !p.keep[widx]) { uu = vv = oc = os = nanf(""); }
    const long long o = (long long)pair * p.n_rows * p.n_cols + widx;
    p.u[o] = uu; p.v[o] = vv; p.cmax[o] = oc; p.s2n[o] = os;
    if (p.peer.n) peer_store(p.peer, pair, (long long)p.n_rows * p.n_cols, widx, uu, vv, oc, os);
}

}  // namespace b2piv

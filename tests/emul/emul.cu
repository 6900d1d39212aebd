// TEST-ONLY host emulator: runs the kernel's barrier-delimited phases (pyorc_b200/csrc/piv_core.cuh) on the CPU,
// "thread by thread, phase by phase", to validate index mathematics without a GPU.  Never loaded by pyorc_b200.
#include "../../pyorc_b200/csrc/piv_core.cuh"
#include "../../pyorc_b200/csrc/piv_rows.cuh"
#include "../../pyorc_b200/csrc/piv_direct.cuh"
#include <vector>
#include <cstring>
#include <cmath>
using namespace b2piv;

template <class C>
static void zero_red(Smem<C>& s) { memset(s.red, 0, sizeof(s.red)); }

#define ALL(expr) for (int tid = 0; tid < C::NT; ++tid) { expr; }

template <class C>
static int run(Params p) {
    std::vector<float2> twx(C::WX), twy(C::WY);
    for (int j = 0; j < C::WX; ++j) twx[j] = make_float2((float)cos(2 * M_PI * j / C::WX), (float)-sin(2 * M_PI * j / C::WX));
    for (int j = 0; j < C::WY; ++j) twy[j] = make_float2((float)cos(2 * M_PI * j / C::WY), (float)-sin(2 * M_PI * j / C::WY));
    Smem<C>* sp = new Smem<C>();
    Smem<C>& s = *sp;
    ALL(phase_init<C>(s, tid, twx.data(), twy.data()));
    const int nw = p.n_rows * p.n_cols;
    const int per_pair = (C::NWIN == 2) ? (nw + 1) / 2 : nw;
    for (int item = 0; item < per_pair * p.n_pairs; ++item) {
        Item it = decode_item<C>(p, item);
        zero_red(s); ALL(phase_load<C>(s, tid, p, it));
        ALL(phase_stats<C>(s, tid, p));
        zero_red(s); ALL(phase_center<C>(s, tid, p));
        ALL(phase_stats_f32<C>(s, tid, p));
        ALL(phase_embed<C>(s, tid, p));
        ALL((fft_pass<C, C::NWIN, 0, 0, 0>(s, tid)));
        ALL((fft_pass<C, C::NWIN, 0, 1, 0>(s, tid)));
        ALL((fft_pass<C, C::NWIN, 1, 0, 0>(s, tid)));
        ALL((fft_pass<C, C::NWIN, 1, 1, 0>(s, tid)));
        ALL(phase_cross<C>(s, tid));
        ALL((fft_pass<C, 1, 1, 1, 1>(s, tid)));
        ALL((fft_pass<C, 1, 1, 0, 1>(s, tid)));
        ALL((fft_pass<C, 1, 0, 1, 1>(s, tid)));
        ALL((fft_pass<C, 1, 0, 0, 1>(s, tid)));
        zero_red(s); ALL(phase_reduce<C>(s, tid, p, it));
        ALL(phase_peak<C>(s, tid, p, it));
    }
    delete sp;
    return 0;
}

extern "C" int b2piv_emul_pairs(const void* frames, int n_frames, int H, int W, int is_f32, int wy, int wx, int oy, int ox,
                                int nwin, int clip_norm, int border_nan, float eps, const unsigned char* keep,
                                float* u, float* v, float* cmax, float* s2n, float* planes) {
    Params p;
    memset(&p, 0, sizeof(p));
    p.frames = frames;
    p.pitch = W * (is_f32 ? 4 : 1);
    p.frame_stride = (long long)H * p.pitch;
    p.is_f32 = is_f32;
    p.n_rows = (H - wy) / (wy - oy) + 1;
    p.n_cols = (W - wx) / (wx - ox) + 1;
    p.sy = wy - oy; p.sx = wx - ox;
    p.n_pairs = n_frames - 1;
    p.clip_norm = clip_norm; p.border_nan = border_nan; p.gauss_eps = eps; p.keep = keep;
    p.u = u; p.v = v; p.cmax = cmax; p.s2n = s2n; p.planes = planes;
    p.ny = wy; p.nx = wx;
    {   // non power-of-two windows: padded plane, same rule as b2piv.cu plane_shape()
        auto pow2ok = [](int n) { return n == 16 || n == 32 || n == 64 || n == 128; };
        auto pad = [](int n) { int w = 16; while (w < 2 * n) w <<= 1; return w; };
        if (!(pow2ok(wy) && pow2ok(wx))) { int a = pad(wy), b = pad(wx); if (a != b && !((a == 32 && b == 64) || (a == 64 && b == 32))) a = b = (a > b ? a : b); wy = a; wx = b; }
    }
    const bool padded = !(wy == p.ny && wx == p.nx);
#define CASE(Y, X, T) if (wy == Y && wx == X) return padded ? (nwin == 2 ? run<Cfg<Y, X, T, 2, true>>(p) : run<Cfg<Y, X, T, 1, true>>(p)) : (nwin == 2 ? run<Cfg<Y, X, T, 2, false>>(p) : run<Cfg<Y, X, T, 1, false>>(p));
    CASE(16, 16, 64) CASE(32, 32, 128) CASE(64, 64, 256) CASE(128, 128, 256) CASE(32, 64, 128) CASE(64, 32, 128) CASE(64, 128, 256)
    return -1;
}


// ---- row-per-thread kernel (piv_rows.cuh): lock-step emulation of the W threads of one group ---------------------
// lock-step emulation of transpose_device (sub-phases separated where the device has a CTA barrier or a __syncwarp)
template <class R, bool FWD>
static void emul_transpose(RSmem<R>& s, std::vector<RRegs<R>>& regs) {
    constexpr int W = R::W;
    auto store = [&](int t, int k, int blk) {
        const int lane = t & 31;
        if (k == 0) tr_store_set<R, 0>(s.X[blk], regs[t], lane); else tr_store_set<R, (R::NWARP == 2 ? 1 : 0)>(s.X[blk], regs[t], lane);
    };
    auto load = [&](int t, int k, int blk) {
        const int lane = t & 31;
        if (k == 0) tr_load_set<R, 0>(s.X[blk], regs[t], lane); else tr_load_set<R, (R::NWARP == 2 ? 1 : 0)>(s.X[blk], regs[t], lane);
    };
    if (R::NWARP == 2) {
        for (int t = 0; t < W; ++t) { const int wq = t >> 5; store(t, 1 - wq, 1 - wq); }
        for (int t = 0; t < W; ++t) { const int wq = t >> 5; load(t, 1 - wq, wq); }
    }
    for (int t = 0; t < W; ++t) { const int wq = t >> 5; store(t, wq, wq); }
    for (int t = 0; t < W; ++t) { const int wq = t >> 5; load(t, wq, wq); }
}

template <class R, bool F32, bool PAD = false>
static int run_rows(RParams p) {
    constexpr int W = R::W;
    RSmem<R>* sp = new RSmem<R>();
    RSmem<R>& s = *sp;
    std::vector<RRegs<R>> regs(W), snap(W);
    for (int unit = 0; unit < p.n_units; ++unit) {
        const RUnit un = decode_unit(p, unit);
        for (int f = un.f0; f <= un.f1; ++f) {
            const bool have_prev = f > un.f0;
            if (F32 && PAD) {
                // padded float32 mode: one (PFW floats x ny rows) box per window from the 16-byte boundary below its start
                int xoff[2];
                const int width = p.pitch / 4;
                for (int w = 0; w < 2; ++w) {
                    const int xa = un.x0[w] & ~3;
                    xoff[w] = un.x0[w] - xa;
                    float* dst = reinterpret_cast<float*>(s.tile() + w * R::PFWIN);
                    for (int row = 0; row < p.ny; ++row)
                        for (int b = 0; b < R::PFW; ++b)
                            dst[row * R::PFW + b] = xa + b < width
                                ? reinterpret_cast<const float*>(p.frames + (long long)f * p.frame_stride + (long long)(un.y0[w] + row) * p.pitch)[xa + b] : 0.f;
                }
                memset(s.red, 0, sizeof(s.red));
                for (int t = 0; t < W; ++t) { rows_f1_pad<R>(s, regs[t], t, p, 0, xoff[0]); rows_f1_pad<R>(s, regs[t], t, p, 1, xoff[1]); }
                for (int t = 0; t < W; ++t) { rows_f2_pad<R>(s, regs[t], t, p, 0); rows_f2_pad<R>(s, regs[t], t, p, 1); }
                for (int t = 0; t < W; ++t) rows_f3_pad<R>(s, regs[t], t, p);
            } else if (!F32) {
            // "TMA": fill the tile from frame f (swizzled exact boxes, or 16-byte wider boxes from the boundary below)
            const bool aligned = !PAD && (p.sx % 16) == 0;
            int xoff[2] = {0, 0};
            for (int w = 0; w < 2; ++w) {
                if (aligned) {
                    for (int row = 0; row < W; ++row)
                        for (int j = 0; j < W / 16; ++j)
                            memcpy(s.tile() + tile_chunk_offset<W>(w, row, j),
                                   p.frames + (long long)f * p.frame_stride + (long long)(un.y0[w] + row) * p.pitch + un.x0[w] + 16 * j, 16);
                } else {
                    const int xa = un.x0[w] & ~15;
                    xoff[w] = un.x0[w] - xa;
                    for (int row = 0; row < W; ++row)
                        for (int b = 0; b < R::WB; ++b)
                            s.tile()[(w * W + row) * R::WB + b] =
                                (xa + b < p.pitch && (!PAD || un.y0[w] + row < p.height))
                                    ? p.frames[(long long)f * p.frame_stride + (long long)(un.y0[w] + row) * p.pitch + xa + b] : 0;
                }
            }
            memset(s.red, 0, sizeof(s.red));
            if (PAD) {
                for (int t = 0; t < W; ++t) rows_p1_pad<R>(s, regs[t], t, p, xoff[0], xoff[1]);
                for (int t = 0; t < W; ++t) rows_p2_pre_pad<R>(s, regs[t], t, p);
            } else {
            for (int t = 0; t < W; ++t) { if (aligned) rows_p1<R, true>(s, regs[t], t); else rows_p1<R, false>(s, regs[t], t, xoff[0], xoff[1]); }
            for (int t = 0; t < W; ++t) rows_p2_pre<R>(s, regs[t], t, p.clip_norm);
            }
            } else {
                // float32: 128-byte-wide boxes, SWIZZLE_128B; one window per TMA phase for 64x64, both for 32x32
                auto fill = [&](int w, int toff) {
                    for (int h = 0; h < W / 32; ++h)
                        for (int row = 0; row < W; ++row)
                            for (int j = 0; j < 8; ++j)
                                memcpy(s.tile() + toff + h * R::FBOX + row * 128 + ((j ^ (row & 7)) << 4),
                                       p.frames + (long long)f * p.frame_stride + (long long)(un.y0[w] + row) * p.pitch + 4 * (un.x0[w] + 32 * h + 4 * j), 16);
                };
                memset(s.red, 0, sizeof(s.red));
                fill(0, 0);
                if (R::F_PHASES == 1) fill(1, R::FWIN);
                for (int t = 0; t < W; ++t) { rows_f1<R>(s, regs[t], t, 0, 0); if (R::F_PHASES == 1) rows_f1<R>(s, regs[t], t, 1, R::FWIN); }
                if (R::F_PHASES == 2) {
                    for (int t = 0; t < W; ++t) rows_f2<R>(s, regs[t], t, 0);
                    fill(1, 0);
                    for (int t = 0; t < W; ++t) rows_f1<R>(s, regs[t], t, 1, 0);
                    for (int t = 0; t < W; ++t) rows_f2<R>(s, regs[t], t, 1);
                } else {
                    for (int t = 0; t < W; ++t) { rows_f2<R>(s, regs[t], t, 0); }
                    for (int t = 0; t < W; ++t) { rows_f2<R>(s, regs[t], t, 1); }
                }
                for (int t = 0; t < W; ++t) rows_f3<R>(s, regs[t], t, p.clip_norm);
            }
            for (int t = 0; t < W; ++t) fft_reg<W, 0>(regs[t].v);
            emul_transpose<R, true>(s, regs);
            for (int t = 0; t < W; ++t) fft_reg<W, 0>(regs[t].v);
            snap = regs;
            for (int ky = 0; ky <= W / 2; ++ky) {
                for (int t = 0; t < W; ++t) {
                    const int pt = (t & ~31) | partner_lane_of<W>(t);
                    regs[t].r0 = regs[t].r1 = make_float2(0.f, 0.f);
                    cross_step_a<R, PAD>(s, regs[t], t, ky, snap[pt].v[(W - ky) % W], have_prev, regs[t].r0, regs[t].r1, &p);
                }
                if (have_prev && ky != 0 && ky != W / 2) {
                    std::vector<float2> q(W);
                    for (int t = 0; t < W; ++t) q[t] = cross_mirror(regs[t].r0, regs[t].r1);
                    for (int t = 0; t < W; ++t) {
                        const int pt = (t & ~31) | partner_lane_of<W>(t);
                        cross_step_b<R>(regs[t], ky, q[pt]);
                    }
                }
            }
            if (have_prev) {
                for (int t = 0; t < W; ++t) fft_reg<W, 0>(regs[t].v);
                emul_transpose<R, false>(s, regs);
                for (int k = 0; k < R::NWARP; ++k) for (int q = 4; q < 8; ++q) s.red[k][q] = 0;
                for (int t = 0; t < W; ++t) {
                    const bool d0 = regs[t].half_alpha_prev[0] == 0.f || regs[t].half_alpha_new[0] == 0.f;
                    const bool d1 = regs[t].half_alpha_prev[1] == 0.f || regs[t].half_alpha_new[1] == 0.f;
                    fft_reg<W, 0>(regs[t].v);
                    rows_p5_post<R, PAD>(s, regs[t], t, d0, d1, &p);
                }
                for (int k = 0; k < R::NWARP; ++k) s.redk[k][0] = s.redk[k][1] = ~0ull;
                for (int t = 0; t < W; ++t) rows_p6<R, PAD>(s, regs[t], t, &p);
                for (int t = 0; t < W; ++t) rows_dump_planes<R, PAD>(regs[t], t, p, un, f - 1);
                for (int t = 0; t < W; ++t) rows_p7<R, PAD>(s, regs[t], t, &p);
                for (int t = 0; t < W; ++t) rows_p8<R, PAD>(s, regs[t], t, p, un, f - 1);
            }
            for (int t = 0; t < W; ++t) {
                regs[t].half_alpha_prev[0] = regs[t].half_alpha_new[0];
                regs[t].half_alpha_prev[1] = regs[t].half_alpha_new[1];
            }
        }
    }
    delete sp;
    return 0;
}

static int emul_rows_any(const unsigned char* frames, int is_f32, int n_frames, int H, int W, int win, int ovl, int run_len, int clip_norm,
                         int border_nan, float eps, const unsigned char* keep, float* u, float* v, float* cmax, float* s2n, float* planes) {
    RParams p;
    memset(&p, 0, sizeof(p));
    p.frames = frames; p.pitch = W * (is_f32 ? 4 : 1); p.frame_stride = (long long)H * p.pitch;
    p.n_rows = (H - win) / (win - ovl) + 1; p.n_cols = (W - win) / (win - ovl) + 1;
    p.sy = p.sx = win - ovl; p.n_pairs = n_frames - 1;
    p.run_len = run_len > 0 && run_len < p.n_pairs ? run_len : p.n_pairs;
    const int nw = p.n_rows * p.n_cols;
    p.n_units = ((nw + 1) / 2) * ((p.n_pairs + p.run_len - 1) / p.run_len);
    p.clip_norm = clip_norm; p.border_nan = border_nan; p.gauss_eps = eps; p.keep = keep;
    p.u = u; p.v = v; p.cmax = cmax; p.s2n = s2n; p.planes = planes;
    if (win == 64) return is_f32 ? run_rows<RCfg<64>, true>(p) : run_rows<RCfg<64>, false>(p);
    if (win == 32) return is_f32 ? run_rows<RCfg<32>, true>(p) : run_rows<RCfg<32>, false>(p);
    return -1;
}
// padded mode: any uint8 window up to 32 px per side, any stride (piv_rows.cuh "Padded mode")
static int emul_rows_pad_any(const unsigned char* frames, int is_f32, int n_frames, int H, int W, int wy, int wx, int oy, int ox, int run_len,
                             int clip_norm, int border_nan, float eps, const unsigned char* keep, float* u, float* v, float* cmax,
                             float* s2n, float* planes) {
    RParams p;
    memset(&p, 0, sizeof(p));
    p.frames = frames; p.pitch = W * (is_f32 ? 4 : 1); p.frame_stride = (long long)H * p.pitch; p.height = H;
    p.n_rows = (H - wy) / (wy - oy) + 1; p.n_cols = (W - wx) / (wx - ox) + 1;
    p.sy = wy - oy; p.sx = wx - ox; p.n_pairs = n_frames - 1;
    p.run_len = run_len > 0 && run_len < p.n_pairs ? run_len : p.n_pairs;
    const int nw = p.n_rows * p.n_cols;
    p.n_units = ((nw + 1) / 2) * ((p.n_pairs + p.run_len - 1) / p.run_len);
    p.clip_norm = clip_norm; p.border_nan = border_nan; p.gauss_eps = eps; p.keep = keep;
    p.u = u; p.v = v; p.cmax = cmax; p.s2n = s2n; p.planes = planes;
    p.ny = wy; p.nx = wx;
    const int m = wy > wx ? wy : wx;
    const int P = 2 * m <= 32 ? 32 : 64;
    if (2 * m > 64) return -1;
    p.pad_scale = (float)(1.0 / ((double)P * P * wy * wx));
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k <= P / 2; ++k) { const double th = two_pi * (double)((k * wy) % P) / P; p.pad_ty[k] = make_float2((float)(1.0 + cos(th)), (float)(-sin(th))); }
    for (int k = 0; k < P; ++k) { const double th = two_pi * (double)((k * wx) % P) / P; p.pad_tx[k] = make_float2((float)(1.0 + cos(th)), (float)(-sin(th))); }
    for (int k = 0; k < P / 4; ++k) { const int left = wx - 4 * k; p.pad_mask[k] = left >= 4 ? 0xffffffffu : (left <= 0 ? 0u : ((1u << (8 * left)) - 1u)); }
    for (int x = 0; x < P; ++x) p.pad_cm[x] = x < wx ? 1.f : 0.f;
    // the kernel dumps W x W planes in natural lag order; crop / shift like planes_reorder_kernel
    std::vector<float> nat;
    const long long n_planes = (long long)p.n_pairs * nw;
    if (planes) { nat.assign((size_t)n_planes * P * P, 0.f); p.planes = nat.data(); }
    const int rc = is_f32 ? (P == 32 ? run_rows<RCfg<32>, true, true>(p) : run_rows<RCfg<64>, true, true>(p))
                          : (P == 32 ? run_rows<RCfg<32>, false, true>(p) : run_rows<RCfg<64>, false, true>(p));
    if (planes)
        for (long long pl = 0; pl < n_planes; ++pl)
            for (int iy = 0; iy < wy; ++iy)
                for (int ix = 0; ix < wx; ++ix) {
                    const int hy = wy / 2, hx = wx / 2;
                    const int qy = iy < hy ? iy + wy - hy : iy - hy, qx = ix < hx ? ix + wx - hx : ix - hx;
                    planes[(pl * wy + iy) * wx + ix] = nat[(pl * P + qy) * P + qx];
                }
    return rc;
}
extern "C" int b2piv_emul_rows_pad(const unsigned char* frames, int n_frames, int H, int W, int wy, int wx, int oy, int ox, int run_len,
                                   int clip_norm, int border_nan, float eps, const unsigned char* keep, float* u, float* v, float* cmax,
                                   float* s2n, float* planes) {
    return emul_rows_pad_any(frames, 0, n_frames, H, W, wy, wx, oy, ox, run_len, clip_norm, border_nan, eps, keep, u, v, cmax, s2n, planes);
}
// padded mode, float32 frames (rows_f1_pad / f2_pad / f3_pad)
extern "C" int b2piv_emul_rows_pad_f32(const float* frames, int n_frames, int H, int W, int wy, int wx, int oy, int ox, int run_len,
                                       int clip_norm, int border_nan, float eps, const unsigned char* keep, float* u, float* v, float* cmax,
                                       float* s2n, float* planes) {
    return emul_rows_pad_any((const unsigned char*)frames, 1, n_frames, H, W, wy, wx, oy, ox, run_len, clip_norm, border_nan, eps, keep, u, v,
                             cmax, s2n, planes);
}
extern "C" int b2piv_emul_rows(const unsigned char* frames, int n_frames, int H, int W, int win, int ovl, int run_len, int clip_norm,
                               int border_nan, float eps, const unsigned char* keep, float* u, float* v, float* cmax, float* s2n,
                               float* planes) {
    return emul_rows_any(frames, 0, n_frames, H, W, win, ovl, run_len, clip_norm, border_nan, eps, keep, u, v, cmax, s2n, planes);
}
extern "C" int b2piv_emul_rows_f32(const float* frames, int n_frames, int H, int W, int win, int ovl, int run_len, int clip_norm,
                                   int border_nan, float eps, const unsigned char* keep, float* u, float* v, float* cmax, float* s2n,
                                   float* planes) {
    return emul_rows_any((const unsigned char*)frames, 1, n_frames, H, W, win, ovl, run_len, clip_norm, border_nan, eps, keep, u, v, cmax,
                         s2n, planes);
}


// ---- direct any-size kernel (piv_direct.cuh) -------------------------------------------------------------------------
extern "C" int b2piv_emul_direct(const void* frames, int n_frames, int H, int W, int is_f32, int wy, int wx, int oy, int ox,
                                 int clip_norm, int border_nan, float eps, float* u, float* v, float* cmax, float* s2n, float* planes) {
    Params p;
    memset(&p, 0, sizeof(p));
    p.frames = frames; p.pitch = W * (is_f32 ? 4 : 1); p.frame_stride = (long long)H * p.pitch; p.is_f32 = is_f32;
    p.n_rows = (H - wy) / (wy - oy) + 1; p.n_cols = (W - wx) / (wx - ox) + 1;
    p.sy = wy - oy; p.sx = wx - ox; p.n_pairs = n_frames - 1;
    p.clip_norm = clip_norm; p.border_nan = border_nan; p.gauss_eps = eps;
    p.u = u; p.v = v; p.cmax = cmax; p.s2n = s2n; p.planes = planes;
    std::vector<unsigned char> mem(direct_smem_bytes(wy, wx) + 64);
    DView s = direct_view(mem.data(), wy, wx);
    const int nw = p.n_rows * p.n_cols;
    for (int pair = 0; pair < p.n_pairs; ++pair)
        for (int w = 0; w < nw; ++w) {
            memset(s.red, 0, 64 * sizeof(unsigned long long));
            for (int t = 0; t < DNT; ++t) direct_load(s, t, p, pair, w);
            for (int t = 0; t < DNT; ++t) direct_center(s, t, p);
            for (int t = 0; t < DNT; ++t) direct_correlate(s, t, p, pair, w);
            for (int t = 0; t < DNT; ++t) direct_peak(s, t, p, pair, w);
        }
    return 0;
}

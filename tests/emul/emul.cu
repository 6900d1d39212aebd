// TEST-ONLY host emulator: runs the kernel's barrier-delimited phases (pyorc_b200/csrc/piv_core.cuh) on the CPU,
// "thread by thread, phase by phase", to validate index mathematics without a GPU.  Never loaded by pyorc_b200.
#include "../../pyorc_b200/csrc/piv_core.cuh"
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
#define CASE(Y, X, T) if (wy == Y && wx == X) return nwin == 2 ? run<Cfg<Y, X, T, 2>>(p) : run<Cfg<Y, X, T, 1>>(p);
    CASE(16, 16, 64) CASE(32, 32, 128) CASE(64, 64, 256) CASE(128, 128, 256) CASE(32, 64, 128) CASE(64, 32, 128)
    return -1;
}

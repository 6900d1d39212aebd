"""GPU parity: the CUDA engine (through the C ABI) against the float64 oracle on identical seeded inputs.

Tolerances (fp32 engine vs float64 oracle; SURVEY.md §8d, tightened after measurement):
  correlation planes  abs <= 5e-6          corr_max abs <= 5e-6       s2n rel <= 2e-5
  u, v (where integer peaks agree)         abs <= 2e-3 px, RMSE <= 2e-4 px
  integer-peak agreement >= 99.9 %;  NaN masks identical.
"""
import numpy as np
import pytest

from oracle import ffpiv_oracle as O
from pyorc_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from pyorc_b200.engine import Engine

    e = Engine(0)
    yield e
    e.close()


def compare(engine, imgs, ws, ov, clip, check_planes=True, signal_threshold=None, variant=0, run_len=0):
    O.CLIP_NORMALIZED = bool(clip)
    engine.set_option("clip_normalized", float(clip))
    engine.set_option("kernel_variant", float(variant))   # 0 auto, 1 generic smem kernel, 2 row-per-thread TMA kernel
    engine.set_option("run_len", float(run_len))
    nr, nc = O.get_array_shape(imgs.shape[-2:], ws, ov)
    u, v, c, s = O.uv_timestep(imgs, nc, nr, ws, ov, signal_threshold=signal_threshold)
    gu, gv, gc, gs = engine.pairs(imgs, ws, ov, signal_threshold=signal_threshold)
    assert gu.shape == u.shape == (imgs.shape[0] - 1, nr, nc)
    assert np.array_equal(np.isnan(gu), np.isnan(u)), "NaN mask of u differs"
    assert np.array_equal(np.isnan(gc), np.isnan(c)), "NaN mask of corr differs"
    fin = np.isfinite(u) & np.isfinite(gu)
    same_peak = np.abs(np.round(gu) - np.round(u)) + np.abs(np.round(gv) - np.round(v)) < 0.5
    agree = same_peak[fin].mean() if fin.any() else 1.0
    # near-tied peaks can resolve differently in fp32; tiny windows (noisy planes) have more of them
    assert agree >= (0.999 if min(ws) >= 16 else 0.995), f"integer peak agreement {agree}"
    m = fin & same_peak
    if m.any():
        assert np.abs(gu[m] - u[m]).max() <= 2e-3 and np.abs(gv[m] - v[m]).max() <= 2e-3
        assert np.sqrt(np.mean((gu[m] - u[m]) ** 2)) <= 2e-4 and np.sqrt(np.mean((gv[m] - v[m]) ** 2)) <= 2e-4
    okc = np.isfinite(c)
    assert np.abs(gc[okc] - c[okc]).max() <= 5e-6
    oks = np.isfinite(s) & (s != 0)
    assert np.array_equal(np.isnan(gs), np.isnan(s))
    assert (np.abs(gs[oks] - s[oks]) / np.abs(s[oks])).max() <= 2e-5
    if check_planes:
        _, _, corr = O.cross_corr(imgs, ws, ov, signal_threshold=signal_threshold)
        gp = engine.corr_planes(imgs, ws, ov, signal_threshold=signal_threshold)
        assert np.array_equal(np.isnan(gp), np.isnan(corr))
        assert np.nanmax(np.abs(gp - corr)) <= 5e-6
    return float(np.sqrt(np.mean((gu[m] - u[m]) ** 2))) if m.any() else 0.0


CASES = [
    ((64, 64), (32, 32), (3, 270, 400)),
    ((32, 32), (16, 16), (3, 150, 210)),
    ((32, 32), (24, 24), (3, 100, 140)),
    ((16, 16), (8, 8), (3, 70, 90)),
    ((128, 128), (64, 64), (2, 300, 420)),
    ((32, 64), (16, 32), (3, 130, 260)),
    ((64, 32), (48, 24), (3, 150, 110)),
    ((64, 128), (32, 64), (2, 200, 400)),
    ((128, 64), (64, 32), (2, 400, 200)),
]


@pytest.mark.parametrize("ws,ov,shape", CASES)
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
@pytest.mark.parametrize("clip", [1, 0])
def test_pairs_match_oracle(engine, ws, ov, shape, dtype, clip):
    imgs = synth.particle_frames(*shape, dtype=dtype)
    compare(engine, imgs, ws, ov, clip)


ROWS_CASES = [
    ((64, 64), (32, 32), (4, 270, 400), 0),
    ((64, 64), (32, 32), (6, 200, 304), 2),    # runs of 2 pairs: unit boundaries inside the stack
    ((64, 64), (48, 48), (3, 150, 176), 1),
    ((32, 32), (16, 16), (5, 150, 208), 0),
    ((32, 32), (0, 0), (4, 100, 144), 3),      # no overlap: stride 32
    ((64, 64), (16, 16), (3, 200, 256), 0),    # stride 48
    ((64, 64), (32, 32), (3, 64 * 2 + 7, 64 * 3 + 16), 0),   # 3 x 5 windows: odd count -> last unit has one window
]


@pytest.mark.parametrize("ws,ov,shape,run_len", ROWS_CASES)
@pytest.mark.parametrize("clip", [1, 0])
def test_rows_kernel_matches_oracle(engine, ws, ov, shape, run_len, clip):
    """The row-per-thread TMA kernel (forced), including forward-spectrum sharing across consecutive pairs."""
    imgs = synth.particle_frames(*shape, dtype=np.uint8)
    imgs[:, :40, :50] = 0
    compare(engine, imgs, ws, ov, clip, variant=2, run_len=run_len)


def test_rows_and_generic_kernels_agree(engine):
    imgs = synth.particle_frames(5, 270, 400, dtype=np.uint8)
    engine.set_option("clip_normalized", 1.0)
    engine.set_option("run_len", 0.0)
    engine.set_option("kernel_variant", 1.0)
    a = engine.pairs(imgs, (64, 64), (32, 32))
    engine.set_option("kernel_variant", 2.0)
    b = engine.pairs(imgs, (64, 64), (32, 32))
    engine.set_option("kernel_variant", 0.0)
    for x, y in zip(a, b):
        assert np.array_equal(np.isnan(x), np.isnan(y))
        assert np.nanmax(np.abs(x - y) / (1 + np.abs(x))) < 2e-5


@pytest.mark.parametrize("ws,ov,shape,run_len", [((32, 32), (24, 24), (4, 100, 144), 3), ((64, 64), (40, 40), (3, 160, 208), 0),
                                                 ((32, 32), (20, 20), (3, 90, 128), 0), ((64, 64), (56, 56), (3, 100, 128), 2)])
def test_rows_kernel_window_starts_not_16_byte_aligned(engine, ws, ov, shape, run_len):
    """x strides 8 / 24 / 12 / 8: TMA boxes start at the 16-byte boundary below, rows are read at an offset
    (BASELINE configs[2] geometry: 32x32 at 75 % overlap)."""
    imgs = synth.particle_frames(*shape, dtype=np.uint8)
    imgs[:, :40, :50] = 0
    compare(engine, imgs, ws, ov, 0, variant=2, run_len=run_len)
    engine.set_option("kernel_variant", 0.0)


ROWS128_CASES = [
    ((128, 128), (64, 64), (4, 300, 420), 0),       # 3 x 5 windows: odd count -> last unit has one window
    ((128, 128), (64, 64), (6, 256 + 9, 384 + 16), 2),   # runs of 2 pairs: unit boundaries inside the stack
    ((128, 128), (32, 32), (3, 330, 450), 0),       # stride 96
    ((128, 128), (112, 112), (3, 170, 200), 1),     # stride 16: heavy overlap, many windows per TMA row band
]


@pytest.mark.parametrize("ws,ov,shape,run_len", ROWS128_CASES)
@pytest.mark.parametrize("clip", [1, 0])
def test_rows128_kernel_matches_oracle(engine, ws, ov, shape, run_len, clip):
    """128x128 windows on the polyphase row-per-thread kernel (forced): four 64x64 sub-groups coupled in the cross phase."""
    imgs = synth.particle_frames(*shape, dtype=np.uint8)
    imgs[:, :140, :150] = 0                        # a dead window (zero variance) next to live ones
    compare(engine, imgs, ws, ov, clip, variant=2, run_len=run_len)
    engine.set_option("kernel_variant", 0.0)


def test_rows128_and_generic_kernels_agree_and_signal_threshold(engine):
    imgs = synth.particle_frames(4, 300, 420, dtype=np.uint8)
    imgs[:, 150:, 200:] = 0
    engine.set_option("clip_normalized", 0.0)
    engine.set_option("run_len", 0.0)
    engine.set_option("kernel_variant", 1.0)
    a = engine.pairs(imgs, (128, 128), (64, 64))
    engine.set_option("kernel_variant", 2.0)
    b = engine.pairs(imgs, (128, 128), (64, 64))
    engine.set_option("kernel_variant", 0.0)
    for x, y in zip(a, b):
        assert np.array_equal(np.isnan(x), np.isnan(y))
        assert np.nanmax(np.abs(x - y) / (1 + np.abs(x))) < 2e-5
    compare(engine, imgs, (128, 128), (64, 64), 0, signal_threshold=0.6, variant=2)
    engine.set_option("kernel_variant", 0.0)


@pytest.mark.parametrize("shape,run_len", [((4, 300, 420), 0), ((6, 280, 432), 2)])
def test_rows128_kernel_float32_frames(engine, shape, run_len):
    """float32 frames (pyorc's time_diff / smooth / edge_detect output) at 128x128 on the polyphase kernel: 32-float swizzled TMA
    boxes into the spectrum blocks, two-pass float moments - against the oracle (planes included) and the shared-memory kernel."""
    imgs = synth.particle_frames(*shape, dtype=np.float32)
    imgs[:, :128, :140] = 0
    imgs[1:] -= 0.25 * imgs[:-1]
    compare(engine, imgs, (128, 128), (64, 64), 0, variant=2, run_len=run_len)
    a = engine.pairs(imgs, (128, 128), (64, 64))
    engine.set_option("kernel_variant", 1.0)
    b = engine.pairs(imgs, (128, 128), (64, 64))
    engine.set_option("kernel_variant", 0.0)
    engine.set_option("run_len", 0.0)
    assert np.nanmax(np.abs(a[2] - b[2])) <= 1e-5


def test_rows128_kernel_needs_16_byte_strides(engine):
    imgs = synth.particle_frames(3, 300, 420, dtype=np.uint8)
    engine.set_option("kernel_variant", 2.0)
    with pytest.raises(NotImplementedError):
        engine.pairs(imgs, (128, 128), (60, 60))          # stride 68
    engine.set_option("kernel_variant", 0.0)
    compare(engine, imgs, (128, 128), (60, 60), 0)      # auto: shared-memory kernel
    compare(engine, imgs.astype(np.float32), (128, 128), (64, 64), 0)   # float32 frames: polyphase kernel as well (round 2)
    compare(engine, imgs.astype(np.float32), (128, 128), (62, 62), 0)   # float32, stride 66 floats (not 16-byte aligned): shared-memory kernel


def test_rows_kernel_refuses_stride_not_multiple_of_4(engine):
    imgs = synth.particle_frames(3, 100, 144, dtype=np.uint8)
    engine.set_option("kernel_variant", 2.0)
    with pytest.raises(NotImplementedError):
        engine.pairs(imgs, (32, 32), (22, 22))            # stride 10
    engine.set_option("kernel_variant", 0.0)
    compare(engine, imgs, (32, 32), (22, 22), 0)        # auto: generic kernel


def test_rows_kernel_takes_any_frame_width_from_the_host(engine):
    """pitch 203 B: TMA needs 16-byte strides, so the engine's device copy of host frames is pitched (a caller-owned
    device tensor with that pitch is refused, see test_rows_kernels_on_device_tensors_need_an_aligned_pitch)."""
    import torch

    imgs = synth.particle_frames(3, 150, 203, dtype=np.uint8)
    compare(engine, imgs, (64, 64), (32, 32), 1, variant=2)
    engine.set_option("kernel_variant", 2.0)
    with pytest.raises(NotImplementedError):
        engine.pairs(torch.from_numpy(imgs).cuda(), (64, 64), (32, 32))
    engine.set_option("kernel_variant", 0.0)
    compare(engine, imgs.astype(np.float32), (64, 64), (32, 32), 0, variant=2)   # float32 rows: 812-byte rows -> 816
    engine.set_option("kernel_variant", 0.0)


DIRECT_CASES = [
    ((10, 10), (5, 5), (3, 100, 120)),      # pyorc's golden-test window
    ((26, 26), (12, 12), (3, 150, 180)),    # camera-config window 25 -> 26, overlap int(25/2)
    ((20, 14), (10, 7), (2, 90, 80)),
    ((50, 50), (25, 25), (2, 160, 210)),
    ((9, 11), (4, 5), (2, 50, 60)),         # odd sizes (never produced by pyorc, allowed by the ABI)
    ((64, 64), (32, 32), (2, 140, 200)),    # power of two through the direct kernel (variant 3)
]


@pytest.mark.parametrize("ws,ov,shape", DIRECT_CASES)
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_direct_kernel_any_window_size(engine, ws, ov, shape, dtype):
    imgs = synth.particle_frames(*shape, dtype=dtype)
    imgs[:, :20, :20] = 0
    compare(engine, imgs, ws, ov, 0, variant=3)
    engine.set_option("kernel_variant", 0.0)


PADDED_CASES = [
    ((26, 26), (12, 12), (3, 150, 180)),    # -> 64x64 plane
    ((20, 14), (10, 7), (2, 90, 80)),       # -> 64x32 plane
    ((50, 50), (25, 25), (2, 160, 210)),    # -> 128x128 plane
    ((24, 40), (12, 20), (2, 100, 160)),    # -> 64x128 plane
    ((14, 14), (7, 7), (3, 80, 90)),        # -> 32x32 plane
]


@pytest.mark.parametrize("ws,ov,shape", PADDED_CASES)
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
@pytest.mark.parametrize("variant", [0, 1])   # 0: auto (uint8 windows up to 32 px take the padded row-per-thread kernel); 1: shared-memory kernel
def test_padded_fft_kernel_non_power_of_two_windows(engine, ws, ov, shape, dtype, variant):
    """pyorc's usual windows (26 from a camera-config 25, 20, 50 ...) through the power-of-two FFT kernels, exactly."""
    imgs = synth.particle_frames(*shape, dtype=dtype)
    imgs[:, :20, :20] = 0
    compare(engine, imgs, ws, ov, 0, variant=variant)
    engine.set_option("kernel_variant", 0.0)


ROWS_PAD_CASES = [
    ((26, 26), (12, 12), (4, 150, 192), 0),     # pyorc's default geometry: window 25 -> 26, overlap 12, stride 14
    ((26, 26), (12, 12), (5, 100, 176), 2),     # short runs (first frame of every unit stores nothing)
    ((20, 20), (10, 10), (3, 90, 128), 0),
    ((30, 18), (15, 9), (3, 100, 96), 0),       # rectangular
    ((10, 10), (5, 5), (3, 60, 80), 0),         # -> 32x32 plane
    ((16, 16), (8, 8), (3, 70, 96), 0),         # power of two below the native sizes
    ((32, 32), (15, 15), (3, 100, 112), 0),     # native size with an odd stride (17): byte-granular window starts
    ((22, 22), (11, 11), (3, 77, 91), 0),       # frame width that is not a multiple of 16: the engine's device copy is pitched
]


@pytest.mark.parametrize("ws,ov,shape,run_len", ROWS_PAD_CASES)
@pytest.mark.parametrize("clip", [0, 1])
def test_rows_kernel_padded_mode(engine, ws, ov, shape, run_len, clip):
    """Padded mode of the row-per-thread kernel (kernel_variant 4): zero-padded window, tiling as a spectrum factor, the
    reference's plane read from the wrap-free lags - against the oracle, planes included."""
    imgs = synth.particle_frames(*shape, dtype=np.uint8)
    imgs[:, : ws[0], : ws[1] + 3] = 0
    compare(engine, imgs, ws, ov, clip, variant=4, run_len=run_len)
    engine.set_option("kernel_variant", 0.0)
    engine.set_option("run_len", 0.0)


@pytest.mark.parametrize("ws,ov,shape,run_len", ROWS_PAD_CASES + [((12, 28), (5, 13), (3, 61, 99), 0), ((26, 26), (12, 12), (3, 96, 127), 0)])
@pytest.mark.parametrize("clip", [0, 1])
def test_rows_kernel_padded_mode_float32_frames(engine, ws, ov, shape, run_len, clip):
    """The padded mode on float32 frames - what pyorc's own recipe hands over (normalize -> edge_detect -> minmax ->
    get_piv(window_size=25), examples/ngwerere/ngwerere.yml: float32 frames, 26 x 26 windows): un-swizzled TMA boxes from the
    16-byte boundary below any window start, two-pass moments over the window's pixels, frames with negative values, a dead
    window, frame widths that end inside a box - against the oracle (planes included) and against the shared-memory kernel."""
    imgs = synth.particle_frames(*shape, dtype=np.float32)
    imgs[1:] -= 0.25 * imgs[:-1]
    imgs[:, : ws[0], : ws[1] + 3] = 0
    compare(engine, imgs, ws, ov, clip, variant=4, run_len=run_len)
    assert engine.last_variant == 4
    a = engine.pairs(imgs, ws, ov)
    engine.set_option("kernel_variant", 1.0)
    b = engine.pairs(imgs, ws, ov)
    engine.set_option("kernel_variant", 0.0)
    engine.set_option("run_len", 0.0)
    c = engine.pairs(imgs, ws, ov)                     # auto: the padded rows kernel unless the size is a compiled FFT shape
    for x, y in zip(a[2:], b[2:]):
        assert np.array_equal(np.isnan(x), np.isnan(y)) and np.nanmax(np.abs(x - y)) <= 1e-5 * max(1.0, np.nanmax(np.abs(y)))
    if ws not in ((16, 16), (32, 32)):
        assert engine.last_variant == 4 and all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, c))


ROWS_PAD128_CASES = [
    ((50, 50), (25, 25), (4, 160, 224), 0),     # a 4K user's 50 px window: 128-point plane, 25 x 25 samples per polyphase component
    ((50, 50), (25, 25), (5, 130, 210), 2),     # short runs, frame width that is not a multiple of 16
    ((40, 40), (20, 20), (3, 130, 180), 0),
    ((62, 62), (31, 31), (3, 160, 190), 0),     # odd component size, odd stride
    ((34, 34), (17, 17), (3, 120, 150), 0),     # the smallest size that needs the 128-point plane
    ((64, 48), (32, 24), (3, 170, 200), 0),     # rectangular, one side at the plane's limit
    ((36, 20), (18, 10), (3, 110, 100), 0),     # one side small
    ((20, 44), (10, 22), (3, 90, 160), 0),
]


@pytest.mark.parametrize("ws,ov,shape,run_len", ROWS_PAD128_CASES)
@pytest.mark.parametrize("clip", [0, 1])
def test_rows128_kernel_padded_mode(engine, ws, ov, shape, run_len, clip):
    """Padded mode of the 128-plane polyphase kernel (even windows of 34 .. 64 px): zero-padded window, 2 x 2 tiling as a spectrum
    factor on the polyphase components, the reference's plane read from the wrap-free lags - against the oracle, planes included,
    and against the shared-memory kernel."""
    imgs = synth.particle_frames(*shape, dtype=np.uint8)
    imgs[:, : ws[0], : ws[1] + 3] = 0                  # a dead window and a partly empty one
    compare(engine, imgs, ws, ov, clip, variant=4, run_len=run_len)
    a = engine.pairs(imgs, ws, ov)
    engine.set_option("kernel_variant", 1.0)
    b = engine.pairs(imgs, ws, ov)
    engine.set_option("kernel_variant", 0.0)
    engine.set_option("run_len", 0.0)
    for x, y in zip(a[2:], b[2:]):
        assert np.array_equal(np.isnan(x), np.isnan(y)) and np.nanmax(np.abs(x - y)) <= 1e-5 * max(1.0, np.nanmax(np.abs(y)))


@pytest.mark.parametrize("ws,ov,shape,run_len", ROWS_PAD128_CASES + [((50, 50), (24, 24), (3, 131, 203), 0)])
@pytest.mark.parametrize("clip", [0, 1])
def test_rows128_kernel_padded_mode_float32_frames(engine, ws, ov, shape, run_len, clip):
    """The padded mode of the 128-plane polyphase kernel on float32 frames (pyorc's edge_detect / smooth / time_diff output with a
    window of 34 .. 64 px, e.g. a 4K user's 50 px window): one 68-float x 64-row un-swizzled TMA box per window from the 16-byte
    boundary below its start, two-pass float moments over the window's pixels, frames with negative values, a dead window, odd
    strides and frame widths - against the oracle (planes included) and against the shared-memory kernel."""
    imgs = synth.particle_frames(*shape, dtype=np.float32)
    imgs[1:] -= 0.25 * imgs[:-1]
    imgs[:, : ws[0], : ws[1] + 3] = 0
    compare(engine, imgs, ws, ov, clip, variant=4, run_len=run_len)
    assert engine.last_variant == 4
    a = engine.pairs(imgs, ws, ov)
    engine.set_option("kernel_variant", 1.0)
    b = engine.pairs(imgs, ws, ov)
    assert engine.last_variant == 1
    engine.set_option("kernel_variant", 0.0)
    engine.set_option("run_len", 0.0)
    c = engine.pairs(imgs, ws, ov)                     # auto: the padded polyphase kernel
    assert engine.last_variant == 4 and all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, c))
    for x, y in zip(a[2:], b[2:]):
        assert np.array_equal(np.isnan(x), np.isnan(y)) and np.nanmax(np.abs(x - y)) <= 1e-5 * max(1.0, np.nanmax(np.abs(y)))


def test_rows_kernels_on_device_tensors_need_an_aligned_pitch(engine):
    """Host frames are copied into a 16-byte pitched buffer by the engine (any width qualifies); a caller-owned device
    tensor is used in place, so an odd pitch is refused by the TMA kernels and taken by the shared-memory kernel."""
    import torch

    O.CLIP_NORMALIZED = False
    engine.set_option("clip_normalized", 0.0)
    imgs = synth.particle_frames(3, 77, 91, dtype=np.uint8)
    d = torch.from_numpy(imgs).cuda()
    engine.set_option("kernel_variant", 4.0)
    with pytest.raises(NotImplementedError):
        engine.pairs(d, (22, 22), (11, 11))
    engine.set_option("kernel_variant", 0.0)
    got = [t.cpu().numpy() for t in engine.pairs(d, (22, 22), (11, 11))]
    host = engine.pairs(imgs, (22, 22), (11, 11))           # padded rows kernel on the pitched copy
    nr, nc = O.get_array_shape(imgs.shape[1:], (22, 22), (11, 11))
    u, v, c, s_ = O.uv_timestep(imgs, nc, nr, (22, 22), (11, 11))
    for g in (got, host):
        ok = np.isfinite(u) & np.isfinite(g[0])
        assert np.array_equal(np.isnan(g[0]), np.isnan(u))
        assert np.abs(g[0][ok] - u[ok]).max() <= 2e-3 and np.abs(g[2] - c).max() <= 5e-6


def test_odd_window_count_and_ragged_edges(engine):
    # 3 x 5 windows (odd count -> last work item holds a single window); frame not a multiple of the stride
    imgs = synth.particle_frames(3, 64 * 2 + 7, 64 * 3 + 13, dtype=np.uint8)
    compare(engine, imgs, (64, 64), (32, 32), 1)


def test_dead_and_saturated_windows(engine):
    imgs = synth.particle_frames(3, 200, 300, dtype=np.uint8)
    imgs[:, :80, :100] = 0        # zero-signal windows: std == 0 -> zero plane -> s2n = 0/0 = NaN
    imgs[:, 120:, 200:] = 255     # saturated constant
    compare(engine, imgs, (64, 64), (32, 32), 1)
    compare(engine, imgs, (32, 32), (16, 16), 0)


def test_signal_threshold(engine):
    imgs = synth.particle_frames(3, 200, 300, dtype=np.uint8)
    imgs[imgs < 60] = 0
    imgs[:, :100, :150] = 0
    compare(engine, imgs, (64, 64), (32, 32), 1, signal_threshold=0.2)


def test_device_resident_path(engine):
    import torch

    engine.set_option("kernel_variant", 0.0)
    engine.set_option("run_len", 0.0)
    imgs = synth.particle_frames(4, 270, 400, dtype=np.uint8)
    engine.set_option("clip_normalized", 1.0)
    hu, hv, hc, hs = engine.pairs(imgs, (64, 64), (32, 32))
    d = torch.from_numpy(imgs).cuda()
    du, dv, dc, ds = engine.pairs(d, (64, 64), (32, 32))
    torch.cuda.synchronize()
    for a, b in ((hu, du), (hv, dv), (hc, dc), (hs, ds)):
        assert np.array_equal(a, b.cpu().numpy(), equal_nan=True)


def test_errors(engine):
    imgs = synth.particle_frames(2, 100, 100, dtype=np.uint8)
    big = synth.particle_frames(2, 200, 200, dtype=np.uint8)
    with pytest.raises(NotImplementedError):
        engine.pairs(big, (130, 130), (64, 64))        # beyond 128 px per side
    with pytest.raises(ValueError):
        engine.pairs(imgs[:1], (64, 64), (32, 32))
    with pytest.raises(ValueError):
        engine.pairs(imgs, (128, 128), (64, 64))  # frame smaller than window


BIG_CASES = [   # a side of 65 .. 128 px that is not a compiled FFT shape: large-window direct kernel (pyorc accepts any even size)
    ((66, 66), (33, 33), (3, 200, 240)),
    ((100, 100), (50, 50), (2, 260, 320)),
    ((126, 126), (62, 62), (2, 260, 400)),
    ((96, 40), (48, 20), (2, 250, 130)),
    ((70, 128), (30, 64), (2, 200, 330)),
]


@pytest.mark.parametrize("ws,ov,shape", BIG_CASES)
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_large_windows_between_the_fft_sizes(engine, ws, ov, shape, dtype):
    imgs = synth.particle_frames(*shape, dtype=dtype)
    imgs[:, : ws[0], : ws[1]] = 3              # a dead window
    compare(engine, imgs, ws, ov, 0, variant=0)
    compare(engine, imgs, ws, ov, 1, variant=0, check_planes=False, signal_threshold=0.999)


ENS_CASES = [
    ((100, 100), (50, 50), (5, 260, 320), 0.2, 2.0),
    ((64, 64), (32, 32), (6, 200, 304), 0.3, 2.0),
    ((32, 32), (16, 16), (6, 100, 160), 0.2, 3.0),
    ((10, 10), (5, 5), (5, 60, 80), 0.0, 0.0),
    ((26, 26), (12, 12), (5, 90, 120), 0.2, 1.5),
    ((128, 128), (64, 64), (4, 300, 420), 0.2, 3.0),
    ((50, 50), (25, 25), (5, 160, 224), 0.2, 2.0),       # padded mode of the 128-plane kernel
    ((36, 44), (18, 22), (5, 120, 150), 0.2, 2.0),
]


@pytest.mark.parametrize("ws,ov,shape,corr_min,s2n_min", ENS_CASES)
@pytest.mark.parametrize("variant", [0, 1])     # 0: row-per-thread kernel where eligible (64x64, 32x32); 1: shared-memory FFT kernel
def test_ensemble_mode_matches_oracle(engine, ws, ov, shape, corr_min, s2n_min, variant):
    """Ensemble correlation (pyorc/velocimetry/ffpiv.py:182-376): thresholds + plane sums on the device, two chunks."""
    O.CLIP_NORMALIZED = False
    engine.set_option("clip_normalized", 0.0)
    engine.set_option("kernel_variant", float(variant))
    imgs = synth.particle_frames(*shape, dtype=np.uint8)
    imgs[:, : ws[0], : ws[1]] = 0                      # one dead window
    imgs[2] = synth.particle_frames(1, shape[1], shape[2], dtype=np.uint8, seed=7)[0]   # a decorrelated frame -> masked pairs
    nr, nc = O.get_array_shape(shape[1:], ws, ov)
    # move each threshold into the widest gap of the actual per-pair metrics near its nominal value, so that no pair
    # sits within rounding distance of a threshold (fp32 engine vs float64 oracle would then mask differently)
    _, _, c_all, s_all = O.uv_timestep(imgs, nc, nr, ws, ov)

    def gap_threshold(vals, nominal):
        if nominal <= 0:
            return 0.0
        v = np.sort(vals[np.isfinite(vals)])
        v = v[(v > 0.5 * nominal) & (v < 2.0 * nominal)]
        if v.size < 2:
            return float(nominal)
        k = int(np.argmax(np.diff(v)))
        return float(0.5 * (v[k] + v[k + 1]))

    corr_min = gap_threshold(c_all.ravel(), corr_min)
    s2n_min = gap_threshold(s_all.ravel(), s2n_min)
    half = shape[0] // 2 + 1
    chunks = [imgs[:half], imgs[half - 1 :]]
    ens = O.Ensemble(nr, nc, ws, ov, corr_min=corr_min, s2n_min=s2n_min, count_min=0.2)
    engine.ens_begin(shape[1:], ws, ov, np.uint8)
    got_c, got_s = [], []
    for ch in chunks:
        ens.add_chunk(ch)
        c, s_ = engine.ens_add(ch, ws, ov, corr_min=corr_min, s2n_min=s2n_min)
        got_c.append(c); got_s.append(s_)
    for a, b in zip(got_c, ens.corr_chunks):
        assert np.array_equal(a == 0, b == 0), "mask of rejected pairs differs"
        assert np.abs(a - b).max() <= 5e-6
    for a, b in zip(got_s, ens.s2n_chunks):
        assert np.abs(a - b).max() <= 2e-5 * max(1.0, np.abs(b).max())
    u, v, cm, sn = ens.finalize()
    gu, gv, cnt = engine.ens_finish(0.2 * len(chunks))
    assert np.array_equal(cnt, np.asarray(ens.corr_count).reshape(-1))
    u, v = u.reshape(-1), v.reshape(-1)
    assert np.array_equal(np.isnan(gu), np.isnan(u))
    ok = np.isfinite(u)
    same = (np.abs(np.round(gu[ok]) - np.round(u[ok])) + np.abs(np.round(gv[ok]) - np.round(v[ok]))) < 0.5
    assert same.mean() >= 0.99
    assert np.abs(gu[ok][same] - u[ok][same]).max() <= 2e-3 and np.abs(gv[ok][same] - v[ok][same]).max() <= 2e-3


@pytest.mark.parametrize("ws,ov,shape,dtype", [((64, 64), (32, 32), (7, 270, 400), np.float32), ((32, 32), (24, 24), (6, 100, 144), np.uint8),
                                               ((32, 32), (16, 16), (9, 150, 208), np.float32), ((64, 64), (40, 40), (5, 160, 208), np.uint8),
                                               ((128, 128), (64, 64), (5, 300, 432), np.uint8),      # polyphase kernel, ensemble epilogue
                                               ((128, 128), (64, 64), (4, 300, 432), np.float32),    # ... with float32 frames
                                               ((50, 50), (25, 25), (5, 160, 224), np.uint8),        # ... and in padded mode
                                               ((50, 46), (25, 23), (5, 160, 224), np.float32),      # ... padded mode, float32 frames (16-byte pitch: device tensor used in place)
                                               ((26, 26), (12, 12), (6, 96, 127), np.float32),       # padded rows kernel, float32 frames
                                               ((10, 14), (5, 7), (5, 60, 83), np.float32)])
def test_ensemble_rows_kernel_device_frames(engine, ws, ov, shape, dtype):
    """Device-resident chunk in ONE launch (a unit walks all frames and adds its planes to the HBM accumulators): float32
    frames and window starts that are not 16-byte aligned, against the oracle's plane sums (no thresholds, so no pair can
    sit on a threshold) and against the shared-memory kernel."""
    import torch

    O.CLIP_NORMALIZED = False
    engine.set_option("clip_normalized", 0.0)
    imgs = synth.particle_frames(*shape, dtype=dtype)
    imgs[:, : ws[0], : ws[1]] = 0
    nr, nc = O.get_array_shape(shape[1:], ws, ov)
    ens = O.Ensemble(nr, nc, ws, ov, corr_min=0.0, s2n_min=0.0, count_min=0.0)
    ens.add_chunk(imgs)
    ref_sum = np.asarray(ens.corr_sum, dtype=np.float64).reshape(nr * nc, -1)
    d = torch.from_numpy(imgs).cuda()
    res = {}
    for variant in (0, 1):
        engine.set_option("kernel_variant", float(variant))
        engine.ens_begin(shape[1:], ws, ov, dtype)
        c, s_ = engine.ens_add(d, ws, ov, corr_min=0.0, s2n_min=0.0)
        plane, count = engine.ens_accumulators()
        res[variant] = (c.cpu().numpy(), s_.cpu().numpy(), plane.cpu().numpy().copy(), count.cpu().numpy().copy(), engine.ens_finish(0.0))
    engine.set_option("kernel_variant", 0.0)
    for variant in (0, 1):
        c, s_, plane, count, (gu, gv, cnt) = res[variant]
        assert np.abs(c - ens.corr_chunks[0]).max() <= 5e-6
        assert np.abs(plane - ref_sum).max() <= 2e-5
        assert np.array_equal(count, np.asarray(ens.corr_count).reshape(-1))
    assert np.abs(res[0][2] - res[1][2]).max() <= 1e-5
    u, v, _, _ = ens.finalize()
    gu, gv, _ = res[0][4]
    assert np.array_equal(np.isnan(gu), np.isnan(u.reshape(-1)))
    ok = np.isfinite(gu)
    assert np.abs(gu[ok] - u.reshape(-1)[ok]).max() <= 2e-3 and np.abs(gv[ok] - v.reshape(-1)[ok]).max() <= 2e-3


def test_ffpiv_api_cross_corr_and_u_v_displacement(engine):
    """ffpiv's own two-call form (cross_corr -> planes -> u_v_displacement) on the GPU, against the oracle's."""
    from pyorc_b200 import ffpiv_api

    O.CLIP_NORMALIZED = False
    engine.set_option("clip_normalized", 0.0)
    engine.set_option("kernel_variant", 0.0)
    imgs = synth.particle_frames(3, 150, 210, dtype=np.uint8)
    imgs[:, :40, :40] = 0
    ws, ov = (32, 32), (16, 16)
    nr, nc = O.get_array_shape(imgs.shape[-2:], ws, ov)
    x, y, corr = ffpiv_api.cross_corr(imgs, window_size=ws, overlap=ov, search_area_size=ws, normalize=False, verbose=False)
    ox, oy, ocorr = O.cross_corr(imgs, ws, ov)
    assert np.array_equal(x, ox) and np.array_equal(y, oy) and corr.dtype == np.float32
    assert np.nanmax(np.abs(corr - ocorr)) <= 5e-6
    u, v = ffpiv_api.u_v_displacement(ocorr, nr, nc)          # same planes in -> isolates the peak-fit kernel
    ou, ov_ = O.u_v_displacement(ocorr, nr, nc)
    assert u.shape == ou.shape == (2, nr, nc)
    assert np.array_equal(np.isnan(u), np.isnan(ou))
    ok = np.isfinite(ou)
    assert np.abs(u[ok] - ou[ok]).max() <= 1e-4 and np.abs(v[ok] - ov_[ok]).max() <= 1e-4


@pytest.mark.parametrize("ws,ov,shape,run_len", [((64, 64), (32, 32), (4, 270, 400), 0), ((32, 32), (16, 16), (5, 150, 208), 2),
                                                 ((64, 64), (44, 44), (3, 160, 204), 0), ((32, 32), (24, 24), (3, 100, 144), 0)])
def test_rows_kernel_float32_frames(engine, ws, ov, shape, run_len):
    """float32 frames (pyorc's time_diff / smooth / edge_detect output) through the row-per-thread TMA kernel."""
    imgs = synth.particle_frames(*shape, dtype=np.float32)
    imgs[:, :40, :52] = 0
    imgs[1:] -= 0.25 * imgs[:-1]
    compare(engine, imgs, ws, ov, 0, variant=2, run_len=run_len)
    engine.set_option("kernel_variant", 0.0)


def test_peer_push_two_engines_one_device():
    """b2piv_peer_push: the gather as a separate copy kernel.  Two engines (stand-ins for two ranks) push their blocks into both
    gather buffers at their offsets - vector path (element counts divisible by 4) and the scalar tail (odd window counts)."""
    import torch

    from pyorc_b200.engine import Engine

    for shape, ws, ov in (((200, 304), (64, 64), (32, 32)), ((150, 210), (32, 32), (16, 16)), ((131, 197), (30, 30), (15, 15))):
        imgs = synth.particle_frames(7, *shape, dtype=np.uint8)
        d = torch.from_numpy(imgs).cuda()
        with Engine(0) as e0, Engine(0) as e1:
            whole = torch.stack(e0.pairs(d, ws, ov))
            nr, nc = e0.plan(shape, ws, ov, np.uint8)
            e1.plan(shape, ws, ov, np.uint8)
            bufs = [torch.full((4, 6, nr, nc), -7.0, device="cuda") for _ in range(2)]
            ptrs = [b.data_ptr() for b in bufs]
            a = e0.pairs(d[:4], ws, ov)            # pairs 0..2
            b = e1.pairs(d[3:], ws, ov)            # pairs 3..5
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            e0.peer_push(a[0]._base, ptrs, 6, 0, stream=side)
            e1.peer_push(b[0]._base, ptrs, 6, 3, stream=side)
            side.synchronize()
            for buf in bufs:
                assert torch.equal(torch.nan_to_num(buf), torch.nan_to_num(whole))
            with pytest.raises(ValueError):
                e1.peer_push(b[0]._base, ptrs, 6, 4)       # 3 pairs from offset 4 do not fit 6
            with pytest.raises(ValueError):
                e1.peer_push(b[0], ptrs, 6, 3)             # one field, not the [4, pairs, rows, cols] block


def test_fused_peer_gather_two_engines_one_device():
    """b2piv_set_peer_outputs: two engines (stand-ins for two ranks, both on cuda:0) each process half of the frame pairs and
    store their results straight into BOTH gather buffers; each buffer ends up with the whole time axis."""
    import torch

    from pyorc_b200.engine import Engine

    imgs = synth.particle_frames(7, 200, 304, dtype=np.uint8)
    d = torch.from_numpy(imgs).cuda()
    ws, ov = (64, 64), (32, 32)
    with Engine(0) as e0, Engine(0) as e1, Engine(0) as ref:
        for e in (e0, e1, ref):
            e.set_option("clip_normalized", 0.0)
        whole = torch.stack(ref.pairs(d, ws, ov))
        nr, nc = e0.plan((200, 304), ws, ov, np.uint8)
        e1.plan((200, 304), ws, ov, np.uint8)
        bufs = [torch.full((4, 6, nr, nc), -7.0, device="cuda") for _ in range(2)]
        ptrs = [b.data_ptr() for b in bufs]
        e0.set_peer_outputs(ptrs, 6, 0)
        e1.set_peer_outputs(ptrs, 6, 4)
        a = e0.pairs(d[:5], ws, ov)            # pairs 0..3
        b = e1.pairs(d[4:], ws, ov)            # pairs 4..5
        torch.cuda.synchronize()
        for buf in bufs:
            assert torch.equal(torch.nan_to_num(buf), torch.nan_to_num(whole))
        assert torch.equal(torch.nan_to_num(torch.stack(a)), torch.nan_to_num(whole[:, :4]))     # local results still written
        with pytest.raises(ValueError):
            e1.pairs(d[3:], ws, ov)            # 3 pairs from offset 4 do not fit 6
        e0.set_peer_outputs(None, 1, 0)
        bufs[0].fill_(-7.0)
        e0.pairs(d[:5], ws, ov)
        torch.cuda.synchronize()
        assert float(bufs[0].max()) == -7.0    # switched off
        # the 128x128 polyphase kernel and the shared-memory kernel have the same epilogue hook
        for wsz, ovl, variant in (((128, 128), (64, 64), 0), ((64, 64), (32, 32), 1)):
            e0.set_option("kernel_variant", float(variant))
            ref.set_option("kernel_variant", float(variant))
            w2 = torch.stack(ref.pairs(d, wsz, ovl))
            r2, c2 = e0.plan((200, 304), wsz, ovl, np.uint8)
            buf = torch.full((4, 8, r2, c2), -7.0, device="cuda")
            e0.set_peer_outputs([buf.data_ptr()], 8, 2)
            e0.pairs(d, wsz, ovl)
            torch.cuda.synchronize()
            assert torch.equal(torch.nan_to_num(buf[:, 2:]), torch.nan_to_num(w2)) and float(buf[:, :2].max()) == -7.0
            e0.set_peer_outputs(None, 1, 0)


def test_pageable_and_pinned_host_frames_give_identical_results(engine):
    """Ordinary numpy memory is staged through the engine's page-locked ring by copy threads (any thread count, any frame
    width - 203 px rows are re-pitched on the way), page-locked memory goes straight to the device: same bits."""
    imgs = synth.particle_frames(23, 190, 203, dtype=np.uint8)
    pinned = engine.pinned_empty(imgs.shape, np.uint8)
    pinned[...] = imgs
    engine.set_option("clip_normalized", 0.0)
    engine.set_option("kernel_variant", 0.0)
    ref = engine.pairs(pinned, (32, 32), (16, 16))
    # (threads, H2D chunks, stage_mode, slice KB, ring groups, non-temporal stores): the Stager of csrc/stager.h with slices from
    # 4 KB (hundreds of groups for these 0.9 MB) to larger than the call, and round 1's three-buffer pool (stage_mode 0)
    for threads, chunks, mode, kb, groups, nt in ((1, 0, 1, 256, 4, 0), (3, 5, 1, 4, 2, 0), (8, 1, 1, 16, 3, 1), (0, 0, 1, 64, 8, 1),
                                                  (5, 22, 1, 8, 4, 0), (2, 3, 1, 4096, 2, 0), (1, 0, 0, 256, 4, 0), (3, 5, 0, 256, 4, 0),
                                                  (0, 0, 1, 512, 6, 1)):      # the defaults again
        engine.set_option("stage_threads", float(threads))
        engine.set_option("copy_chunks", float(chunks))
        engine.set_option("stage_mode", float(mode))
        engine.set_option("stage_slice_kb", float(kb))
        engine.set_option("stage_groups", float(groups))
        engine.set_option("stage_nt", float(nt))
        for _ in range(2):                      # the second call reuses the ring while nothing of the first is in flight
            got = engine.pairs(imgs, (32, 32), (16, 16))
            for a, b in zip(ref, got):
                assert np.array_equal(a, b, equal_nan=True)
    engine.set_option("copy_chunks", 0.0)
    f32 = imgs.astype(np.float32)
    a = engine.pairs(f32, (32, 32), (16, 16))
    p32 = engine.pinned_empty(f32.shape, np.float32)
    p32[...] = f32
    b = engine.pairs(p32, (32, 32), (16, 16))
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))


# ---- round 2: multi-engine / multi-device binding, stream-ordered ensemble, install() with the real engine ------------------
def _golden_frames():
    import os

    from pyorc_b200 import _xr

    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ngwerere_proj.npz"))
    fr, t, res = d["frames"], d["time_s"], float(d["resolution"])
    H, W = fr.shape[1:]
    y = np.flipud(np.linspace(res / 2, res * (H - 0.5), H))
    x = np.linspace(res / 2, res * (W - 0.5), W)
    return _xr.DataArray(fr, ("time", "y", "x"), {"time": t, "y": y, "x": x}), res, d["pinned_vx_timestep"], d["pinned_vx_ensemble"]


def test_get_b2piv_devices_equals_single_engine():
    """`get_b2piv(devices=[...])`: sub-ranges of every chunk on several engines (here three engines on cuda:0, one host thread
    each) give the Dataset of a single engine - bit-identical per time step, equal up to float32 summation order in ensemble mode."""
    from pyorc_b200 import _xr, velocimetry

    n = 14
    imgs = synth.particle_frames(n, 200, 304, dtype=np.uint8)
    da = _xr.DataArray(imgs, ("time", "y", "x"), {"time": np.arange(n) / 30.0})
    ws, ov = (64, 64), (32, 32)
    nr, nc = O.get_array_shape(imgs.shape[-2:], ws, ov)
    args = (da, np.arange(nr), np.arange(nc), np.full(n - 1, 1 / 30), ws, ov, ws, 0.01, 0.01)
    one = velocimetry.get_b2piv(*args, chunksize=8)
    many = velocimetry.get_b2piv(*args, chunksize=8, devices=[0, 0, 0])
    for k in ("v_x", "v_y", "corr", "s2n"):
        assert np.array_equal(one[k].values, many[k].values, equal_nan=True), k
    assert np.array_equal(one.coords["time"], many.coords["time"])
    kw = dict(ensemble_corr=True, corr_min=0.1, s2n_min=1.5, count_min=0.2, chunksize=8)
    one = velocimetry.get_b2piv(*args, **kw)
    many = velocimetry.get_b2piv(*args, **kw, devices=[0, 0, 0])
    assert np.isfinite(one["v_x"].values).mean() > 0.9
    for k in ("corr", "s2n"):
        assert np.array_equal(one[k].values, many[k].values, equal_nan=True), k
    for k in ("v_x", "v_y"):
        assert np.array_equal(np.isnan(one[k].values), np.isnan(many[k].values))
        assert np.nanmax(np.abs(one[k].values - many[k].values)) <= 2e-3 * 0.01 * 30      # 2e-3 px / frame in m / s


def test_ensemble_is_stream_ordered_without_host_synchronisation(engine):
    """begin / add / finish on different streams: accumulate device frames on a side stream behind a long
    dependency, then finish at once on the engine's stream (host call) and on torch's current stream (device call) - both must
    wait for the accumulation and equal the fully synchronous result."""
    import torch

    imgs = synth.particle_frames(9, 270, 400, dtype=np.uint8)
    ws, ov = (64, 64), (32, 32)
    engine.set_option("clip_normalized", 0.0)
    engine.set_option("kernel_variant", 0.0)
    engine.ens_begin(imgs.shape[-2:], ws, ov, np.uint8)
    engine.ens_add(imgs, ws, ov, corr_min=0.1, s2n_min=1.5)
    u0, v0, c0 = engine.ens_finish(0.2)
    d = torch.from_numpy(imgs).cuda()
    side = torch.cuda.Stream()
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for finish_on_device in (False, True):
        torch.cuda.synchronize()
        engine.ens_begin(imgs.shape[-2:], ws, ov, np.uint8, device_ordered=True)     # zero-fill on torch's current stream
        with torch.cuda.stream(side):
            for _ in range(20):
                big.add_(1)                                                             # keeps the side stream busy for a while
            engine.ens_add(d, ws, ov, corr_min=0.1, s2n_min=1.5)                        # launched behind it, not awaited
        if finish_on_device:
            ut, vt = engine.ens_finish_device(0.2)
            u, v = ut.cpu().numpy(), vt.cpu().numpy()
        else:
            u, v, c = engine.ens_finish(0.2)
            assert np.array_equal(c, c0)
        assert np.array_equal(u, u0, equal_nan=True) and np.array_equal(v, v0, equal_nan=True)
    torch.cuda.synchronize()


def test_ensemble_window_slices_finish_like_the_whole(engine):
    """`ens_finish_device(first, n)` on window slices (what a rank does after the reduce-scatter) = the whole-field finish."""
    from pyorc_b200 import parallel

    imgs = synth.particle_frames(5, 270, 400, dtype=np.uint8)
    ws, ov = (32, 32), (16, 16)
    nr, nc = engine.ens_begin(imgs.shape[-2:], ws, ov, np.uint8)
    engine.ens_add(imgs, ws, ov, corr_min=0.1, s2n_min=1.5)
    u0, v0, _ = engine.ens_finish(0.2)
    for world in (2, 3, 8):
        u = np.empty_like(u0)
        v = np.empty_like(v0)
        for a, b in parallel.window_slices(nr * nc, world):
            ut, vt = engine.ens_finish_device(0.2, int(a), int(b - a))
            u[a:b], v[a:b] = ut.cpu().numpy(), vt.cpu().numpy()
        assert np.array_equal(u, u0, equal_nan=True) and np.array_equal(v, v0, equal_nan=True)


def test_install_dispatch_real_engine(monkeypatch):
    """`pyorc_b200.frames.install()` on a pyorc-shaped module pair (the real pyorc cannot travel to this box - see
    tests/test_reference_dropin.py for its own body): `Frames.get_piv(engine="b200")` reaches the CUDA engine through the
    patched `pyorc.velocimetry.ffpiv.get_ffpiv` call site and reproduces pyorc's pinned vectors (tests/test_frames.py:139-153)."""
    import sys
    import types

    from pyorc_b200 import frames as b2frames, window

    da, res, pin_ts, pin_ens = _golden_frames()
    calls = []
    ffpiv_mod = types.ModuleType("pyorc.velocimetry.ffpiv")

    def get_ffpiv(frames, y, x, dt, *a, **kw):       # stands for the reference's CPU arm
        calls.append(kw.get("engine"))
        raise RuntimeError("reference arm reached")

    ffpiv_mod.get_ffpiv = get_ffpiv
    frames_mod = types.ModuleType("pyorc.api.frames")

    class Frames:                                      # the call site of pyorc/api/frames.py:156-188, nothing else
        def __init__(self, obj):
            self._obj = obj

        def get_piv(self, window_size=None, overlap=None, engine="numba", ensemble_corr=False, **kwargs):
            dt = self._obj["time"].diff(dim="time")
            ws = window.round_to_even(2 * (window_size,))
            if overlap is None:
                overlap = 2 * (int(round(window_size) / 2),)
            cols, rows = window.get_rect_coordinates(dim_size=self._obj[0].shape, window_size=ws, search_area_size=ws, overlap=overlap)
            if engine not in ["numba", "numpy"]:
                raise ValueError(f"Selected PIV engine {engine} does not exist.")
            kwargs = {**kwargs, "search_area_size": ws, "window_size": ws, "overlap": overlap, "res_x": res, "res_y": res}
            return ffpiv_mod.get_ffpiv(self._obj, self._obj.y.values[rows], self._obj.x.values[cols], dt, engine=engine,
                                       ensemble_corr=ensemble_corr, **kwargs)

    frames_mod.Frames = Frames
    for name, m in {"pyorc": types.ModuleType("pyorc"), "pyorc.api": types.ModuleType("pyorc.api"), "pyorc.api.frames": frames_mod,
                    "pyorc.velocimetry": types.ModuleType("pyorc.velocimetry"), "pyorc.velocimetry.ffpiv": ffpiv_mod}.items():
        monkeypatch.setitem(sys.modules, name, m)
    assert b2frames.install()
    try:
        for ens, pin in ((False, pin_ts), (True, pin_ens)):
            piv = Frames(da).get_piv(window_size=10, engine="b200", ensemble_corr=ens, s2n_min=0, corr_min=0, count_min=0, devices=[0, 0])
            got = piv.mean(dim="time", keep_attrs=True)["v_x"].values.flatten()[-4:]
            assert np.allclose(got, pin, rtol=0, atol=2e-6, equal_nan=True), (got, pin)
        assert calls == []
        with pytest.raises(RuntimeError, match="reference arm reached"):
            Frames(da).get_piv(window_size=10, engine="numba")
        assert calls == ["numba"]
    finally:
        b2frames.uninstall()


@pytest.mark.parametrize("tmem", [1, 0])
@pytest.mark.parametrize("ws,ov,shape,parts", [((64, 64), (32, 32), (7, 200, 304), 12), ((64, 64), (32, 32), (4, 270, 400), 6),
                                                   ((64, 64), (40, 40), (9, 160, 208), 18), ((64, 64), (32, 32), (12, 200, 304), 30)])
def test_rows64_even_work_partition_and_tensor_memory(engine, ws, ov, shape, parts, tmem):
    """64x64 kernel variants: parked spectra in Tensor Memory (6 groups per SM) or in shared memory, work dealt as an even 1-D
    partition of the (window pair, frame pair) space (forced here on small problems: parts end inside window pairs, parts
    with several segments, more parts than window pairs) - same planes, peaks and NaN masks as the oracle."""
    imgs = synth.particle_frames(*shape, dtype=np.uint8)
    imgs[1:, 60:150, 100:230] = 7            # dead windows
    engine.set_option("tmem", float(tmem))
    engine.set_option("unit_parts", float(parts))
    try:
        compare(engine, imgs, ws, ov, 0, variant=2, run_len=0)
        engine.set_option("unit_parts", 0.0)
        compare(engine, imgs, ws, ov, 0, variant=2, run_len=3)    # the unit / wave scheme with runs of three pairs
    finally:
        engine.set_option("unit_parts", 0.0)
        engine.set_option("tmem", 1.0)


def test_fused_unit_conversion_is_numpys_arithmetic(engine):
    """`pairs(..., units=(res_x, res_y, dt))` = `(u * res / dt[:, None, None]).astype(float32)` bit for bit (float32 product,
    float64 division, one rounding - what numpy does for a float32 `u`, pyorc/velocimetry/ffpiv.py:418-419), non-uniform dt."""
    imgs = synth.particle_frames(6, 200, 304, dtype=np.uint8)
    imgs[:, :64, :64] = 0
    ws, ov = (64, 64), (32, 32)
    engine.set_option("kernel_variant", 0.0)
    u, v, c, s = engine.pairs(imgs, ws, ov)
    dt = np.array([1 / 30.0, 1 / 25.0, 0.04, 1 / 29.97, 0.0333])
    res_x, res_y = 0.01, 0.0125
    vx, vy, c2, s2 = engine.pairs(imgs, ws, ov, units=(res_x, res_y, dt))
    want_x = (u * res_x / np.expand_dims(dt, (1, 2))).astype(np.float32)
    want_y = (v * res_y / np.expand_dims(dt, (1, 2))).astype(np.float32)
    assert np.array_equal(vx, want_x, equal_nan=True) and np.array_equal(vy, want_y, equal_nan=True)
    assert np.array_equal(c, c2) and np.array_equal(s, s2, equal_nan=True)
    with pytest.raises(ValueError):
        engine.pairs(imgs, ws, ov, units=(res_x, res_y, dt[:3]))


@pytest.mark.parametrize("ws,ov,shape,run_len", [((128, 128), (64, 64), (30, 300, 420), 4), ((128, 128), (64, 64), (26, 300, 420), 0),
                                                   ((64, 64), (32, 32), (40, 200, 304), 5), ((64, 64), (32, 32), (40, 200, 304), 0),
                                                   ((32, 32), (24, 24), (34, 100, 144), 3)])
def test_long_sequences_wrap_the_run_length_many_times(engine, ws, ov, shape, run_len):
    """Work units that start and stop many times along a long frame sequence (the full-size tests only wrap a run a
    few times): 128x128 with runs of 4 pairs over 29 pairs, 64x64 / 32x32 likewise, and the default unit / partition schemes."""
    imgs = synth.particle_frames(*shape, dtype=np.uint8)
    imgs[7:9, : ws[0], : ws[1]] = 9          # a window that is dead for two frames in the middle of a run
    compare(engine, imgs, ws, ov, 0, variant=2, run_len=run_len, check_planes=False)
    engine.set_option("kernel_variant", 0.0)
    engine.set_option("run_len", 0.0)


def test_available_memory_cache():
    """window.available_memory: the driver is asked when no need is given, when the cached value is old, or when it is less than twice
    the need; otherwise the recent answer is reused (cudaMemGetInfo costs 1 - 7 ms, a whole get_b2piv call 4.6 ms)."""
    import torch

    from pyorc_b200 import window

    window._FREE_CACHE.clear()
    a = window.available_memory(0)
    assert 0 < a <= torch.cuda.get_device_properties(0).total_memory
    hold = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # changes the free memory (unless the caching allocator had it)
    assert window.available_memory(0, need=1e6) == a                      # reused: a >= 2 * need
    assert window.available_memory(0, need=a) <= a                        # too close to the limit: asked again
    assert window.available_memory(0, need=1e6, max_age=0.0) > 0          # too old: asked again
    window._FREE_CACHE[0] = (window._FREE_CACHE[0][0], 123.0)
    assert window.available_memory(0, need=50.0) == 123.0
    assert window.available_memory(0) != 123.0                            # no need given: always the driver
    del hold

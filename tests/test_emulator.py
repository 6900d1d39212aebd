"""CPU: the CUDA kernel's barrier-delimited phases, compiled for the host (tests/emul), against the oracle.
Checks the FFT factorisation, digit-swapped spectrum order, cross-spectrum packing, fftshift indexing, first-argmax
and Gaussian fit without a GPU.  The emulator is test infrastructure only."""
import ctypes

import numpy as np
import pytest

from oracle import ffpiv_oracle as O
from pyorc_b200 import synth


@pytest.fixture(scope="module")
def emul():
    import __graft_entry__ as g

    lib = ctypes.CDLL(g.build_emulator())

    def run(imgs, ws, ov, nwin=2, clip=1, border_nan=1):
        imgs = np.ascontiguousarray(imgs)
        n, H, W = imgs.shape
        nr, nc = O.get_array_shape((H, W), ws, ov)
        outs = [np.full((n - 1, nr, nc), -7, np.float32) for _ in range(4)]
        planes = np.zeros((n - 1, nr * nc, ws[0], ws[1]), np.float32)
        rc = lib.b2piv_emul_pairs(
            imgs.ctypes.data_as(ctypes.c_void_p), n, H, W, int(imgs.dtype == np.float32), ws[0], ws[1], ov[0], ov[1], nwin, clip,
            border_nan, ctypes.c_float(1e-7), None, *[o.ctypes.data_as(ctypes.c_void_p) for o in outs], planes.ctypes.data_as(ctypes.c_void_p),
        )
        assert rc == 0
        return outs, planes

    return run


@pytest.mark.parametrize(
    "ws,ov,shape",
    [((64, 64), (32, 32), (2, 140, 200)), ((32, 32), (24, 24), (3, 70, 90)), ((16, 16), (8, 8), (2, 50, 60)),
     ((32, 64), (16, 32), (2, 70, 140)), ((64, 32), (32, 16), (2, 140, 70)), ((128, 128), (64, 64), (2, 130, 200))],
)
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
@pytest.mark.parametrize("clip", [1, 0])
def test_phases_match_oracle(emul, ws, ov, shape, dtype, clip):
    O.CLIP_NORMALIZED = bool(clip)
    imgs = synth.particle_frames(*shape, dtype=dtype)
    nr, nc = O.get_array_shape(shape[1:], ws, ov)
    _, _, corr = O.cross_corr(imgs, ws, ov)
    u, v, c, s = O.uv_timestep(imgs, nc, nr, ws, ov)
    for nwin in (2, 1):
        (eu, ev, ec, es), pl = emul(imgs, ws, ov, nwin=nwin, clip=clip)
        assert np.abs(pl - corr).max() < 2e-6
        assert np.array_equal(np.isnan(eu), np.isnan(u))
        ok = np.isfinite(u)
        assert np.abs(eu[ok] - u[ok]).max() < 1e-3 and np.abs(ev[ok] - v[ok]).max() < 1e-3
        assert np.abs(ec - c).max() < 2e-6
        assert np.nanmax(np.abs(es - s) / np.abs(s)) < 1e-5


@pytest.fixture(scope="module")
def emul_rows():
    import __graft_entry__ as g

    lib = ctypes.CDLL(g.build_emulator())

    def run(imgs, win, ovl, run_len=0, clip=1, border_nan=1):
        imgs = np.ascontiguousarray(imgs)
        n, H, W = imgs.shape
        nr, nc = O.get_array_shape((H, W), (win, win), (ovl, ovl))
        outs = [np.full((n - 1, nr, nc), -7, np.float32) for _ in range(4)]
        planes = np.zeros((n - 1, nr * nc, win, win), np.float32)
        fn = lib.b2piv_emul_rows_f32 if imgs.dtype == np.float32 else lib.b2piv_emul_rows
        rc = fn(
            imgs.ctypes.data_as(ctypes.c_void_p), n, H, W, win, ovl, run_len, clip, border_nan, ctypes.c_float(1e-7), None,
            *[o.ctypes.data_as(ctypes.c_void_p) for o in outs], planes.ctypes.data_as(ctypes.c_void_p),
        )
        assert rc == 0
        return outs, planes

    return run


@pytest.mark.parametrize(
    "win,ovl,shape,run_len",
    [(64, 32, (4, 200, 304), 0), (64, 32, (5, 136, 208), 2), (32, 16, (4, 100, 160), 0), (32, 24, (4, 80, 96), 3), (64, 48, (3, 140, 160), 1),
     (64, 40, (3, 140, 176), 0), (32, 20, (3, 80, 112), 2)],   # strides 24 / 12: window starts only 8- / 4-byte aligned
)
@pytest.mark.parametrize("clip", [1, 0])
def test_rows_phases_match_oracle(emul_rows, win, ovl, shape, run_len, clip):
    """Row-per-thread kernel phases (piv_rows.cuh): swizzled tile reads, register FFTs, partner-lane Hermitian split,
    forward-spectrum sharing across consecutive pairs, first-argmax, dead windows."""
    O.CLIP_NORMALIZED = bool(clip)
    imgs = synth.particle_frames(*shape, dtype=np.uint8)
    imgs[:, :40, :50] = 0
    ws, ov = (win, win), (ovl, ovl)
    nr, nc = O.get_array_shape(shape[1:], ws, ov)
    _, _, corr = O.cross_corr(imgs, ws, ov)
    u, v, c, s = O.uv_timestep(imgs, nc, nr, ws, ov)
    (eu, ev, ec, es), pl = emul_rows(imgs, win, ovl, run_len, clip)
    assert np.abs(pl - corr).max() < 2e-6
    assert np.array_equal(np.isnan(eu), np.isnan(u)) and np.array_equal(np.isnan(es), np.isnan(s))
    ok = np.isfinite(u)
    assert np.abs(eu[ok] - u[ok]).max() < 1e-3 and np.abs(ev[ok] - v[ok]).max() < 1e-3
    assert np.abs(ec - c).max() < 2e-6
    assert np.nanmax(np.abs(es - s) / np.abs(s)) < 1e-5


def test_rows_phases_config0_ngwerere(emul_rows):
    """BASELINE.json configs[0] on the CPU: the 32x32 row-per-thread kernel phases on the three orthorectified Ngwerere frames of
    pyorc's own test (tests/golden/ngwerere_proj.npz, made by the reference's code), every window against the oracle - real river
    imagery instead of synthetic particles (the GPU counterpart: tests/test_full_size.py::test_config0_ngwerere_32x32)."""
    import os

    O.CLIP_NORMALIZED = False
    frames = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ngwerere_proj.npz"))["frames"]
    ws, ov = (32, 32), (16, 16)
    nr, nc = O.get_array_shape(frames.shape[-2:], ws, ov)
    assert (nr, nc) == (28, 22)
    _, _, corr = O.cross_corr(frames, ws, ov)
    u, v, c, s = O.uv_timestep(frames, nc, nr, ws, ov)
    (eu, ev, ec, es), pl = emul_rows(frames, 32, 16, 0, 0)
    assert np.abs(pl - corr).max() < 2e-6
    assert np.array_equal(np.isnan(eu), np.isnan(u)) and np.array_equal(np.isnan(es), np.isnan(s))
    ok = np.isfinite(u)
    assert np.abs(eu[ok] - u[ok]).max() < 1e-4 and np.abs(ev[ok] - v[ok]).max() < 1e-4
    assert np.abs(ec - c).max() < 2e-6
    assert np.nanmax(np.abs(es - s) / np.abs(s)) < 1e-5


@pytest.fixture(scope="module")
def emul_direct():
    import __graft_entry__ as g

    lib = ctypes.CDLL(g.build_emulator())

    def run(imgs, ws, ov, clip=0, border_nan=1):
        imgs = np.ascontiguousarray(imgs)
        n, H, W = imgs.shape
        nr, nc = O.get_array_shape((H, W), ws, ov)
        outs = [np.full((n - 1, nr, nc), -7, np.float32) for _ in range(4)]
        planes = np.zeros((n - 1, nr * nc, ws[0], ws[1]), np.float32)
        rc = lib.b2piv_emul_direct(
            imgs.ctypes.data_as(ctypes.c_void_p), n, H, W, int(imgs.dtype == np.float32), ws[0], ws[1], ov[0], ov[1], clip, border_nan,
            ctypes.c_float(1e-7), *[o.ctypes.data_as(ctypes.c_void_p) for o in outs], planes.ctypes.data_as(ctypes.c_void_p),
        )
        assert rc == 0
        return outs, planes

    return run


@pytest.mark.parametrize("ws,ov,shape", [((10, 10), (5, 5), (3, 60, 70)), ((26, 26), (12, 12), (2, 80, 100)), ((20, 14), (10, 7), (2, 70, 60)),
                                         ((9, 11), (4, 5), (2, 40, 50)), ((32, 32), (16, 16), (2, 70, 90))])
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_direct_phases_match_oracle(emul_direct, ws, ov, shape, dtype):
    """Any-size direct-correlation kernel (pyorc's non power-of-two windows: 10, 20, 26, ...)."""
    O.CLIP_NORMALIZED = False
    imgs = synth.particle_frames(*shape, dtype=dtype)
    imgs[:, :12, :12] = 0
    nr, nc = O.get_array_shape(shape[1:], ws, ov)
    _, _, corr = O.cross_corr(imgs, ws, ov)
    u, v, c, s = O.uv_timestep(imgs, nc, nr, ws, ov)
    (eu, ev, ec, es), pl = emul_direct(imgs, ws, ov)
    assert np.abs(pl - corr).max() < 2e-6
    assert np.array_equal(np.isnan(eu), np.isnan(u)) and np.array_equal(np.isnan(es), np.isnan(s))
    ok = np.isfinite(u)
    assert np.abs(eu[ok] - u[ok]).max() < 1e-3 and np.abs(ev[ok] - v[ok]).max() < 1e-3
    assert np.abs(ec - c).max() < 2e-6


def test_direct_phases_reproduce_reference_golden(emul_direct):
    """The kernel code path itself (on the CPU) against pyorc's pinned v_x vector (tests/test_frames.py:143)."""
    import os

    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ngwerere_proj.npz"))
    frames, t, res, pin = d["frames"], d["time_s"], float(d["resolution"]), d["pinned_vx_timestep"]
    (eu, ev, ec, es), _ = emul_direct(frames, (10, 10), (5, 5))
    vx = (eu * res / np.diff(t)[:, None, None]).astype(np.float32)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        got = np.nanmean(vx, axis=0).flatten()[-4:]
    assert np.allclose(got, pin, rtol=0, atol=2e-6), (got, pin)


@pytest.mark.parametrize("ws,ov,shape", [((26, 26), (12, 12), (2, 80, 100)), ((20, 14), (10, 7), (2, 70, 60)), ((50, 50), (25, 25), (2, 110, 130)),
                                         ((10, 10), (5, 5), (2, 40, 50)), ((24, 40), (12, 20), (2, 80, 120))])
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_padded_fft_phases_match_oracle(emul, ws, ov, shape, dtype):
    """Windows that are not a power of two run EXACTLY through the power-of-two FFT kernel: zero-padded a, periodically
    tiled b, plane >= 2n (piv_core.cuh phase_embed)."""
    O.CLIP_NORMALIZED = False
    imgs = synth.particle_frames(*shape, dtype=dtype)
    imgs[:, :12, :12] = 0
    nr, nc = O.get_array_shape(shape[1:], ws, ov)
    _, _, corr = O.cross_corr(imgs, ws, ov)
    u, v, c, s = O.uv_timestep(imgs, nc, nr, ws, ov)
    for nwin in (2, 1):
        (eu, ev, ec, es), pl = emul(imgs, ws, ov, nwin=nwin, clip=0)
        assert np.abs(pl - corr).max() < 3e-6
        assert np.array_equal(np.isnan(eu), np.isnan(u)) and np.array_equal(np.isnan(es), np.isnan(s))
        ok = np.isfinite(u)
        assert np.abs(eu[ok] - u[ok]).max() < 1e-3 and np.abs(ev[ok] - v[ok]).max() < 1e-3
        assert np.abs(ec - c).max() < 3e-6


@pytest.mark.parametrize("win,ovl,shape,run_len", [(64, 32, (4, 200, 304), 0), (32, 16, (4, 100, 160), 2), (64, 44, (3, 140, 164), 0), (32, 20, (3, 80, 112), 0)])
def test_rows_phases_float32_frames(emul_rows, win, ovl, shape, run_len):
    """float32 frames through the row-per-thread kernel: 128-byte swizzled boxes, two-pass moments, one window per TMA
    phase for 64x64."""
    O.CLIP_NORMALIZED = False
    imgs = synth.particle_frames(*shape, dtype=np.float32)
    imgs[:, :40, :52] = 0
    imgs[1:] -= 0.25 * imgs[:-1]           # time_diff-like frames with negative values
    ws, ov = (win, win), (ovl, ovl)
    nr, nc = O.get_array_shape(shape[1:], ws, ov)
    _, _, corr = O.cross_corr(imgs, ws, ov)
    u, v, c, s = O.uv_timestep(imgs, nc, nr, ws, ov)
    (eu, ev, ec, es), pl = emul_rows(imgs, win, ovl, run_len, 0)
    assert np.abs(pl - corr).max() < 3e-6
    assert np.array_equal(np.isnan(eu), np.isnan(u)) and np.array_equal(np.isnan(es), np.isnan(s))
    ok = np.isfinite(u)
    assert np.abs(eu[ok] - u[ok]).max() < 1e-3 and np.abs(ev[ok] - v[ok]).max() < 1e-3
    assert np.abs(ec - c).max() < 3e-6


@pytest.fixture(scope="module")
def emul_rows_pad():
    import __graft_entry__ as g

    lib = ctypes.CDLL(g.build_emulator())

    def run(imgs, ws, ov, run_len=0, clip=0, border_nan=1):
        imgs = np.ascontiguousarray(imgs)
        n, H, W = imgs.shape
        nr, nc = O.get_array_shape((H, W), ws, ov)
        outs = [np.full((n - 1, nr, nc), -7, np.float32) for _ in range(4)]
        planes = np.zeros((n - 1, nr * nc, ws[0], ws[1]), np.float32)
        fn = lib.b2piv_emul_rows_pad_f32 if imgs.dtype == np.float32 else lib.b2piv_emul_rows_pad
        rc = fn(
            imgs.ctypes.data_as(ctypes.c_void_p), n, H, W, ws[0], ws[1], ov[0], ov[1], run_len, clip, border_nan, ctypes.c_float(1e-7), None,
            *[o.ctypes.data_as(ctypes.c_void_p) for o in outs], planes.ctypes.data_as(ctypes.c_void_p),
        )
        assert rc == 0
        return outs, planes

    return run


@pytest.mark.parametrize("ws,ov,shape,run_len", [((26, 26), (12, 12), (3, 96, 128), 0), ((20, 20), (10, 10), (4, 70, 96), 2),
                                                  ((10, 10), (5, 5), (3, 48, 64), 0), ((30, 18), (15, 9), (3, 100, 80), 0),
                                                  ((32, 32), (15, 15), (3, 100, 112), 1), ((16, 16), (8, 8), (3, 50, 64), 0),
                                                  ((26, 26), (12, 12), (2, 90, 96), 0)])
@pytest.mark.parametrize("clip", [0, 1])
def test_rows_padded_phases_match_oracle(emul_rows_pad, ws, ov, shape, run_len, clip):
    """Padded mode of the row-per-thread kernel (pyorc's non power-of-two windows, any stride): zero-padded embedding,
    byte-granular row offsets, the 2 x 2 tiling applied as a spectrum factor, block-restricted max / mean / first-argmax."""
    O.CLIP_NORMALIZED = bool(clip)
    imgs = synth.particle_frames(*shape, dtype=np.uint8)
    imgs[:, : ws[0], : ws[1] + 3] = 0          # a dead window (and a partly dark neighbour)
    nr, nc = O.get_array_shape(shape[1:], ws, ov)
    _, _, corr = O.cross_corr(imgs, ws, ov)
    u, v, c, s = O.uv_timestep(imgs, nc, nr, ws, ov)
    (eu, ev, ec, es), pl = emul_rows_pad(imgs, ws, ov, run_len, clip)
    assert np.abs(pl - corr).max() < 3e-6
    assert np.array_equal(np.isnan(eu), np.isnan(u)) and np.array_equal(np.isnan(es), np.isnan(s))
    ok = np.isfinite(u)
    same = np.abs(np.round(eu[ok]) - np.round(u[ok])) + np.abs(np.round(ev[ok]) - np.round(v[ok])) < 0.5
    assert same.mean() > 0.99
    assert np.abs(eu[ok][same] - u[ok][same]).max() < 2e-3 and np.abs(ev[ok][same] - v[ok][same]).max() < 2e-3
    assert np.abs(ec - c).max() < 3e-6
    assert np.nanmax(np.abs(es - s) / np.abs(s)) < 2e-5


@pytest.mark.parametrize("ws,ov,shape,run_len", [((26, 26), (12, 12), (3, 96, 128), 0), ((20, 20), (10, 10), (4, 70, 96), 2),
                                                  ((10, 10), (5, 5), (3, 48, 64), 0), ((30, 18), (15, 9), (3, 100, 81), 0),
                                                  ((32, 32), (15, 15), (3, 100, 113), 1), ((16, 16), (8, 8), (3, 50, 64), 0),
                                                  ((12, 28), (5, 13), (3, 61, 99), 0)])
@pytest.mark.parametrize("clip", [0, 1])
def test_rows_padded_phases_float32_frames(emul_rows_pad, ws, ov, shape, run_len, clip):
    """Padded mode of the row-per-thread kernel on float32 frames - what pyorc's own recipe produces (normalize -> edge_detect ->
    minmax -> get_piv(window_size=25): float32 frames, 26 x 26 windows): boxes from the 16-byte boundary below any window start,
    two-pass moments over the window's pixels, frames with negative values, a dead window, frame widths that end inside a box."""
    O.CLIP_NORMALIZED = bool(clip)
    imgs = synth.particle_frames(*shape, dtype=np.float32)
    imgs[1:] -= 0.25 * imgs[:-1]           # time_diff / edge_detect-like frames with negative values
    imgs[:, : ws[0], : ws[1] + 3] = 0
    nr, nc = O.get_array_shape(shape[1:], ws, ov)
    _, _, corr = O.cross_corr(imgs, ws, ov)
    u, v, c, s = O.uv_timestep(imgs, nc, nr, ws, ov)
    (eu, ev, ec, es), pl = emul_rows_pad(imgs, ws, ov, run_len, clip)
    assert np.abs(pl - corr).max() < 3e-6
    assert np.array_equal(np.isnan(eu), np.isnan(u)) and np.array_equal(np.isnan(es), np.isnan(s))
    ok = np.isfinite(u)
    same = np.abs(np.round(eu[ok]) - np.round(u[ok])) + np.abs(np.round(ev[ok]) - np.round(v[ok])) < 0.5
    assert same.mean() > 0.99
    assert np.abs(eu[ok][same] - u[ok][same]).max() < 2e-3 and np.abs(ev[ok][same] - v[ok][same]).max() < 2e-3
    assert np.abs(ec - c).max() < 3e-6
    assert np.nanmax(np.abs(es - s) / np.abs(s)) < 2e-5

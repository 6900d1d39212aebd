"""Pre-processing filters: the GPU kernels against the numpy/OpenCV restatement of pyorc's own filter
expressions.  normalize / minmax / time_diff are bit-exact (same IEEE float32 operation order); the Gaussian filters
agree with cv2.GaussianBlur to float32 rounding."""
import numpy as np
import pytest

from oracle import preprocess_oracle as P
from pyorc_b200 import synth


def test_oracle_gaussian_kernel_rule_matches_opencv():
    """The tap rule restated in b2piv.cu (gauss_taps): dyadic tables up to 9 taps, sigma = 0.3*((k-1)/2-1)+0.8 beyond."""
    import cv2

    tabs = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
            7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
            9: [4 / 256, 13 / 256, 30 / 256, 51 / 256, 60 / 256, 51 / 256, 30 / 256, 13 / 256, 4 / 256]}
    for k, t in tabs.items():
        assert np.array_equal(cv2.getGaussianKernel(k, 0, cv2.CV_32F).ravel(), np.float32(t))
    for k in (11, 13, 21, 31):
        sigma = 0.3 * ((k - 1) * 0.5 - 1) + 0.8
        x = np.arange(k) - (k - 1) / 2
        g = np.exp(-x**2 / (2 * sigma**2))
        assert np.abs(cv2.getGaussianKernel(k, 0, cv2.CV_32F).ravel() - g / g.sum()).max() < 1e-7


def _golden(name):
    import os

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))


def test_oracle_pinned_on_the_references_filter_goldens():
    """pyorc pins two filter results: test_smooth (tests/test_frames.py:86-93, last four values of smooth() on the raw
    grayscale Ngwerere frames) and test_edge_detect (:96-103, last four values of edge_detect() on the projected frames,
    atol 0 on x86).  The fixtures hold exactly those inputs (made by the reference's own decode / projection code,
    tests/golden/make_ngwerere_golden.py); the oracle reproduces both pins to the last digit."""
    tail = _golden("ngwerere_gray_tail.npz")
    assert np.array_equal(P.smooth(tail["gray_tail"]).ravel()[-4:], tail["pinned_smooth_last4"].astype(np.float32))
    proj = _golden("ngwerere_proj.npz")["frames"]
    assert np.array_equal(P.edge_detect(proj).ravel()[-4:], tail["pinned_edge_detect_proj_last4"].astype(np.float32))


@pytest.mark.gpu
def test_gpu_filters_reproduce_the_references_pins():
    from pyorc_b200 import preprocess as G

    tail = _golden("ngwerere_gray_tail.npz")
    got = G.smooth(tail["gray_tail"])
    assert got.dtype == np.float32 and np.abs(got.ravel()[-4:] - tail["pinned_smooth_last4"]).max() <= 1e-4
    proj = _golden("ngwerere_proj.npz")["frames"]
    got = G.edge_detect(proj)
    assert got.shape == proj.shape and np.abs(got.ravel()[-4:] - tail["pinned_edge_detect_proj_last4"]).max() <= 1e-4
    assert np.abs(got - P.edge_detect(proj)).max() <= 1e-4


def test_oracle_normalize_properties():
    fr = synth.particle_frames(30, 40, 50, dtype=np.uint8)
    out = P.normalize(fr, samples=15)
    assert out.dtype == np.uint8 and out.shape == fr.shape
    assert (out.reshape(30, -1).max(axis=1) == 255).all() and (out.reshape(30, -1).min(axis=1) == 0).all()
    with pytest.raises(AssertionError):
        P.normalize(fr[:5], samples=15)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(120, 200), (121, 203)])     # vector path / scalar path (frame size not a multiple of 4)
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_gpu_normalize_minmax_timediff_bit_exact(dtype, shape):
    from pyorc_b200 import preprocess as G

    fr = synth.particle_frames(31, shape[0], shape[1], dtype=dtype)
    fr[7] = fr[7, 0, 0]                                   # a flat frame: 0/0 -> uint8 0
    if dtype == np.uint8:
        assert np.array_equal(G.normalize(fr, samples=15), P.normalize(fr, samples=15))
        assert np.array_equal(G.normalize(fr, samples=4), P.normalize(fr, samples=4))
    else:
        a, b = G.normalize(fr, samples=15), P.normalize(fr, samples=15)     # float mean: summation order differs by an ulp
        assert (np.abs(a.astype(int) - b.astype(int)) <= 1).all() and (a != b).mean() < 1e-3
    assert np.array_equal(G.minmax(fr, min=20, max=180), P.minmax(fr, min=20, max=180))
    assert np.array_equal(G.minmax(fr, max=100), P.minmax(fr, max=100)) and G.minmax(fr, max=100).dtype == fr.dtype   # same values; dtype kept
    # fractional bounds: numpy promotes uint8 frames to floating point and keeps the fraction (pyorc/api/frames.py:343-361)
    want = np.maximum(np.minimum(fr, 180.25), 20.5)
    got = G.minmax(fr, min=20.5, max=180.25)
    assert got.dtype == np.float32 and np.array_equal(got, want.astype(np.float32)) and got.min() == 20.5
    for thres, ab in ((0.0, False), (5.0, False), (2.0, True)):
        assert np.array_equal(G.time_diff(fr, thres=thres, abs=ab), P.time_diff(fr, thres=thres, abs=ab))
    with pytest.raises(AssertionError):
        G.normalize(fr[:5], samples=15)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_gpu_gaussian_filters_match_opencv(dtype):
    from pyorc_b200 import preprocess as G

    fr = synth.particle_frames(3, 97, 131, dtype=dtype)   # odd sizes: partial tiles + reflect-101 borders
    for wdw in (1, 2, 3, 5):
        ref = P.smooth(fr, wdw)
        got = G.smooth(fr, wdw)
        assert got.dtype == np.float32 and np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
    for w1, w2 in ((1, 2), (2, 4), (1, 7)):
        ref = P.edge_detect(fr, w1, w2)
        got = G.edge_detect(fr, w1, w2)
        assert np.abs(got - ref).max() <= 2e-4 * 255


@pytest.mark.gpu
def test_gpu_pipeline_stays_on_device_and_feeds_piv():
    """normalize -> get_piv without leaving the device, against oracle filters + oracle PIV."""
    import torch

    from oracle import ffpiv_oracle as O
    from pyorc_b200 import preprocess as G
    from pyorc_b200.engine import get_engine

    O.CLIP_NORMALIZED = False
    fr = synth.particle_frames(16, 200, 304, dtype=np.uint8)
    d = torch.from_numpy(fr).cuda()
    dn = G.normalize(d, samples=15)
    assert dn.is_cuda and dn.dtype == torch.uint8
    eng = get_engine(0)
    eng.set_option("clip_normalized", 0.0)
    eng.set_option("kernel_variant", 0.0)
    u, v, c, s = eng.pairs(dn, (64, 64), (32, 32))
    ref_frames = P.normalize(fr, samples=15)
    nr, nc = O.get_array_shape(fr.shape[1:], (64, 64), (32, 32))
    ou, ov, oc, os_ = O.uv_timestep(ref_frames, nc, nr, (64, 64), (32, 32))
    ok = np.isfinite(ou)
    assert np.abs(u.cpu().numpy()[ok] - ou[ok]).max() <= 2e-3 and np.abs(c.cpu().numpy() - oc).max() <= 5e-6


@pytest.mark.gpu
def test_gpu_pyorc_recipe_float32_frames_and_26px_windows():
    """pyorc's own example recipe (examples/ngwerere/ngwerere.yml): frames.normalize -> edge_detect(wdw_1=1, wdw_2=2) -> minmax(-5, 5)
    -> get_piv(window_size=25) on the projected Ngwerere frames.  The filters hand FLOAT32 frames to the PIV and the camera-config
    window becomes 26 x 26 with overlap 12 (frames.py:159-171): on the device that is the padded float32 mode of the row-per-thread
    kernel.  Filters and PIV stay in HBM; the result equals the float64 oracle on the same filtered frames, and get_piv on the host
    copy of those frames returns the same fields in m / s."""
    import os

    import torch

    from oracle import ffpiv_oracle as O
    from pyorc_b200 import _xr
    from pyorc_b200 import frames as b2frames
    from pyorc_b200 import preprocess as G
    from pyorc_b200.engine import get_engine

    O.CLIP_NORMALIZED = False
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ngwerere_proj.npz"))
    fr, t, res = np.ascontiguousarray(g["frames"][:, :, :368]), g["time_s"], float(g["resolution"])
    # (368 of the 371 columns: a caller-owned float32 device tensor needs a 16-byte pitch for the TMA kernels; host frames of any
    # width are re-pitched by the engine - test_rows_kernel_padded_mode_float32_frames covers odd widths)
    d = torch.from_numpy(fr).cuda()
    de = G.minmax(G.edge_detect(G.normalize(d, samples=3), 1, 2), -5, 5)   # 3 frames in the fixture: every frame is a sample
    assert de.is_cuda and de.dtype == torch.float32
    # the oracle filters agree with the device filters (their own tests pin them); the PIV is compared on identical inputs
    ref_f = P.minmax(P.edge_detect(P.normalize(fr, samples=3), 1, 2), -5, 5)
    he = de.cpu().numpy()
    assert np.abs(he - ref_f).max() <= 2e-4 * 255
    ws, ov = (26, 26), (12, 12)
    eng = get_engine(0)
    eng.set_option("clip_normalized", 0.0)
    eng.set_option("kernel_variant", 0.0)
    u, v, c, s = (x.cpu().numpy() for x in eng.pairs(de, ws, ov))
    assert eng.last_variant == 4                     # the padded row-per-thread kernel, not the shared-memory fallback
    nr, nc = O.get_array_shape(fr.shape[1:], ws, ov)
    ou, ovv, oc, os_ = O.uv_timestep(he, nc, nr, ws, ov)
    assert np.array_equal(np.isnan(u), np.isnan(ou))
    ok = np.isfinite(ou)
    same = np.abs(np.round(u[ok]) - np.round(ou[ok])) + np.abs(np.round(v[ok]) - np.round(ovv[ok])) < 0.5
    assert same.mean() >= 0.995
    assert np.abs(u[ok][same] - ou[ok][same]).max() <= 2e-3 and np.abs(v[ok][same] - ovv[ok][same]).max() <= 2e-3
    assert np.abs(c - oc).max() <= 5e-6
    # the reference-shaped call: window_size=25 -> 26 x 26, overlap 12, velocities in m / s on (time, y, x)
    da = _xr.DataArray(he, ("time", "y", "x"), {"time": t, "y": np.arange(he.shape[1]), "x": np.arange(he.shape[2])})
    ds = b2frames.get_piv(da, window_size=25, resolution=res)
    assert ds["v_x"].values.shape == (2, nr, nc)
    dtp = np.diff(t)[:, None, None]
    assert np.allclose(ds["v_x"].values, (u * np.float32(res) / dtp).astype(np.float32), rtol=0, atol=1e-6, equal_nan=True)
    assert np.allclose(ds["corr"].values, c, rtol=0, atol=1e-6)

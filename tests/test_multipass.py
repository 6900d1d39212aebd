"""Two-pass PIV of BASELINE.json configs[2] ("2-pass deform").  There is no reference implementation (ffpiv is single pass),
so the scheme is defined in oracle/multipass_oracle.py (parity UNPINNED, stated there): the CPU tests check the definition on
hand-made fields, the GPU tests check the CUDA path against it stage by stage and against the imposed synthetic field."""
import numpy as np
import pytest

from oracle import ffpiv_oracle as O
from oracle import multipass_oracle as MP
from pyorc_b200 import synth

COARSE, FINE = ((64, 64), (48, 48)), ((32, 32), (24, 24))


# ---- CPU: the definition -------------------------------------------------------------------------------------------------
def test_validate_replaces_outliers_and_nans_by_the_neighbourhood_median():
    u = np.full((1, 5, 6), 3.0)
    v = np.full((1, 5, 6), -1.0)
    u += np.arange(6)[None, None, :] * 0.01
    u[0, 2, 3] = 40.0          # spurious vector
    v[0, 1, 1] = np.nan        # invalid vector
    u[0, 4, 5] = 3.3           # |3.3 - 3.04| / (0.01 + 0.1) > 2  -> replaced as well
    ou, ov = MP.validate(u, v)
    assert abs(ou[0, 2, 3] - 3.03) < 0.011 and ov[0, 2, 3] == -1.0
    assert np.isfinite(ov[0, 1, 1]) and ov[0, 1, 1] == -1.0 and abs(ou[0, 1, 1] - 3.01) < 0.011
    assert abs(ou[0, 4, 5] - 3.04) < 0.011
    keep = np.ones_like(u, bool)
    keep[0, 2, 3] = keep[0, 1, 1] = keep[0, 4, 5] = False
    assert np.array_equal(ou[keep], u[keep]) and np.array_equal(ov[keep], v[keep])
    # a lone vector without any valid neighbour is kept; a lone NaN becomes 0
    lone = np.full((1, 1, 1), 7.0)
    assert MP.validate(lone, lone)[0][0, 0, 0] == 7.0
    assert MP.validate(np.full((1, 1, 1), np.nan), lone)[0][0, 0, 0] == 0.0


def test_predictor_interpolates_rounds_and_clamps():
    H, W = 256, 320
    nr1, nc1 = O.get_array_shape((H, W), *COARSE)
    u = np.full((1, nr1, nc1), 2.5)              # rint(2.5) = 2 (half to even)
    v = np.full((1, nr1, nc1), -3.5)             # rint(-3.5) = -4
    dy, dx = MP.predictor(u, v, (H, W), COARSE, FINE)
    nr2, nc2 = O.get_array_shape((H, W), *FINE)
    assert dy.shape == (1, nr2, nc2)
    y2, x2 = O.window_origins((H, W), *FINE)
    assert (dx[0, :, :-1] == 2).all() and (dy[0, 1:, :] == -4).all()
    assert (dy[0, 0, :] == 0).all()                                   # row 0 cannot move up: clamped to the frame
    assert ((x2[None, :] + dx[0] + 32) <= W).all() and ((y2[:, None] + dy[0]) >= 0).all()
    # a linear field is reproduced by the bilinear interpolation between the coarse centres
    yc = (np.arange(nr1) * 16 + 32)[:, None] * np.ones((1, nc1))
    dy2, _ = MP.predictor(np.zeros((1, nr1, nc1)), 0.05 * yc[None], (H, W), COARSE, FINE)
    inner = (y2 + 16 >= 32) & (y2 + 16 <= yc[-1, 0])
    assert np.array_equal(dy2[0, inner, 3], np.rint(0.05 * (y2[inner] + 16)).astype(int))


def test_two_pass_recovers_large_displacements_better_than_a_single_fine_pass():
    """10 px shifts are beyond what a 32 px window resolves well (a third of the particles leave the window); the coarse
    predictor brings the second pass back to a near-zero residual."""
    rng = np.random.default_rng(4)
    base = synth.particle_frames(1, 200 + 40, 260 + 40, dtype=np.uint8, seed=11)[0]
    a = base[20:220, 20:280]
    b = base[20 - 7:220 - 7, 20 - 10:280 - 10]       # content moved by +7 rows, +10 columns
    imgs = np.stack([a, b])
    u, v, cm, sn, dy, dx = MP.two_pass(imgs, COARSE, FINE)
    inner = (slice(None), slice(1, -1), slice(1, -1))
    assert np.median(dx[inner]) == 10 and np.median(dy[inner]) == 7
    ok = np.isfinite(u[inner])
    assert ok.mean() > 0.95
    assert np.abs(u[inner][ok] - 10).max() < 0.2 and np.abs(v[inner][ok] - 7).max() < 0.2
    nr, nc = O.get_array_shape(imgs.shape[-2:], *FINE)
    u1, v1, c1, _ = O.uv_timestep(imgs, nc, nr, *FINE)
    assert np.nanmean(cm[inner]) > np.nanmean(c1[inner]) + 0.15      # the peak is much stronger once the window follows the flow


# ---- GPU ---------------------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from pyorc_b200.engine import Engine

    e = Engine(0)
    e.set_option("clip_normalized", 0.0)
    O.CLIP_NORMALIZED = False
    yield e
    e.close()


@gpu
def test_gpu_predictor_equals_the_definition_on_the_same_pass1_fields(engine):
    import torch

    imgs = synth.particle_frames(4, 300, 420, dtype=np.uint8)
    imgs[:, :70, :90] = 0                                    # dead coarse windows -> NaN vectors to validate away
    d = torch.from_numpy(imgs).cuda()
    u1, v1, _, _ = engine.pairs(d, *COARSE)
    hu, hv = u1.cpu().numpy().copy(), v1.cpu().numpy().copy()
    hu[1, 3, 4] += 25.0                                      # a spurious vector
    u1 = torch.from_numpy(hu).cuda()
    shift = engine.predictor(u1, v1, imgs.shape[-2:], COARSE, FINE).cpu().numpy()
    vu, vv = MP.validate(hu, hv)
    dy, dx = MP.predictor(vu, vv, imgs.shape[-2:], COARSE, FINE)
    assert np.array_equal(shift[..., 0], dy) and np.array_equal(shift[..., 1], dx)
    assert np.isnan(hu).any() and np.abs(dx).max() < 12


@gpu
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_gpu_shifted_pass_matches_the_definition(engine, dtype):
    import torch

    imgs = synth.particle_frames(3, 200, 280, dtype=dtype)
    nr, nc = O.get_array_shape(imgs.shape[-2:], *FINE)
    rng = np.random.default_rng(2)
    y0, x0 = O.window_origins(imgs.shape[-2:], *FINE)
    dy = np.clip(rng.integers(-6, 7, (2, nr, nc)), -y0[None, :, None], (200 - 32 - y0)[None, :, None])
    dx = np.clip(rng.integers(-6, 7, (2, nr, nc)), -x0[None, None, :], (280 - 32 - x0)[None, None, :])
    shift = torch.from_numpy(np.stack([dy, dx], axis=-1).astype(np.int16)).cuda()
    gu, gv, gc, gs = (t.cpu().numpy() for t in engine.pairs_shifted(torch.from_numpy(imgs).cuda(), *FINE, shift))
    u, v, c, s = MP.shifted_pass(imgs, dy, dx, *FINE)
    assert np.array_equal(np.isnan(gu), np.isnan(u))
    fin = np.isfinite(u)
    same = fin & (np.abs(np.round(gu) - np.round(u)) + np.abs(np.round(gv) - np.round(v)) < 0.5)
    assert same[fin].mean() >= 0.995
    assert np.abs(gu[same] - u[same]).max() <= 2e-3 and np.abs(gv[same] - v[same]).max() <= 2e-3
    assert np.abs(gc - c).max() <= 5e-6
    ok = np.isfinite(s) & (s != 0)
    assert (np.abs(gs[ok] - s[ok]) / s[ok]).max() <= 2e-5
    # zero shifts reproduce the ordinary single pass: bit for bit on the shared-memory kernel (same code path) ...
    engine.set_option("kernel_variant", 1.0)
    ref = engine.pairs(torch.from_numpy(imgs).cuda(), *FINE)
    z = engine.pairs_shifted(torch.from_numpy(imgs).cuda(), *FINE, torch.zeros_like(shift))
    for a, b in zip(ref, z):
        assert np.array_equal(a.cpu().numpy(), b.cpu().numpy(), equal_nan=True)
    # ... and the displaced row-per-thread kernel (uint8 frames, the default) agrees with the shared-memory kernel
    gen = [t.cpu().numpy() for t in engine.pairs_shifted(torch.from_numpy(imgs).cuda(), *FINE, shift)]
    engine.set_option("kernel_variant", 0.0)
    for a, b in zip((gu, gv, gc, gs), gen):
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.nanmax(np.abs(a - b) / (1 + np.abs(a))) < 2e-5


@gpu
@pytest.mark.parametrize("ov,shape,run_len", [((24, 24), (5, 200, 288), 0), ((24, 24), (6, 130, 176), 2), ((16, 16), (4, 150, 208), 0),
                                              ((20, 20), (3, 120, 160), 1)])
def test_gpu_displaced_rows_kernel(engine, ov, shape, run_len):
    """The displaced second pass on the row-per-thread kernel (forced): arbitrary byte offsets, unit boundaries inside the
    stack, shifts up to the frame border, dead windows."""
    import torch

    imgs = synth.particle_frames(*shape, dtype=np.uint8)
    imgs[:, :50, :60] = 0
    n, H, W = shape
    ws = (32, 32)
    nr, nc = O.get_array_shape((H, W), ws, ov)
    rng = np.random.default_rng(7)
    y0, x0 = O.window_origins((H, W), ws, ov)
    dy = np.clip(rng.integers(-40, 41, (n - 1, nr, nc)), -y0[None, :, None], (H - 32 - y0)[None, :, None])
    dx = np.clip(rng.integers(-40, 41, (n - 1, nr, nc)), -x0[None, None, :], (W - 32 - x0)[None, None, :])
    shift = torch.from_numpy(np.stack([dy, dx], axis=-1).astype(np.int16)).cuda()
    engine.set_option("kernel_variant", 2.0)
    engine.set_option("run_len", float(run_len))
    gu, gv, gc, gs = (t.cpu().numpy() for t in engine.pairs_shifted(torch.from_numpy(imgs).cuda(), ws, ov, shift))
    engine.set_option("kernel_variant", 0.0)
    engine.set_option("run_len", 0.0)
    u, v, c, s = MP.shifted_pass(imgs, dy, dx, ws, ov)
    assert np.array_equal(np.isnan(gu), np.isnan(u))
    fin = np.isfinite(u)
    same = fin & (np.abs(np.round(gu) - np.round(u)) + np.abs(np.round(gv) - np.round(v)) < 0.5)
    assert same[fin].mean() >= 0.99          # random 40 px shifts decorrelate most windows: noisy planes, a few near-ties
    assert np.abs(gu[same] - u[same]).max() <= 2e-3 and np.abs(gv[same] - v[same]).max() <= 2e-3
    assert np.abs(gc - c).max() <= 5e-6
    ok = np.isfinite(s) & (s != 0)
    assert np.array_equal(np.isnan(gs), np.isnan(s))
    assert (np.abs(gs[ok] - s[ok]) / s[ok]).max() <= 2e-5


@gpu
def test_gpu_two_pass_end_to_end_against_definition_and_truth(engine):
    """configs[2] geometry at reduced frame size, host arrays in / out."""
    H, W = 360, 480
    imgs = synth.particle_frames(3, H, W, dtype=np.uint8)
    gu, gv, gc, gs = engine.pairs_two_pass(imgs, COARSE, FINE)
    u, v, c, s, dy, dx = MP.two_pass(imgs, COARSE, FINE)
    assert gu.shape == u.shape == (2,) + O.get_array_shape((H, W), *FINE)
    fin = np.isfinite(u) & np.isfinite(gu)
    assert np.array_equal(np.isnan(gu), np.isnan(u))
    close = np.abs(gu[fin] - u[fin]) + np.abs(gv[fin] - v[fin]) < 4e-3
    assert close.mean() >= 0.995          # a pass-1 vector at a rounding boundary may move a predictor by one pixel
    assert np.abs(gc[fin][close] - c[fin][close]).max() <= 5e-6
    # truth: the imposed field at the window centres
    y0, x0 = O.window_origins((H, W), *FINE)
    yc, xc = np.meshgrid(y0 + 16.0, x0 + 16.0, indexing="ij")
    tx, ty = synth.displacement_field(H, W, yc, xc)
    nr, nc = gu.shape[1:]
    one = engine.pairs(imgs, *FINE)
    e2 = np.nanmedian(np.abs(gu - tx[None])) + np.nanmedian(np.abs(gv - ty[None]))
    e1 = np.nanmedian(np.abs(one[0] - tx[None])) + np.nanmedian(np.abs(one[1] - ty[None]))
    assert e2 < 0.2 and e2 <= e1 + 0.01
    assert np.nanmean(gc) > np.nanmean(one[2])           # following the flow strengthens the correlation peak


@gpu
def test_gpu_two_pass_config2_full_frame(engine):
    """BASELINE configs[2] at full 1080p frame size (132 x 237 windows per pair), a shard of 4 pairs, device resident."""
    import torch

    H, W = 1080, 1920
    d = synth.particle_frames_torch(5, H, W, torch.device("cuda", 0), dtype="uint8")
    u, v, c, s = engine.pairs_two_pass(d, COARSE, FINE)
    assert tuple(u.shape) == (4, 132, 237) and u.is_cuda
    y0, x0 = O.window_origins((H, W), *FINE)
    yc, xc = np.meshgrid(y0 + 16.0, x0 + 16.0, indexing="ij")
    tx, ty = synth.displacement_field(H, W, yc, xc)
    hu, hv = u.cpu().numpy(), v.cpu().numpy()
    assert np.isfinite(hu).mean() > 0.97
    assert np.nanmedian(np.abs(hu - tx[None])) < 0.1 and np.nanmedian(np.abs(hv - ty[None])) < 0.1
    # spot check against the definition on one pair of a frame crop is covered above; here: determinism
    u2, v2, _, _ = engine.pairs_two_pass(d, COARSE, FINE)
    assert torch.equal(torch.nan_to_num(u), torch.nan_to_num(u2)) and torch.equal(torch.nan_to_num(v), torch.nan_to_num(v2))


@gpu
def test_gpu_get_piv_with_coarse_pass(engine):
    """The two-pass scheme behind the get_piv-shaped binding: get_piv(window_size=32, overlap=(24, 24), coarse_pass=...) returns the
    usual Dataset (v_x, v_y, corr, s2n on time / y / x, m/s) and equals Engine.pairs_two_pass times res / dt."""
    from pyorc_b200 import _xr
    from pyorc_b200 import frames as b2frames
    from pyorc_b200.engine import get_engine

    get_engine(0).set_option("clip_normalized", 0.0)
    H, W, res = 200, 288, 0.01
    imgs = synth.particle_frames(5, H, W, dtype=np.uint8)
    da = _xr.DataArray(imgs, ("time", "y", "x"), {"time": np.arange(5) / 30.0, "y": np.flipud(np.linspace(res / 2, res * (H - 0.5), H)),
                                                  "x": np.linspace(res / 2, res * (W - 0.5), W)})
    ds = b2frames.get_piv(da, window_size=32, overlap=(24, 24), engine="b200", resolution=res, coarse_pass=COARSE)
    u, v, c, s = engine.pairs_two_pass(imgs, COARSE, FINE)
    assert ds["v_x"].values.shape == u.shape
    assert np.allclose(ds["v_x"].values, (u * res * 30.0).astype(np.float32), equal_nan=True, rtol=1e-6, atol=0)
    assert np.allclose(ds["v_y"].values, (v * res * 30.0).astype(np.float32), equal_nan=True, rtol=1e-6, atol=0)
    assert np.array_equal(ds["corr"].values, c, equal_nan=True)


# ---- window deformation (the "deform" of configs[2]) -------------------------------------------------------------------------
COARSE50 = ((64, 64), (32, 32))


def test_deform_definition_uniform_field_is_a_plain_shift():
    rng = np.random.default_rng(5)
    imgs = rng.integers(0, 255, (2, 96, 128)).astype(np.uint8)
    nr, nc = O.get_array_shape((96, 128), *COARSE50)
    u = np.full((1, nr, nc), 3.0)
    v = np.full((1, nr, nc), -2.0)
    stack = MP.deform(imgs, u, v, COARSE50[0:2])
    assert stack.shape == (2, 96, 128) and stack.dtype == np.float32
    assert np.array_equal(stack[0], imgs[0].astype(np.float32))
    # B'(y, x) = frame1(y - 2, x + 3) wherever the sample stays inside the frame
    assert np.array_equal(stack[1][2:, :-3], imgs[1][:-2, 3:].astype(np.float32))
    # half-pixel predictor: the mean of two neighbours
    stack = MP.deform(imgs, u + 0.5, v, COARSE50)
    assert np.allclose(stack[1][2:, :-4], 0.5 * (imgs[1][:-2, 3:-1].astype(np.float64) + imgs[1][:-2, 4:]), atol=1e-4)
    dv, du = MP.predictor_float(u + 0.5, v, (96, 128), COARSE50, FINE)
    assert np.all(du == 3.5) and np.all(dv == -2.0) and du.shape == (1,) + O.get_array_shape((96, 128), *FINE)


def test_two_pass_deform_definition_recovers_the_imposed_field():
    H, W = 192, 256
    imgs = synth.particle_frames(2, H, W, dtype=np.uint8)
    u, v, c, s, stack, (dv, du) = MP.two_pass_deform(imgs, COARSE50, FINE)
    y0, x0 = O.window_origins((H, W), *FINE)
    yc, xc = np.meshgrid(y0 + 16.0, x0 + 16.0, indexing="ij")
    tx, ty = synth.displacement_field(H, W, yc, xc)
    assert np.nanmedian(np.abs(u[0] - tx)) < 0.1 and np.nanmedian(np.abs(v[0] - ty)) < 0.1
    one = O.uv_timestep(imgs, *O.get_array_shape((H, W), *FINE)[::-1], *FINE)
    assert np.nanmean(c) > np.nanmean(one[2])       # the warped frame correlates better than the raw one


@gpu
@pytest.mark.parametrize("dtype", [np.uint8, np.float32])
def test_gpu_deformed_stack_and_predictor_equal_the_definition(engine, dtype):
    import torch

    H, W = 200, 288
    imgs = synth.particle_frames(4, H, W, dtype=dtype)
    nr1, nc1 = O.get_array_shape((H, W), *COARSE50)
    u1, v1, _, _ = O.uv_timestep(imgs, nc1, nr1, *COARSE50)
    u1, v1 = u1.astype(np.float32), v1.astype(np.float32)      # the pass-1 fields the engine hands over are float32
    u1[0, 1, 2] = np.nan                    # a hole and an outlier for the validation step
    v1[1, 0, 0] = 25.0
    vu, vv = MP.validate(u1, v1)
    want = MP.deform(imgs, vu, vv, COARSE50)
    wdv, wdu = MP.predictor_float(vu, vv, (H, W), COARSE50, FINE)
    d = torch.from_numpy(imgs).cuda()
    stack, pred = engine.deform(d, torch.from_numpy(u1).cuda(), torch.from_numpy(v1).cuda(), COARSE50, FINE)
    got = stack.cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got[0::2], want[0::2])
    assert np.abs(got[1::2] - want[1::2]).max() <= 1e-4          # float64 arithmetic in the same order: differences are last-bit roundings
    assert (got[1::2] != want[1::2]).mean() < 1e-3
    p = pred.cpu().numpy()
    assert np.abs(p[..., 0] - wdv).max() <= 1e-6 and np.abs(p[..., 1] - wdu).max() <= 1e-6


@gpu
def test_gpu_two_pass_deform_against_definition_and_truth(engine):
    H, W = 360, 480
    imgs = synth.particle_frames(3, H, W, dtype=np.uint8)
    engine.set_option("clip_normalized", 0.0)
    O.CLIP_NORMALIZED = False
    gu, gv, gc, gs = engine.pairs_two_pass(imgs, COARSE50, FINE, mode="deform")
    u, v, c, s, _, _ = MP.two_pass_deform(imgs, COARSE50, FINE)
    assert gu.shape == u.shape == (2,) + O.get_array_shape((H, W), *FINE)
    assert np.array_equal(np.isnan(gu), np.isnan(u))
    fin = np.isfinite(u) & np.isfinite(gu)
    close = np.abs(gu[fin] - u[fin]) + np.abs(gv[fin] - v[fin]) < 4e-3
    assert close.mean() >= 0.995
    assert np.abs(gc[fin][close] - c[fin][close]).max() <= 1e-5
    y0, x0 = O.window_origins((H, W), *FINE)
    yc, xc = np.meshgrid(y0 + 16.0, x0 + 16.0, indexing="ij")
    tx, ty = synth.displacement_field(H, W, yc, xc)
    off = engine.pairs_two_pass(imgs, COARSE50, FINE, mode="offset")
    one = engine.pairs(imgs, *FINE)
    err = lambda r: np.nanmedian(np.abs(r[0] - tx[None])) + np.nanmedian(np.abs(r[1] - ty[None]))   # noqa: E731
    assert err((gu, gv)) < 0.15 and err((gu, gv)) <= err(off) + 0.01 and err((gu, gv)) <= err(one) + 0.01
    assert np.nanmean(gc) >= np.nanmean(off[2]) - 1e-3 and np.nanmean(gc) > np.nanmean(one[2])
    # chunking of the float32 stack does not change anything
    g2 = engine.pairs_two_pass(imgs, COARSE50, FINE, mode="deform", chunk_pairs=1)
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip((gu, gv, gc, gs), g2))


@gpu
def test_gpu_get_piv_with_deformation(engine):
    """`get_piv(..., coarse_pass=..., multipass="deform")`: the deformation scheme behind the get_piv-shaped binding."""
    from pyorc_b200 import _xr
    from pyorc_b200 import frames as b2frames
    from pyorc_b200.engine import get_engine

    get_engine(0).set_option("clip_normalized", 0.0)
    engine.set_option("clip_normalized", 0.0)
    H, W, res = 200, 288, 0.01
    imgs = synth.particle_frames(5, H, W, dtype=np.uint8)
    da = _xr.DataArray(imgs, ("time", "y", "x"), {"time": np.arange(5) / 30.0, "y": np.flipud(np.linspace(res / 2, res * (H - 0.5), H)),
                                                  "x": np.linspace(res / 2, res * (W - 0.5), W)})
    ds = b2frames.get_piv(da, window_size=32, overlap=(24, 24), engine="b200", resolution=res, coarse_pass=COARSE50, multipass="deform", chunksize=3)
    u, v, c, s = engine.pairs_two_pass(imgs, COARSE50, FINE, mode="deform")
    assert ds["v_x"].values.shape == u.shape
    assert np.allclose(ds["v_x"].values, (u * res * 30.0).astype(np.float32), equal_nan=True, rtol=1e-6, atol=0)
    assert np.array_equal(ds["corr"].values, c, equal_nan=True)
    with pytest.raises(ValueError):
        b2frames.get_piv(da, window_size=32, overlap=(24, 24), engine="b200", resolution=res, coarse_pass=COARSE50, multipass="spline")

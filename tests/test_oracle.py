"""CPU: the float64 oracle itself - analytic known answers and the reference-side semantics it must follow."""
import numpy as np
import pytest

from oracle import ffpiv_oracle as O
from pyorc_b200 import synth, window


def blob_image(H, W, rng, n=None, sigma=1.6):
    n = n or int(0.03 * H * W)
    px, py, amp = rng.uniform(0, W, n), rng.uniform(0, H, n), rng.uniform(100, 255, n)
    return px, py, amp


def render(px, py, amp, H, W, sigma=1.6):
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.zeros((H, W))
    for x, y, a in zip(px, py, amp):
        # periodic rendering so that a circular shift is exact
        dx = np.minimum(np.abs(xx - x), W - np.abs(xx - x))
        dy = np.minimum(np.abs(yy - y), H - np.abs(yy - y))
        img += a * np.exp(-(dx**2 + dy**2) / (2 * sigma**2))
    return img


@pytest.mark.parametrize("clip", [True, False])
@pytest.mark.parametrize("shift", [(2.3, 3.6), (-1.7, 3.3), (0.0, 0.0), (4.5, -5.25)])
def test_known_subpixel_shift_single_window(clip, shift):
    """One 64x64 periodic window, frame B = frame A shifted by (dy, dx): +u = right, +v = down."""
    O.CLIP_NORMALIZED = clip
    rng = np.random.default_rng(3)
    px, py, amp = blob_image(64, 64, rng, n=60)
    a = render(px, py, amp, 64, 64)
    b = render(px + shift[1], py + shift[0], amp, 64, 64)
    imgs = np.stack([a, b]).astype(np.float32)
    u, v, c, s = O.uv_timestep(imgs, 1, 1, (64, 64), (32, 32))
    assert abs(u[0, 0, 0] - shift[1]) < 0.05 and abs(v[0, 0, 0] - shift[0]) < 0.05
    assert 0.6 < c[0, 0, 0] <= 1.0 and s[0, 0, 0] > 3


def test_zero_signal_window_gives_zero_plane_and_nan_s2n():
    imgs = np.zeros((2, 64, 64), np.uint8)
    _, _, corr = O.cross_corr(imgs, (64, 64), (32, 32))
    assert corr.dtype == np.float32 and corr.shape == (1, 1, 64, 64) and not corr.any()
    u, v, c, s = O.uv_timestep(imgs, 1, 1, (64, 64), (32, 32))
    assert c[0, 0, 0] == 0 and np.isnan(s[0, 0, 0])
    assert np.isnan(u[0, 0, 0]) and np.isnan(v[0, 0, 0])  # argmax 0 -> border -> BORDER_RULE


def test_argmax_first_occurrence_and_border_rules():
    plane = np.zeros((1, 1, 8, 8), np.float32)
    plane[0, 0, 3, 4] = 0.5
    plane[0, 0, 5, 2] = 0.5  # tie: row-major first wins -> (3, 4)
    plane[0, 0, 2, 4] = plane[0, 0, 4, 4] = 0.25
    plane[0, 0, 3, 3] = plane[0, 0, 3, 5] = 0.25
    u, v = O.u_v_displacement(plane, 1, 1)
    assert u[0, 0, 0] == pytest.approx(0.0, abs=1e-6) and v[0, 0, 0] == pytest.approx(-1.0, abs=1e-6)
    edge = np.zeros((1, 1, 8, 8), np.float32)
    edge[0, 0, 0, 5] = 1.0
    old = O.BORDER_RULE
    try:
        O.BORDER_RULE = "nan"
        u, v = O.u_v_displacement(edge, 1, 1)
        assert np.isnan(u).all() and np.isnan(v).all()
        O.BORDER_RULE = "integer"
        u, v = O.u_v_displacement(edge, 1, 1)
        assert u[0, 0, 0] == 1 and v[0, 0, 0] == -4
    finally:
        O.BORDER_RULE = old


def test_corr_is_float32_in_unit_interval_and_centre_is_zero_shift():
    O.CLIP_NORMALIZED = True
    imgs = synth.particle_frames(2, 96, 128, dtype=np.uint8)
    same = np.stack([imgs[0], imgs[0]])
    x, y, corr = O.cross_corr(same, (32, 32), (16, 16))
    assert corr.dtype == np.float32 and corr.min() >= 0 and corr.max() <= 1
    idx = corr.reshape(corr.shape[0], corr.shape[1], -1).argmax(-1)
    assert (idx == 16 * 32 + 16).all()  # auto-correlation peaks at (wy//2, wx//2)
    assert np.array_equal(x, window.get_rect_coordinates((96, 128), (32, 32), (16, 16))[0])
    assert np.array_equal(y, window.get_rect_coordinates((96, 128), (32, 32), (16, 16))[1])


def test_window_geometry_against_survey_table():
    """SURVEY.md §8(a) sizes (OpenPIV field-shape rule)."""
    assert O.get_array_shape((475, 371), (32, 32), (16, 16)) == (28, 22)
    assert O.get_array_shape((1080, 1920), (64, 64), (32, 32)) == (32, 59)
    assert O.get_array_shape((1080, 1920), (32, 32), (24, 24)) == (132, 237)
    assert O.get_array_shape((2160, 3840), (64, 64), (32, 32)) == (66, 119)
    assert O.get_array_shape((4320, 7680), (128, 128), (64, 64)) == (66, 119)
    for dims, ws, ov in [((475, 371), (10, 10), (5, 5)), ((1080, 1920), (64, 64), (32, 32)), ((200, 333), (26, 26), (12, 12))]:
        assert window.get_array_shape(dims, ws, ov) == O.get_array_shape(dims, ws, ov)
        for a, b in zip(window.get_rect_coordinates(dims, ws, ov), O.get_rect_coordinates(dims, ws, ov)):
            assert a.dtype == np.int64 and np.array_equal(a, b)
    assert window.round_to_even((25, 25)) == (26, 26) == O.round_to_even((25, 25))
    assert window.round_to_even((10, 64.0)) == (10, 64)


def test_subwindows_layout_row_major():
    img = np.arange(2 * 40 * 50).reshape(2, 40, 50)
    st = O.subwindows(img, (16, 16), (8, 8))
    nr, nc = O.get_array_shape((40, 50), (16, 16), (8, 8))
    assert st.shape == (2, nr * nc, 16, 16)
    r, c = 2, 3
    assert np.array_equal(st[1, r * nc + c], img[1, r * 8 : r * 8 + 16, c * 8 : c * 8 + 16])


def test_signal_threshold_masks_planes_with_nan():
    imgs = synth.particle_frames(3, 96, 128, dtype=np.uint8)
    imgs[:, :40, :60] = 0
    _, _, corr = O.cross_corr(imgs, (32, 32), (16, 16), signal_threshold=0.5)
    nr, nc = O.get_array_shape((96, 128), (32, 32), (16, 16))
    assert np.isnan(corr[:, 0]).all() and np.isfinite(corr[:, nr * nc - 1]).all()
    u, v, c, s = O.uv_timestep(imgs, nc, nr, (32, 32), (16, 16), signal_threshold=0.5)
    assert np.isnan(c[:, 0, 0]).all() and np.isnan(u[:, 0, 0]).all()


def test_ensemble_matches_manual_average():
    """Ensemble restatement (ffpiv.py:200-376): with thresholds at 0 the mean plane is the plain average."""
    O.CLIP_NORMALIZED = True
    imgs = synth.particle_frames(5, 96, 128, dtype=np.uint8)
    ws, ov = (32, 32), (16, 16)
    nr, nc = O.get_array_shape(imgs.shape[-2:], ws, ov)
    ens = O.Ensemble(nr, nc, ws, ov, corr_min=0.0, s2n_min=0.0, count_min=0.0)
    ens.add_chunk(imgs[:3])
    ens.add_chunk(imgs[2:])
    u, v, cm, sn = ens.finalize()
    _, _, corr = O.cross_corr(imgs, ws, ov)
    u2, v2 = O.u_v_displacement(corr.mean(axis=0, keepdims=True), nr, nc)
    assert np.allclose(u, u2, atol=1e-4, equal_nan=True) and np.allclose(v, v2, atol=1e-4, equal_nan=True)
    assert cm.shape == (1, nr, nc) and sn.shape == (1, nr, nc)


def test_polyphase_identity_behind_the_128x128_kernel():
    """piv_rows128.cuh never runs a 128-point transform: with a_p[m] = a[2m + p] the period-128 circular cross-correlation has
    the polyphase components c_q[n'] = sum_p corr64(a_p, b_{p xor q})[n' + (p and q)], i.e. in the 64x64 frequency domain
    C_q[k] = sum_p conj(A_p[k]) B_{p xor q}[k] exp(+2 pi i (k1 s1 + k2 s2) / 64), s = p and q.  Checked here in float64 against
    the reference formula irfft2(conj(rfft2 a) rfft2 b)."""
    rng = np.random.default_rng(12)
    a, b = rng.standard_normal((128, 128)), rng.standard_normal((128, 128))
    ref = np.fft.irfft2(np.conj(np.fft.rfft2(a)) * np.fft.rfft2(b), s=(128, 128))
    comp = lambda x, p1, p2: x[p1::2, p2::2]
    A = {(p1, p2): np.fft.fft2(comp(a, p1, p2)) for p1 in (0, 1) for p2 in (0, 1)}
    B = {(p1, p2): np.fft.fft2(comp(b, p1, p2)) for p1 in (0, 1) for p2 in (0, 1)}
    k1, k2 = np.meshgrid(np.arange(64), np.arange(64), indexing="ij")
    got = np.empty_like(ref)
    for q1 in (0, 1):
        for q2 in (0, 1):
            C = np.zeros((64, 64), complex)
            for p1 in (0, 1):
                for p2 in (0, 1):
                    s1, s2 = p1 & q1, p2 & q2
                    C += np.conj(A[p1, p2]) * B[p1 ^ q1, p2 ^ q2] * np.exp(2j * np.pi * (k1 * s1 + k2 * s2) / 64)
            cq = np.fft.ifft2(C)
            assert np.abs(cq.imag).max() < 1e-9
            got[q1::2, q2::2] = cq.real
    assert np.abs(got - ref).max() < 1e-9 * np.abs(ref).max() + 1e-9
    # the index map of the kernel's epilogue: reference (fftshifted) row of element m1 of component q1 is (2 m1 + q1 + 64) % 128
    sh = np.fft.fftshift(ref)
    m1, q1, m2, q2 = 5, 1, 40, 0
    assert sh[(2 * m1 + q1 + 64) % 128, (2 * m2 + q2 + 64) % 128] == ref[2 * m1 + q1, 2 * m2 + q2]


def test_displaced_second_pass_needs_two_forward_transforms_per_frame():
    """piv_rows_shift_kernel: pair k correlates the undisplaced window of frame k with the DISPLACED window of frame k+1, pair
    k+1 the undisplaced window of frame k+1 with the displaced one of frame k+2 - the displaced spectrum cannot be reused as
    the next pair's `a` (a shift of the window is not a phase ramp of its spectrum: the content changes at the borders)."""
    rng = np.random.default_rng(3)
    img = rng.standard_normal((96, 96))
    w0 = img[32:64, 32:64]
    w_shift = img[35:67, 30:62]                       # the window displaced by (+3, -2)
    k1, k2 = np.meshgrid(np.fft.fftfreq(32), np.fft.fftfreq(32), indexing="ij")
    ramp = np.fft.fft2(w0) * np.exp(2j * np.pi * (3 * k1 - 2 * k2))
    assert np.abs(np.fft.ifft2(ramp).real - w_shift).max() > 0.5     # circular shift != displaced window


@pytest.mark.parametrize("n,P", [((26, 26), 64), ((10, 14), 32), ((30, 18), 64), ((32, 32), 64)])
def test_padded_embedding_identity_behind_the_padded_rows_kernels(n, P):
    """piv_rows.cuh "Padded mode": the period-n circular correlation of the reference equals, at the lags 0 .. n-1, the circular
    correlation on a P x P plane (P >= 2 n) of the zero-padded window `a` with the window `b` tiled 2 x 2 - and the tiled plane's
    spectrum is the zero-padded one times T(k) = (1 + w^(ny k1)) (1 + w^(nx k2)), w = exp(-2 pi i / P), so ONE forward transform
    per window and frame serves both roles.  Lag q >= n/2 is the reference's negative lag q - n: its fftshifted plane is
    element (q + n/2) % n <- lag q.  Checked in float64 against irfft2(conj(rfft2 a) rfft2 b)."""
    ny, nx = n
    rng = np.random.default_rng(5)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    ref = np.fft.irfft2(np.conj(np.fft.rfft2(a)) * np.fft.rfft2(b), s=n)
    za, zb = np.zeros((P, P)), np.zeros((P, P))
    za[:ny, :nx], zb[:ny, :nx] = a, b
    k1, k2 = np.meshgrid(np.arange(P), np.arange(P), indexing="ij")
    T = (1 + np.exp(-2j * np.pi * ny * k1 / P)) * (1 + np.exp(-2j * np.pi * nx * k2 / P))
    tiled = np.zeros((P, P))
    tiled[: 2 * ny, : 2 * nx] = np.tile(b, (2, 2))
    assert np.abs(np.fft.fft2(zb) * T - np.fft.fft2(tiled)).max() < 1e-9
    plane = np.fft.ifft2(np.conj(np.fft.fft2(za)) * np.fft.fft2(zb) * T)
    assert np.abs(plane.imag).max() < 1e-9
    assert np.abs(plane.real[:ny, :nx] - ref).max() < 1e-9
    sh = np.fft.fftshift(ref)
    q1, q2 = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    assert np.array_equal(sh[(q1 + ny // 2) % ny, (q2 + nx // 2) % nx], ref)


@pytest.mark.parametrize("n", [(50, 50), (34, 34), (64, 48), (36, 20), (62, 62)])
def test_padded_polyphase_identity_behind_the_padded_128_plane_kernel(n):
    """piv_rows128.cuh "Padded mode" (even windows of 34 .. 64 px, uint8 and float32 frames): the embedding above on the 128 x 128
    plane, computed through the four 64 x 64 polyphase components.  A component holds (ny/2) x (nx/2) samples of the window, the
    tiling shift n = 2 (n/2) stays inside a component, so T(k) = (1 + w64^(k1 ny/2)) (1 + w64^(k2 nx/2)) multiplies every C_q alike;
    lag (2 m1 + q1, 2 m2 + q2) of component q is the reference's plane at that lag."""
    ny, nx = n
    rng = np.random.default_rng(9)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    ref = np.fft.irfft2(np.conj(np.fft.rfft2(a)) * np.fft.rfft2(b), s=n)
    za, zb = np.zeros((128, 128)), np.zeros((128, 128))
    za[:ny, :nx], zb[:ny, :nx] = a, b
    comp = lambda x, p1, p2: x[p1::2, p2::2]
    A = {(p1, p2): np.fft.fft2(comp(za, p1, p2)) for p1 in (0, 1) for p2 in (0, 1)}
    B = {(p1, p2): np.fft.fft2(comp(zb, p1, p2)) for p1 in (0, 1) for p2 in (0, 1)}
    k1, k2 = np.meshgrid(np.arange(64), np.arange(64), indexing="ij")
    T = (1 + np.exp(-2j * np.pi * k1 * (ny // 2) / 64)) * (1 + np.exp(-2j * np.pi * k2 * (nx // 2) / 64))
    got = np.empty((128, 128))
    for q1 in (0, 1):
        for q2 in (0, 1):
            C = np.zeros((64, 64), complex)
            for p1 in (0, 1):
                for p2 in (0, 1):
                    s1, s2 = p1 & q1, p2 & q2
                    C += np.conj(A[p1, p2]) * B[p1 ^ q1, p2 ^ q2] * np.exp(2j * np.pi * (k1 * s1 + k2 * s2) / 64)
            cq = np.fft.ifft2(C * T)
            assert np.abs(cq.imag).max() < 1e-9
            got[q1::2, q2::2] = cq.real
    assert np.abs(got[:ny, :nx] - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())

"""BASELINE.json's configurations at FULL size on the GPU, checked through size-independent properties plus an oracle
spot check on randomly sampled windows (the oracle cannot process a whole 1080p x 100 stack in seconds, a few hundred
windows it can):

  * spot check   random (pair, window) samples: u, v, corr_max, s2n against the float64 oracle on those windows
  * chunking     a stack processed whole == processed in two overlapping halves (1-frame halo), bit for bit
                 (the property behind pyorc's chunk loop, ffpiv.py:399-440, and behind frame-pair sharding)
  * reversal     swapping the two frames of a pair mirrors the correlation plane: same corr_max / s2n, negated u, v
  * identity     a frame against itself: u = v = 0, corr_max = 1
  * truth        the imposed synthetic displacement field is recovered (median error well below a pixel)
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
    e.set_option("clip_normalized", 0.0)
    e.set_option("kernel_variant", 0.0)
    e.set_option("run_len", 0.0)
    O.CLIP_NORMALIZED = False
    yield e
    e.close()


def frames_on_device(n, H, W):
    import torch

    return synth.particle_frames_torch(n, H, W, torch.device("cuda", 0), dtype="uint8")


def run(engine, d_frames, ws, ov):
    return tuple(t.cpu().numpy() for t in engine.pairs(d_frames, ws, ov))


def spot_check(host, res, ws, ov, n_samples=256, seed=1):
    """Oracle on sampled windows only: the same float64 ncc + float32 plane + peak fit as oracle.uv_timestep."""
    u, v, c, s = res
    n_pairs, nr, nc = u.shape
    rng = np.random.default_rng(seed)
    k = rng.integers(0, n_pairs, n_samples)
    r = rng.integers(0, nr, n_samples)
    q = rng.integers(0, nc, n_samples)
    y0, x0 = O.window_origins(host.shape[-2:], ws, ov)
    wa = np.stack([host[kk, y0[rr]:y0[rr] + ws[0], x0[qq]:x0[qq] + ws[1]] for kk, rr, qq in zip(k, r, q)])
    wb = np.stack([host[kk + 1, y0[rr]:y0[rr] + ws[0], x0[qq]:x0[qq] + ws[1]] for kk, rr, qq in zip(k, r, q)])
    corr = O.ncc(wa, wb).astype(np.float32)
    cmax = corr.max(axis=(-1, -2))
    with np.errstate(all="ignore"):
        s2n = cmax / corr.mean(axis=(-1, -2))
    pk = O.peak_position(corr)
    ov_, ou_ = pk[:, 0] - ws[0] // 2, pk[:, 1] - ws[1] // 2
    gu, gv, gc, gs = u[k, r, q], v[k, r, q], c[k, r, q], s[k, r, q]
    assert np.array_equal(np.isnan(gu), np.isnan(ou_))
    same = np.isfinite(ou_) & (np.abs(np.round(gu) - np.round(ou_)) + np.abs(np.round(gv) - np.round(ov_)) < 0.5)
    assert same.sum() >= 0.99 * np.isfinite(ou_).sum()
    assert np.abs(gu[same] - ou_[same]).max() <= 2e-3 and np.abs(gv[same] - ov_[same]).max() <= 2e-3
    assert np.abs(gc - cmax).max() <= 5e-6
    ok = np.isfinite(s2n) & (s2n != 0)
    assert (np.abs(gs[ok] - s2n[ok]) / s2n[ok]).max() <= 2e-5


def check_chunking(engine, d, ws, ov, whole):
    h = d.shape[0] // 2
    a = run(engine, d[: h + 1], ws, ov)
    b = run(engine, d[h:], ws, ov)
    for w, x, y in zip(whole, a, b):
        assert np.array_equal(np.concatenate([x, y]), w, equal_nan=True)


def check_reversal_and_identity(engine, d, ws, ov):
    import torch

    fwd = run(engine, d[:2], ws, ov)
    bwd = run(engine, torch.stack([d[1], d[0]]), ws, ov)
    fin = np.isfinite(fwd[0]) & np.isfinite(bwd[0])
    # the mirrored plane has the same samples, so max and mean agree to rounding; an even-sized plane's mirror moves the
    # Nyquist row/column, which can turn an interior peak into a border peak - compare where both are finite
    assert np.abs(fwd[2] - bwd[2]).max() <= 2e-6
    assert fin.mean() > 0.9
    assert np.abs(fwd[0][fin] + bwd[0][fin]).max() <= 2e-3 and np.abs(fwd[1][fin] + bwd[1][fin]).max() <= 2e-3
    same = run(engine, torch.stack([d[0], d[0]]), ws, ov)
    live = np.isfinite(same[0])
    assert live.mean() > 0.99
    assert np.abs(same[0][live]).max() <= 1e-4 and np.abs(same[1][live]).max() <= 1e-4
    assert np.abs(same[2][live] - 1.0).max() <= 1e-5


def check_truth(res, H, W, ws, ov):
    u, v = res[0], res[1]
    y0, x0 = O.window_origins((H, W), ws, ov)
    yc, xc = np.meshgrid(y0 + ws[0] / 2, x0 + ws[1] / 2, indexing="ij")
    dx, dy = synth.displacement_field(H, W, yc, xc)
    assert np.nanmedian(np.abs(u - dx[None])) < 0.15 and np.nanmedian(np.abs(v - dy[None])) < 0.15


def test_config0_ngwerere_32x32(engine):
    """BASELINE.json configs[0]: the Ngwerere sample clip, 2 frame pairs, 32x32 windows, 50 % overlap - the three orthorectified
    frames that reach ffpiv in pyorc's own test (tests/golden/ngwerere_proj.npz, made by the reference's decode / projection
    code), EVERY window against the float64 oracle, from host memory as pyorc hands them over (475 x 371 uint8: the device copy is
    re-pitched), whole and in two chunks with the 1-frame halo."""
    import os

    frames = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ngwerere_proj.npz"))["frames"]
    assert frames.shape == (3, 475, 371) and frames.dtype == np.uint8
    ws, ov = (32, 32), (16, 16)
    nr, nc = O.get_array_shape(frames.shape[-2:], ws, ov)
    assert (nr, nc) == (28, 22)                                     # 616 windows per pair
    u, v, c, s = O.uv_timestep(frames, nc, nr, ws, ov)
    gu, gv, gc, gs = (np.array(a) for a in engine.pairs(frames, ws, ov))
    assert gu.shape == (2, nr, nc)
    assert np.array_equal(np.isnan(gu), np.isnan(u)) and np.array_equal(np.isnan(gs), np.isnan(s))
    fin = np.isfinite(u)
    same = fin & (np.abs(np.round(gu) - np.round(u)) + np.abs(np.round(gv) - np.round(v)) < 0.5)
    assert same.sum() >= 0.999 * fin.sum()
    assert np.abs(gu[same] - u[same]).max() <= 2e-3 and np.abs(gv[same] - v[same]).max() <= 2e-3
    assert np.sqrt(np.mean((gu[same] - u[same]) ** 2)) <= 2e-4 and np.sqrt(np.mean((gv[same] - v[same]) ** 2)) <= 2e-4
    assert np.abs(gc - c).max() <= 5e-6
    ok = np.isfinite(s) & (s != 0)
    assert (np.abs(gs[ok] - s[ok]) / np.abs(s[ok])).max() <= 2e-5
    a = engine.pairs(frames[:2], ws, ov)
    b = engine.pairs(frames[1:], ws, ov)
    for w, x, y in zip((gu, gv, gc, gs), a, b):
        assert np.array_equal(np.concatenate([x, y]), w, equal_nan=True)


def test_config1_1080p_100_pairs_64x64(engine):
    """BASELINE.json configs[1]: synthetic 1080p, 100 frame pairs, 64x64 windows, 50 % overlap (the bench workload)."""
    H, W, ws, ov = 1080, 1920, (64, 64), (32, 32)
    d = frames_on_device(101, H, W)
    whole = run(engine, d, ws, ov)
    assert whole[0].shape == (100, 32, 59)
    spot_check(d.cpu().numpy(), whole, ws, ov, n_samples=384)
    check_chunking(engine, d, ws, ov, whole)
    check_reversal_and_identity(engine, d, ws, ov)
    check_truth(whole, H, W, ws, ov)


def test_config2_1080p_32x32_75_percent_overlap(engine):
    """configs[2] geometry (single pass; the 2-pass deformation has no reference): 132 x 237 windows/pair."""
    H, W, ws, ov = 1080, 1920, (32, 32), (24, 24)
    d = frames_on_device(21, H, W)
    whole = run(engine, d, ws, ov)
    assert whole[0].shape == (20, 132, 237)
    spot_check(d.cpu().numpy(), whole, ws, ov, n_samples=384)
    check_chunking(engine, d, ws, ov, whole)
    check_reversal_and_identity(engine, d, ws, ov)


def test_config3_4k_64x64(engine):
    """configs[3] frame size (3840 x 2160), a shard of 8 pairs: 66 x 119 windows/pair."""
    H, W, ws, ov = 2160, 3840, (64, 64), (32, 32)
    d = frames_on_device(9, H, W)
    whole = run(engine, d, ws, ov)
    assert whole[0].shape == (8, 66, 119)
    spot_check(d.cpu().numpy(), whole, ws, ov, n_samples=256)
    check_chunking(engine, d, ws, ov, whole)
    check_truth(whole, H, W, ws, ov)


def test_config4_8k_128x128(engine):
    """configs[4] frame size (7680 x 4320) and 128x128 windows, a shard of 3 pairs: 66 x 119 windows/pair."""
    H, W, ws, ov = 4320, 7680, (128, 128), (64, 64)
    d = frames_on_device(4, H, W)
    whole = run(engine, d, ws, ov)
    assert whole[0].shape == (3, 66, 119)
    spot_check(d.cpu().numpy(), whole, ws, ov, n_samples=96)
    check_chunking(engine, d, ws, ov, whole)
    check_reversal_and_identity(engine, d, ws, ov)

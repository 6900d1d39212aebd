"""The reference's own golden vectors for the hot path: pyorc tests/test_frames.py:139-153 (`test_get_piv`).

tests/golden/ngwerere_proj.npz holds the three orthorectified Ngwerere frames that reach ffpiv in that test, generated
by the reference's own code (tests/golden/make_ngwerere_golden.py).  CPU: the oracle must reproduce both pinned vectors
(this is what pins the oracle).  GPU: the CUDA engine, through the same get_piv-style binding, must reproduce them too.
"""
import os
import warnings

import numpy as np
import pytest

from oracle import ffpiv_oracle as O
from pyorc_b200 import _xr

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ngwerere_proj.npz")
PIN_ATOL = 3e-8   # the reference prints its pins with 8 significant digits; its own tolerance is atol=0 + rtol 1e-5


@pytest.fixture(scope="module")
def ngwerere():
    d = np.load(FIX)
    return d["frames"], d["time_s"], float(d["resolution"]), d["pinned_vx_timestep"], d["pinned_vx_ensemble"]


def test_fixture_matches_reference_facts(ngwerere):
    frames, t, res, _, _ = ngwerere
    assert frames.shape == (3, 475, 371) and frames.dtype == np.uint8   # tests/test_frames.py:36 (0.01 m resolution)
    assert res == 0.01 and np.allclose(np.diff(t), 1 / 30)


def test_oracle_reproduces_pinned_vx_per_timestep(ngwerere):
    frames, t, res, pin, _ = ngwerere
    O.CLIP_NORMALIZED = False
    ws, ov = (10, 10), (5, 5)                                 # window_size=10 -> overlap int(round(10)/2) (frames.py:170-171)
    nr, nc = O.get_array_shape(frames.shape[1:], ws, ov)
    u, v, c, s = O.uv_timestep(frames, nc, nr, ws, ov)
    vx = (u * res / np.diff(t)[:, None, None]).astype(np.float32)   # ffpiv.py:418
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        got = np.nanmean(vx, axis=0).flatten()[-4:]               # piv.mean(dim="time")["v_x"].values.flatten()[-4:]
    assert np.allclose(got, pin, rtol=0, atol=PIN_ATOL), (got, pin)


def test_oracle_reproduces_pinned_vx_ensemble(ngwerere):
    frames, t, res, _, pin = ngwerere
    O.CLIP_NORMALIZED = False
    ws, ov = (10, 10), (5, 5)
    nr, nc = O.get_array_shape(frames.shape[1:], ws, ov)
    ens = O.Ensemble(nr, nc, ws, ov, corr_min=0, s2n_min=0, count_min=0)
    ens.add_chunk(frames)
    u, v, cm, sn = ens.finalize()
    got = (u * res / np.diff(t).mean()).astype(np.float32).flatten()[-4:]
    assert np.allclose(got, pin, rtol=0, atol=PIN_ATOL), (got, pin)


def test_clipped_normalisation_is_ruled_out_by_the_pin(ngwerere):
    frames, t, res, pin, _ = ngwerere
    O.CLIP_NORMALIZED = True
    try:
        nr, nc = O.get_array_shape(frames.shape[1:], (10, 10), (5, 5))
        u, *_ = O.uv_timestep(frames, nc, nr, (10, 10), (5, 5))
        vx = (u * res / np.diff(t)[:, None, None]).astype(np.float32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            got = np.nanmean(vx, axis=0).flatten()[-4:]
        assert np.abs(got - pin).max() > 1e-2
    finally:
        O.CLIP_NORMALIZED = False


def _as_frames(frames, t, res):
    H, W = frames.shape[1:]
    y = np.flipud(np.linspace(res / 2, res * (H - 0.5), H))
    x = np.linspace(res / 2, res * (W - 0.5), W)
    return _xr.DataArray(frames, ("time", "y", "x"), {"time": t, "y": y, "x": x})


@pytest.mark.gpu
@pytest.mark.parametrize("ensemble_corr", [False, True])
def test_gpu_engine_reproduces_pinned_vx(ngwerere, ensemble_corr):
    """`test_get_piv` of the reference, run through the B200 binding: window_size=10, s2n_min=corr_min=count_min=0."""
    from pyorc_b200 import frames as b2frames

    frames, t, res, pin_ts, pin_ens = ngwerere
    piv = b2frames.get_piv(_as_frames(frames, t, res), window_size=10, ensemble_corr=ensemble_corr, engine="b200",
                           resolution=res, s2n_min=0, corr_min=0, count_min=0)
    piv_mean = piv.mean(dim="time", keep_attrs=True)
    got = piv_mean["v_x"].values.flatten()[-4:]
    pin = pin_ens if ensemble_corr else pin_ts
    # fp32 engine vs the float64 reference: 2e-6 m/s = 7e-6 px/frame
    assert np.allclose(got, pin, rtol=0, atol=2e-6, equal_nan=True), (got, pin)

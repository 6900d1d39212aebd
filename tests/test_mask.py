"""Velocimetry mask stack + result packing: oracle restatement of pyorc/api/mask.py checked on the
CPU against hand-worked cases and the reference's quirks; CUDA kernels (through the C ABI) checked against the oracle on
the GPU - masks bit for bit (float32 arithmetic in numpy's operation order), atan2 up to the last ulp at the threshold."""
import warnings

import numpy as np
import pytest

from oracle import mask_oracle as MO

F32 = np.float32


def fields(T=37, ny=23, nx=31, seed=5, nan_frac=0.12):
    rng = np.random.default_rng(seed)
    vx = (0.6 + 0.35 * rng.standard_normal((T, ny, nx))).astype(F32)
    vy = (-0.2 + 0.25 * rng.standard_normal((T, ny, nx))).astype(F32)
    c = rng.uniform(0, 1, (T, ny, nx)).astype(F32)
    s = rng.uniform(1, 40, (T, ny, nx)).astype(F32)
    bad = rng.uniform(size=(T, ny, nx)) < nan_frac
    bad[:, 3, 4] = True            # a location that is never valid
    bad[:, 5, 6] = False
    bad[1:, 7, 8] = True           # a location with one valid sample (std = 0)
    for a in (vx, vy, c, s):
        a[bad] = np.nan
    vx[T // 2, 10, 10], vy[T // 2, 10, 10] = 0.0, 0.0      # zero speed, atan2(0, 0)
    return vx, vy, c, s


# ---------------------------------------------------------------------------------------------------------------------------
# CPU: the oracle itself
# ---------------------------------------------------------------------------------------------------------------------------
def test_oracle_stack_window_keeps_the_reference_quirk():
    # helpers.py:675-676: x strides -wdw..wdw inclusive, y strides -wdw..wdw-1 (range excludes the maximum)
    assert MO.strides(1) == [(-1, -1), (-1, 0), (0, -1), (0, 0), (1, -1), (1, 0)]
    assert len(MO.strides(2)) == 5 * 4
    assert MO.strides(1, wdw_y_max=2) == [(x, y) for x in (-1, 0, 1) for y in (-1, 0, 1)]


def test_oracle_shift_is_xarray_shift():
    a = np.arange(12, dtype=F32).reshape(3, 4)
    s = MO.shift_yx(a, 1, 0)        # da.shift(x=1): data moves to higher x, NaN enters at x = 0
    assert np.isnan(s[:, 0]).all() and np.array_equal(s[:, 1:], a[:, :-1])
    s = MO.shift_yx(a, 0, -1)       # da.shift(y=-1): data moves to lower y
    assert np.isnan(s[-1]).all() and np.array_equal(s[:-1], a[1:])
    assert np.isnan(MO.shift_yx(a, 7, 0)).all()


def test_oracle_rolling_window_convention():
    # xarray pads window // 2 before and window - 1 - window // 2 after (Variable.rolling_window, center=True)
    s = np.arange(8, dtype=F32)[:, None, None]
    r5 = MO.rolling_max_centered(s, 5)[:, 0, 0]
    assert np.isnan(r5[[0, 1, 6, 7]]).all() and np.array_equal(r5[2:6], [4, 5, 6, 7])
    r4 = MO.rolling_max_centered(s, 4)[:, 0, 0]
    assert np.isnan(r4[[0, 1, 7]]).all() and np.array_equal(r4[2:7], [3, 4, 5, 6, 7])


def test_oracle_masks_hand_cases():
    vx = np.array([[[3.0]], [[0.05]], [[np.nan]], [[30.0]]], F32)
    vy = np.array([[[4.0]], [[0.0]], [[1.0]], [[40.0]]], F32)
    assert MO.minmax(vx, vy, 0.1, 5.5).ravel().tolist() == [True, False, False, False]
    assert MO.count(vx, 0.7).tolist() == [[True]] and MO.count(vx, 0.75).tolist() == [[False]]     # 3 of 4 valid
    # variance: the reference clamps the mean from below with 1e30, so every location with data passes
    assert MO.variance(vx, vy, tolerance=1e-6).tolist() == [[True]]
    assert MO.variance(np.full((3, 1, 1), np.nan, F32), vy[:3], mode="and").tolist() == [[False]]
    # angle: flow to the right (v_x > 0, v_y = 0) is pi / 2
    assert MO.angle(np.array([1.0], F32), np.array([0.0], F32)).tolist() == [True]
    assert MO.angle(np.array([-1.0], F32), np.array([0.0], F32)).tolist() == [False]


def test_oracle_window_replace_fills_from_neighbours():
    a = np.ones((1, 5, 5), F32)
    a[0, 2, 2] = np.nan
    (out,) = MO.window_replace([a])
    assert out[0, 2, 2] == 1.0 and np.isfinite(out).all()
    b = np.full((1, 5, 5), np.nan, F32)
    b[0, 0, 0] = 2.0
    (out,) = MO.window_replace([b], iter=2)
    # strides (xs, ys) with ys in {-1, 0}: a value spreads to x +- 1 and to y - 1 ... (result[i] = a[i - ys]) -> lower y only
    assert out[0, 0, 1] == 2.0 and np.isnan(out[0, 1, 0]) and out[0, 0, 2] == 2.0


def test_oracle_encoding_round_trip():
    a = np.array([0.0, 0.004, 0.005, 0.015, -0.025, 1.23456, np.nan, 400.0, -400.0], F32)
    q = MO.encode_int16(a)
    assert q.dtype == np.int16
    assert q.tolist() == [0, 0, 0, 2, -2, 123, -9999, 32767, -32768]    # float32 quotients: 0.5 -> 0 (half to even), 1.5000001 -> 2, -2.5 -> -2
    d = MO.decode_int16(q)
    assert np.isnan(d[6]) and abs(d[5] - 1.23) < 1e-6
    ok = np.isfinite(a) & (np.abs(a) < 300)
    assert np.abs(d[ok] - a[ok]).max() <= 0.005 + 1e-6


# ---------------------------------------------------------------------------------------------------------------------------
# CPU: THE PIN - the reference's own example output of its mask stack (tests/golden/ngwerere_masks.npz, made from the two netCDF
# files the reference ships by tests/golden/make_mask_golden.py)
# ---------------------------------------------------------------------------------------------------------------------------
def _reference_example():
    import os

    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ngwerere_masks.npz"))
    shape = tuple(int(v) for v in d["shape"])
    kept = np.unpackbits(d["kept_bits"])[: int(np.prod(shape))].reshape(shape).astype(bool)
    return d, kept


def _notebook_03_mask_sequence(fields):
    """examples/03_Plotting_and_masking_velocimetry_results.ipynb, cell 10: seven masks with pyorc's defaults, each applied in place
    (`ds[var].where(mask)` on every variable, mask.py:84-88) before the next one is computed; `angle(...)` is called WITHOUT
    inplace there and changes nothing; `window_mean(..., reduce_time=True)` decides on the time-mean fields (mask.py:51-52)."""
    f = [np.array(a, copy=True) for a in fields]
    steps = [lambda f: MO.corr(f[2]),
             lambda f: MO.minmax(f[0], f[1]),
             lambda f: MO.rolling(f[0], f[1]),
             lambda f: MO.outliers(f[0], f[1]),
             lambda f: MO.variance(f[0], f[1]),
             lambda f: MO.count(f[0]),
             lambda f: MO.window_mean(MO.time_stats(f[0])[0], MO.time_stats(f[1])[0], tolerance=0.5, wdw=2)]
    kept_after = []
    for fn in steps:
        f = MO.apply_masks(f, [fn(f)])
        kept_after.append(float(np.isfinite(f[0]).mean()))
    return f, kept_after


def test_oracle_reproduces_the_reference_mask_example_exactly():
    """pyorc ships the input (ngwerere_piv.nc) AND the output (ngwerere_masked.nc) of its own mask stack - xarray + pyorc/api/mask.py
    on 125 x 59 x 66 values.  The numpy restatement must leave exactly the same 93 824 of 486 750 values standing (80.7 % are
    masked away, so every one of corr, minmax, rolling, outliers, count and window_mean decides thousands of values; `variance`
    must change nothing - the reference's `np.maximum(mean, 1e30)` quirk).  This pins oracle/mask_oracle.py: the float32
    arithmetic and comparison semantics, skipna means / std (ddof 0), the centred rolling maximum, the stack_window strides
    (including the missing last y stride), reduce_time and the in-place sequence, and the CF decode in front of it."""
    d, kept = _reference_example()
    assert float(d["scale_factor"]) == MO.SCALE and int(d["fill_value"]) == MO.FILL      # const.py:80-83 as found in the files
    fields = [MO.decode_int16(d[k]) for k in ("v_x", "v_y", "corr")]
    assert all(a.dtype == F32 for a in fields)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        out, kept_after = _notebook_03_mask_sequence(fields)
    got = np.isfinite(out[0])
    assert got.sum() == kept.sum() == 93824
    assert np.array_equal(got, kept)                                   # not one value differs
    assert np.array_equal(np.isfinite(out[2]), kept)                   # the same mask on every variable
    assert kept_after[4] == kept_after[3]                              # variance: no effect (the 1e30 quirk, mask.py:273-274)
    assert kept_after[0] > kept_after[1] > kept_after[2] > kept_after[3] > kept_after[5] > kept_after[6]
    # survivors keep their packed value through decode -> where -> encode (what xarray wrote into ngwerere_masked.nc)
    for k, a in zip(("v_x", "v_y", "corr"), out):
        q = MO.encode_int16(a)
        assert np.array_equal(q[kept], d[k][kept]) and (q[~kept] == MO.FILL).all()


def test_mask_fixture_is_what_the_reference_files_hold():
    """Where /root/reference exists (this container), the committed fixture is re-derived from the two netCDF4 files with the
    minimal HDF5 parser (tests/golden/h5min.py): packed fields, surviving values, CF attributes."""
    import os
    import sys

    ref = "/root/reference/examples/ngwerere"
    if not os.path.exists(os.path.join(ref, "ngwerere_masked.nc")):
        pytest.skip("/root/reference is not present (GPU box)")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import h5min

    d, kept = _reference_example()
    hp, hm = h5min.H5(os.path.join(ref, "ngwerere_piv.nc")), h5min.H5(os.path.join(ref, "ngwerere_masked.nc"))
    np_, nm = h5min.dataset_names(hp), h5min.dataset_names(hm)
    assert {"v_x", "v_y", "corr", "s2n", "time", "y", "x"} <= set(np_) and set(np_) == set(nm)
    for k in ("v_x", "v_y", "corr"):
        a = hp.read(np_[k])
        assert a.dtype == np.int16 and np.array_equal(a, d[k])
        m = hm.read(nm[k])
        assert np.array_equal(m != -9999, kept) and np.array_equal(m[kept], a[kept])
    assert [float(np.asarray(v).reshape(-1)[0]) for v in h5min.dense_attributes(hm, "scale_factor")] == [0.01] * 4
    fills = [v for v in h5min.dense_attributes(hm, "_FillValue") if np.asarray(v).dtype == np.int16]
    assert len(fills) == 4 and all(int(np.asarray(v).reshape(-1)[0]) == -9999 for v in fills)
    assert np.allclose(hp.read(np_["time"]), d["time"])


def test_reference_mask_example_distinguishes_float32_from_float64():
    """The same sequence in float64 arithmetic misses the reference's output in a few dozen places (values that sit exactly on a
    threshold after the 0.01 quantisation): the pin is sharp enough to see the arithmetic type - xarray decodes the int16 fields
    to float32 and pyorc's masks stay in float32."""
    d, kept = _reference_example()
    f32 = MO.F32
    try:
        MO.F32 = np.float64
        fields = [np.where(d[k] == MO.FILL, np.nan, d[k].astype(np.float64) * 0.01) for k in ("v_x", "v_y", "corr")]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            out, _ = _notebook_03_mask_sequence(fields)
    finally:
        MO.F32 = f32
    diff = int((np.isfinite(out[0]) != kept).sum())
    assert 0 < diff < 200, diff


# ---------------------------------------------------------------------------------------------------------------------------
# CPU: the accessor's wrapper rules need no device until a kernel is called
# ---------------------------------------------------------------------------------------------------------------------------
def make_ds(vx, vy, c, s, time=True):
    from pyorc_b200 import _xr

    dims = ("time", "y", "x") if time else ("y", "x")
    coords = {"y": np.arange(vx.shape[-2]) * 0.5, "x": np.arange(vx.shape[-1]) * 0.5}
    if time:
        coords["time"] = np.arange(vx.shape[0]) / 30.0
    return _xr.Dataset({"v_x": (dims, vx), "v_y": (dims, vy), "corr": (dims, c), "s2n": (dims, s)}, coords)


def test_accessor_rejects_non_velocimetry_and_missing_time():
    from pyorc_b200 import _xr
    from pyorc_b200.mask import Masks

    with pytest.raises(AssertionError):
        Masks(_xr.Dataset({"v_x": (("y", "x"), np.zeros((2, 2), F32))}, {}))
    vx, vy, c, s = fields(T=1, ny=12, nx=12)
    ds2 = make_ds(vx[0], vy[0], c[0], s[0], time=False)
    with pytest.raises(AssertionError, match='requires dimension "time"'):
        Masks(ds2).count()
    ds1 = make_ds(vx, vy, c, s)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m = Masks(ds1).outliers()          # a single time step: warning + all-True [y, x] mask, no kernel needed
    assert any("multiple timesteps" in str(x.message) for x in w)
    assert m.values.shape == (12, 12) and m.values.all() and m.dims == ("y", "x")


# ---------------------------------------------------------------------------------------------------------------------------
# GPU: kernels against the oracle
# ---------------------------------------------------------------------------------------------------------------------------
gpu = pytest.mark.gpu


def same(got, want):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape and got.dtype == want.dtype, (got.shape, want.shape, got.dtype, want.dtype)
    assert np.array_equal(got, want), f"{(got != want).sum()} of {got.size} differ"


@gpu
def test_gpu_elementwise_masks():
    from pyorc_b200 import mask as M

    vx, vy, c, s = fields()
    same(M.minmax(vx, vy, 0.3, 0.9), MO.minmax(vx, vy, 0.3, 0.9))
    same(M.minmax(vx, vy), MO.minmax(vx, vy))
    same(M.corr(c, 0.4), MO.corr(c, 0.4))
    same(M.s2n(s), MO.s2n(s))
    got, want = M.angle(vx, vy), MO.angle(vx, vy)
    a = np.arctan2(vx, vy)
    with np.errstate(invalid="ignore"):
        edge = np.abs(np.abs(a - F32(0.5 * np.pi)) - F32(0.25 * np.pi)) < 1e-6     # atan2f may differ in the last ulp
    assert np.array_equal(got[~edge], want[~edge]) and want.sum() > 100
    got = M.angle(vx, vy, angle_expected=np.pi, angle_tolerance=0.5)
    want = MO.angle(vx, vy, angle_expected=np.pi, angle_tolerance=0.5)
    assert (got != want).sum() <= 2


@gpu
def test_gpu_time_statistics_match_numpy_bit_for_bit():
    from pyorc_b200 import mask as M

    vx, vy, c, s = fields(T=101)
    cnt, mean, std = M.time_stats(vx)
    om, os_ = MO.time_stats(vx)
    assert np.array_equal(cnt, (~np.isnan(vx)).sum(axis=0))
    assert np.array_equal(mean, om, equal_nan=True)
    assert np.array_equal(std, os_, equal_nan=True)
    assert np.isnan(mean[3, 4]) and std[7, 8] == 0.0
    same(M.count(vx, 0.9), MO.count(vx, 0.9))
    same(M.count(vx), MO.count(vx))
    # the segmented kernel (16 <= T <= 1024, loads parallel in time) and the one-thread-per-location kernel give the same bits
    for T in (12, 16, 33, 64, 1024, 1030):
        a = fields(T=T, ny=12, nx=45, seed=T)[0]
        cnt, mean, std = M.time_stats(a)
        om, os_ = MO.time_stats(a)
        assert np.array_equal(cnt, (~np.isnan(a)).sum(axis=0)), T
        assert np.array_equal(mean, om, equal_nan=True) and np.array_equal(std, os_, equal_nan=True), T


@gpu
@pytest.mark.parametrize("mode", ["or", "and"])
def test_gpu_time_masks(mode):
    from pyorc_b200 import mask as M

    vx, vy, c, s = fields(T=64)
    same(M.outliers(vx, vy, 1.0, mode), MO.outliers(vx, vy, 1.0, mode))
    same(M.outliers(vx, vy, 2.5, mode), MO.outliers(vx, vy, 2.5, mode))
    same(M.variance(vx, vy, 5, mode), MO.variance(vx, vy, 5, mode))
    for wdw in (5, 4, 1, 9):
        same(M.rolling(vx, vy, wdw, 0.5), MO.rolling(vx, vy, wdw, 0.5))
    same(M.rolling(vx, vy, 5, 0.9), MO.rolling(vx, vy, 5, 0.9))


@gpu
def test_gpu_window_masks():
    from pyorc_b200 import mask as M

    vx, vy, c, s = fields(T=9, nan_frac=0.3)
    for kw in ({}, {"wdw": 2}, {"wdw": 1, "wdw_y_max": 2}, {"wdw": 1, "wdw_x_min": -3, "wdw_x_max": 0}):
        same(M.window_nan(vx, 0.7, **kw), MO.window_nan(vx, 0.7, **kw))
        for mode in ("or", "and"):
            same(M.window_mean(vx, vy, 0.7, mode=mode, **kw), MO.window_mean(vx, vy, 0.7, mode=mode, **kw))
        got = M.window_replace([vx, vy, c, s], iter=2, **kw)
        want = MO.window_replace([vx, vy, c, s], iter=2, **kw)
        for g, w in zip(got, want):
            assert np.array_equal(g, w, equal_nan=True)
    same(M.window_nan(vx[0], 0.5), MO.window_nan(vx[0], 0.5))       # [y, x] input


@gpu
def test_gpu_apply_and_packing():
    from pyorc_b200 import mask as M

    vx, vy, c, s = fields()
    m1, m2 = MO.minmax(vx, vy, 0.3, 0.9), MO.count(vx, 0.85)
    got = M.apply_masks([vx, vy, c, s], [m1, m2])
    want = MO.apply_masks([vx, vy, c, s], [m1, m2])
    for g, w in zip(got, want):
        assert np.array_equal(g, w, equal_nan=True)
    big = np.concatenate([vx.ravel(), np.array([400.0, -400.0, np.inf, 0.005, 0.015, -0.025], F32)])
    q = M.encode_int16(big)
    same(q, MO.encode_int16(big))
    assert np.array_equal(M.decode_int16(q), MO.decode_int16(q), equal_nan=True)
    u2, v2 = M.rotate_u_v(vx, -vy, 0.3)
    ou, ovv = MO.rotate_u_v(vx, -vy, 0.3)
    assert u2.dtype == np.float64 and np.array_equal(u2, ou, equal_nan=True) and np.array_equal(v2, ovv, equal_nan=True)


@gpu
def test_gpu_accessor_matches_reference_wrapper_semantics():
    from pyorc_b200.mask import Masks, pack_dataset

    vx, vy, c, s = fields(T=12)
    ds = make_ds(vx, vy, c, s)
    mk = Masks(ds)
    m = mk.minmax(s_min=0.3, s_max=0.9)
    assert m.dims == ("time", "y", "x")
    same(m.values, MO.minmax(vx, vy, 0.3, 0.9))
    # reduce_time: the mask of the time-mean fields, [y, x]
    mr = mk.minmax(reduce_time=True, s_min=0.3, s_max=0.9)
    mean = [MO.time_stats(a)[0] for a in (vx, vy)]
    assert mr.dims == ("y", "x")
    same(mr.values, MO.minmax(mean[0], mean[1], 0.3, 0.9))
    # window masks run per time step and keep the time dimension (mask.py:75-77)
    same(mk.window_nan(wdw=1).values, MO.window_nan(vx, wdw=1))
    mc = mk.count(tolerance=0.85)
    assert mc.dims == ("y", "x")
    # a list of masks -> masked copy; the original is untouched
    out = mk([m, mc])
    want = MO.apply_masks([vx, vy, c, s], [m.values, mc.values])
    for k, w in zip(("v_x", "v_y", "corr", "s2n"), want):
        assert np.array_equal(out[k].values, w, equal_nan=True)
    assert np.array_equal(ds["v_x"].values, vx, equal_nan=True)
    # inplace
    mk.corr(inplace=True, tolerance=0.5)
    assert np.array_equal(ds["s2n"].values, np.where(MO.corr(c, 0.5), s, np.nan), equal_nan=True)
    rep = Masks(ds).window_replace(wdw=1, iter=1)
    wr = MO.window_replace([ds[k].values for k in ("v_x", "v_y", "corr", "s2n")])
    assert np.array_equal(rep["v_x"].values, wr[0], equal_nan=True)
    packed = pack_dataset(ds)
    assert packed["v_x"].dtype == np.int16 and np.array_equal(packed["corr"], MO.encode_int16(ds["corr"].values))


@gpu
def test_gpu_masks_on_device_resident_results_config2_size():
    """The result grid of BASELINE configs[2] (132 x 237 windows, 60 time steps): tensors stay on the device."""
    import torch

    from pyorc_b200 import mask as M

    vx, vy, c, s = fields(T=60, ny=132, nx=237, seed=9)
    d = [torch.from_numpy(a).cuda() for a in (vx, vy, c, s)]
    m = M.outliers(d[0], d[1], 1.5)
    assert m.is_cuda and m.dtype == torch.bool
    same(m.cpu().numpy(), MO.outliers(vx, vy, 1.5))
    m2 = M.window_mean(d[0], d[1], 0.7)
    same(m2.cpu().numpy(), MO.window_mean(vx, vy, 0.7))
    out = M.apply_masks(d, [m, m2])
    want = MO.apply_masks([vx, vy, c, s], [m.cpu().numpy(), m2.cpu().numpy()])
    assert all(np.array_equal(o.cpu().numpy(), w, equal_nan=True) for o, w in zip(out, want))
    assert np.array_equal(d[0].cpu().numpy(), vx, equal_nan=True)       # inputs untouched
    q = M.encode_int16(out[0])
    same(q.cpu().numpy(), MO.encode_int16(want[0]))


# ---------------------------------------------------------------------------------------------------------------------------
# GPU: pyorc's own mask tests (tests/test_mask.py of the reference: they only check that every method RUNS, raises and warns
# as documented) repeated call for call on the `piv` fixture's counterpart - get_piv() of the projected Ngwerere frames
# (tests/conftest.py:390-398; window_size 25 from the camera configuration) - with every returned mask also checked against
# the oracle.
# ---------------------------------------------------------------------------------------------------------------------------
def _ngwerere_piv(ensemble_corr=False):
    import os

    from pyorc_b200 import _xr
    from pyorc_b200 import frames as b2frames

    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ngwerere_proj.npz"))
    fr, t, res = d["frames"], d["time_s"], float(d["resolution"])
    H, W = fr.shape[1:]
    da = _xr.DataArray(fr, ("time", "y", "x"), {"time": t, "y": np.flipud(np.linspace(res / 2, res * (H - 0.5), H)),
                                                "x": np.linspace(res / 2, res * (W - 0.5), W)})
    return b2frames.get_piv(da, window_size=25, engine="b200", resolution=res, ensemble_corr=ensemble_corr)


def _vals(piv):
    return [np.asarray(piv[k].values, np.float32) for k in ("v_x", "v_y", "corr", "s2n")]


@gpu
def test_reference_mask_suite_runs_like_pyorc():
    from pyorc_b200.mask import Masks

    piv = _ngwerere_piv()
    vx, vy, c, s = _vals(piv)
    assert vx.ndim == 3 and vx.shape[0] == 2
    # test_mask_minmax / test_mask: masks with and without time, combined
    m1 = Masks(piv).minmax(inplace=False)
    same(m1.values, MO.minmax(vx, vy))
    piv_mean = piv.mean(dim="time", keep_attrs=True)
    m2 = Masks(piv_mean).minmax(inplace=False)
    assert m2.dims == ("y", "x")
    out = Masks(piv)([m1, m2])
    want = MO.apply_masks([vx, vy, c, s], [m1.values, m2.values])
    assert all(np.array_equal(out[k].values, w, equal_nan=True) for k, w in zip(("v_x", "v_y", "corr", "s2n"), want))
    m3 = Masks(piv).angle()
    m4 = Masks(piv_mean).window_mean()
    Masks(piv)([m3, m4])
    # every method "runs" with the reference tests' arguments and equals the oracle
    same(Masks(piv).corr(tolerance=0.3).values, MO.corr(c, 0.3))
    same(Masks(piv).count().values, MO.count(vx))
    same(Masks(piv).rolling(tolerance=0.4).values, MO.rolling(vx, vy, tolerance=0.4))
    same(Masks(piv).outliers(mode="or").values, MO.outliers(vx, vy, mode="or"))
    same(Masks(piv).variance(tolerance=1.0, mode="or").values, MO.variance(vx, vy, 1.0, "or"))
    same(Masks(piv).window_mean().values, MO.window_mean(vx, vy))
    # test_mask_window_nan: first a filter that creates missings, in place, then the window filter in place
    p2 = _ngwerere_piv()
    Masks(p2).minmax(s_max=0.6, inplace=True)
    vx2 = np.asarray(p2["v_x"].values, np.float32)
    assert np.isnan(vx2).sum() > np.isnan(vx).sum()
    mw = Masks(p2).window_nan(inplace=True)
    same(mw.values, MO.window_nan(vx2))
    assert np.array_equal(np.isnan(p2["v_x"].values), ~mw.values | np.isnan(vx2))
    # test_error_no_time / test_error_single_time_step
    with pytest.raises(AssertionError, match='This mask requires dimension "time"'):
        Masks(piv_mean).variance()
    with pytest.warns(UserWarning, match="This mask requires multiple timesteps"):
        one = Masks(_ngwerere_piv(ensemble_corr=True)).count(inplace=True, tolerance=0.3)      # test_mask_count_ens_corr
    assert bool(one.values.all())

"""TEST INFRASTRUCTURE ONLY - CPU restatement of pyorc's velocimetry mask stack and result encoding.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product never does.

Each function follows pyorc/api/mask.py (pyorc @ be7d7c8) line by line, with the xarray operations the reference uses
restated in numpy on ``[time, y, x]`` float32 arrays (xarray keeps float32 for ``v_x, v_y, corr, s2n``;
ffpiv.py:325-326,418-419,469-470):

  ``da.count(dim)``, ``da.mean(dim)``, ``da.std(dim)``   skipna=True, ddof=0  -> np.nanmean / np.nanstd (float32 in,
                                                          float32 accumulators, sequential along a leading axis)
  ``da.shift(x=k)``                                       result[i] = da[i - k], NaN where i - k is outside
  ``da.rolling(time=w, center=True).max()``               min_periods = w; window of label i covers
                                                          [i - w // 2, i + w - 1 - w // 2]; NaN where it leaves the axis
  ``da.where(m)``, ``da.fillna(o)``                       np.where

PARITY PIN STATUS: PINNED on the reference's own example output for seven of the eleven masks.  xarray is not installed in this
image and pyorc's mask tests (tests/test_mask.py) assert no values, but the reference SHIPS the input and the output of its mask
stack: examples/ngwerere/ngwerere_piv.nc and ngwerere_masked.nc (notebook 03, cells 10 + 16: corr, minmax, rolling, outliers,
variance, count, window_mean(wdw=2, tolerance=0.5, reduce_time=True), each in place, then set_encoding + to_netcdf).  Read with a
minimal HDF5 parser (tests/golden/h5min.py, make_mask_golden.py), this restatement reproduces that output EXACTLY: the same
93 824 of 486 750 values survive, not one differs, and the survivors re-encode to the file's int16 values
(tests/test_mask.py::test_oracle_reproduces_the_reference_mask_example_exactly); float64 arithmetic instead of float32 misses it in
a few dozen threshold ties, so the arithmetic type is pinned too.  Also pinned by the files: the CF attributes (scale_factor 0.01,
_FillValue -9999, int16) and decode -> encode being the identity on them.  NOT covered by that example and still checked against
the reference's source only: `angle`, `s2n`, `window_nan`, `window_replace`, `rotate_u_v`, and encoding values beyond the int16
range (the example's s2n field shows that the reference's cast WRAPS them, -5536 for 600.0; this restatement saturates).
Reference quirks are kept (the example confirms the first: `variance` removes nothing there):
``variance`` clamps the mean from BELOW with 1e30 (mask.py:273-274, `np.maximum`), and ``helpers.stack_window`` leaves
out the last positive y stride (helpers.py:676, `range(wdw_y_min, wdw_y_max)`).
"""
from __future__ import annotations

import warnings

import numpy as np

F32 = np.float32


def _quiet():
    ctx = warnings.catch_warnings()
    ctx.__enter__()
    warnings.simplefilter("ignore", category=RuntimeWarning)
    return ctx


def speed(vx, vy):
    """``(v_x ** 2 + v_y ** 2) ** 0.5`` in float32 (mask.py:158, :299): square, add, sqrt - each rounded to float32."""
    vx, vy = np.asarray(vx, F32), np.asarray(vy, F32)
    with np.errstate(all="ignore"):
        return np.sqrt(vx * vx + vy * vy)


def minmax(vx, vy, s_min=0.1, s_max=5.0):
    """mask.py:147-161."""
    s = speed(vx, vy)
    with np.errstate(invalid="ignore"):
        return (s > F32(s_min)) & (s < F32(s_max))


def angle(vx, vy, angle_expected=0.5 * np.pi, angle_tolerance=0.25 * np.pi):
    """mask.py:163-186: ``arctan2(v_x, v_y)`` (clockwise from "up"), float32."""
    a = np.arctan2(np.asarray(vx, F32), np.asarray(vy, F32))
    with np.errstate(invalid="ignore"):
        return np.abs(a - F32(angle_expected)) < F32(angle_tolerance)


def count(vx, tolerance=0.33):
    """mask.py:188-201: valid samples in time > tolerance * len(time) -> mask [y, x]."""
    vx = np.asarray(vx)
    return (~np.isnan(vx)).sum(axis=0) > tolerance * vx.shape[0]


def corr(c, tolerance=0.1):
    """mask.py:203-213."""
    with np.errstate(invalid="ignore"):
        return np.asarray(c, F32) > F32(tolerance)


def s2n(s, tolerance=10):
    """mask.py:215-225."""
    with np.errstate(invalid="ignore"):
        return np.asarray(s, F32) > F32(tolerance)


def time_stats(v):
    """``mean(dim="time")`` and ``std(dim="time")`` with skipna (float32)."""
    v = np.asarray(v, F32)
    ctx = _quiet()
    try:
        with np.errstate(all="ignore"):
            return np.nanmean(v, axis=0), np.nanstd(v, axis=0)
    finally:
        ctx.__exit__(None, None, None)


def outliers(vx, vy, tolerance=1.0, mode="or"):
    """mask.py:227-252: |(v - mean_t) / std_t| < tolerance per component -> mask [time, y, x]."""
    vx, vy = np.asarray(vx, F32), np.asarray(vy, F32)
    xm, xs = time_stats(vx)
    ym, ys = time_stats(vy)
    with np.errstate(all="ignore"):
        xc = np.abs((vx - xm) / xs) < F32(tolerance)
        yc = np.abs((vy - ym) / ys) < F32(tolerance)
    return (xc | yc) if mode == "or" else (xc & yc)


def variance(vx, vy, tolerance=5, mode="and"):
    """mask.py:254-285 (including ``np.maximum(mean, 1e30)``) -> mask [y, x]."""
    xm, xs = time_stats(vx)
    ym, ys = time_stats(vy)
    with np.errstate(all="ignore"):
        xm = np.maximum(xm, F32(1e30))
        ym = np.maximum(ym, F32(1e30))
        xc = np.abs(xs / xm) < F32(tolerance)
        yc = np.abs(ys / ym) < F32(tolerance)
    return (xc | yc) if mode == "or" else (xc & yc)


def rolling_max_centered(s, wdw):
    """``s.rolling(time=wdw, center=True).max()`` with the default min_periods (= wdw)."""
    s = np.asarray(s, F32)
    n = s.shape[0]
    out = np.full(s.shape, np.nan, F32)
    lo, hi = wdw // 2, wdw - 1 - wdw // 2
    for i in range(n):
        a, b = i - lo, i + hi
        if a < 0 or b >= n:
            continue
        out[i] = s[a:b + 1].max(axis=0)
    return out


def rolling(vx, vy, wdw=5, tolerance=0.5):
    """mask.py:287-303."""
    s = speed(vx, vy)
    s_roll = rolling_max_centered(np.where(np.isnan(s), F32(0), s), wdw)
    with np.errstate(invalid="ignore"):
        return s > F32(tolerance) * s_roll


def strides(wdw=1, wdw_x_min=None, wdw_x_max=None, wdw_y_min=None, wdw_y_max=None):
    """The (x_stride, y_stride) list of helpers.stack_window (helpers.py:667-677), in its order."""
    wdw_x_min = -wdw if wdw_x_min is None else wdw_x_min
    wdw_x_max = wdw if wdw_x_max is None else wdw_x_max
    wdw_y_min = -wdw if wdw_y_min is None else wdw_y_min
    wdw_y_max = wdw if wdw_y_max is None else wdw_y_max
    return [(xs, ys) for xs in range(wdw_x_min, wdw_x_max + 1) for ys in range(wdw_y_min, wdw_y_max)]


def shift_yx(a, xs, ys):
    """``a.shift(x=xs, y=ys)`` on the last two axes (y, x): result[.., i, j] = a[.., i - ys, j - xs], NaN outside."""
    a = np.asarray(a, F32)
    out = np.full(a.shape, np.nan, F32)
    ny, nx = a.shape[-2:]
    y0, y1 = max(0, ys), min(ny, ny + ys)
    x0, x1 = max(0, xs), min(nx, nx + xs)
    if y0 < y1 and x0 < x1:
        out[..., y0:y1, x0:x1] = a[..., y0 - ys:y1 - ys, x0 - xs:x1 - xs]
    return out


def stack_window(a, **kw):
    """helpers.py:638-679 for one variable: [stride, ..., y, x]."""
    return np.stack([shift_yx(a, xs, ys) for xs, ys in strides(**kw)])


def window_nan(vx, tolerance=0.7, wdw=1, **kw):
    """mask.py:305-337 (applied per time step by the wrapper, mask.py:75-77)."""
    st = stack_window(vx, wdw=wdw, **kw)
    return (~np.isnan(st)).sum(axis=0) >= tolerance * st.shape[0]


def window_mean_of(a, wdw=1, **kw):
    """``stack_window(...).mean(dim="stride")``: float32 sum in stride order over the valid samples / their count."""
    st = stack_window(a, wdw=wdw, **kw)
    ctx = _quiet()
    try:
        with np.errstate(all="ignore"):
            return np.nanmean(st, axis=0)
    finally:
        ctx.__exit__(None, None, None)


def window_mean(vx, vy, tolerance=0.7, wdw=1, mode="or", **kw):
    """mask.py:339-377."""
    vx, vy = np.asarray(vx, F32), np.asarray(vy, F32)
    mx, my = window_mean_of(vx, wdw=wdw, **kw), window_mean_of(vy, wdw=wdw, **kw)
    with np.errstate(all="ignore"):
        xc = np.abs(vx - mx) / mx < F32(tolerance)
        yc = np.abs(vy - my) / my < F32(tolerance)
    return (xc | yc) if mode == "or" else (xc & yc)


def window_replace(fields, wdw=1, iter=1, **kw):  # noqa: A002  (reference argument name)
    """mask.py:379-403: NaNs of every variable filled with the window mean, ``iter`` times."""
    out = [np.array(f, F32, copy=True) for f in fields]
    for _ in range(iter):
        means = [window_mean_of(f, wdw=wdw, **kw) for f in out]
        out = [np.where(np.isnan(f), m, f) for f, m in zip(out, means)]
    return out


def apply_masks(fields, masks):
    """``ds[var].where(m)`` for every variable and mask (mask.py:131-144); masks broadcast over time."""
    out = [np.array(f, F32, copy=True) for f in fields]
    for m in masks:
        m = np.asarray(m, bool)
        out = [np.where(m, f, F32(np.nan)) for f in out]
    return out


# ---- f-4: CF packing pyorc sets for v_x, v_y, corr, s2n (const.py:80-83; api/velocimetry.py:239-253) ---------------
SCALE, FILL = 0.01, -9999


def encode_int16(a, scale_factor=SCALE, fill_value=FILL):
    """What xarray's CF encoders write for ``{"dtype": "int16", "scale_factor": 0.01, "_FillValue": -9999}``: float32
    ``a / scale_factor``, NaN -> _FillValue, round half to even, cast.  Values beyond the int16 range saturate here (the
    reference's cast is undefined for them)."""
    a = np.asarray(a, F32)
    with np.errstate(all="ignore"):
        q = np.rint(a / F32(scale_factor))
    q = np.where(np.isnan(q), F32(fill_value), np.clip(q, -32768, 32767))
    return q.astype(np.int16)


def decode_int16(q, scale_factor=SCALE, fill_value=FILL):
    """The inverse xarray applies on reading: _FillValue -> NaN, ``q * scale_factor`` in float32."""
    q = np.asarray(q, np.int16)
    return np.where(q == fill_value, F32(np.nan), q.astype(F32) * F32(scale_factor)).astype(F32)


def rotate_u_v(u, v, theta):
    """helpers.py:602-630 with a float64 angle on float32 fields: the products promote to float64 (r is a float64 array
    element, a numpy scalar, not a Python float)."""
    c, s = np.cos(theta), np.sin(theta)
    r = np.array(((c, -s), (s, c)))
    u, v = np.asarray(u), np.asarray(v)
    return r[0, 0] * u + r[0, 1] * v, r[1, 0] * u + r[1, 1] * v

"""Velocimetry masks and result packing on the GPU - counterparts of pyorc's ``ds.velocimetry.mask.*`` methods
(pyorc/api/mask.py:147-403), ``Velocimetry.set_encoding`` (pyorc/api/velocimetry.py:239-253, pyorc/const.py:80-83) and
``helpers.rotate_u_v`` (pyorc/helpers.py:602-630), so that the fields the PIV engine leaves in HBM can be filtered and
packed without the eleven xarray passes of the reference.

Two layers, same names and arguments as the reference:

* functions on arrays - ``minmax(v_x, v_y, s_min, s_max)`` ... : torch CUDA tensors in -> torch bool tensors out (stream
  ordered, nothing leaves the device); numpy in -> numpy out (copied through the device);
* :class:`Masks` - the accessor: ``Masks(ds).minmax(inplace=False, reduce_time=False, s_min=0.1, s_max=5.0)`` on the
  Dataset ``get_piv`` returns (xarray when installed, the ``_xr`` stand-in otherwise), with the reference wrapper's
  semantics (``inplace``, ``reduce_time``, time requirements, the single-time-step warning; mask.py:23-90).

No CPU fallback: every function calls the CUDA library through its C ABI (include/b2piv.h).
"""

from __future__ import annotations

import copy
import ctypes
import warnings

import numpy as np

from .engine import Engine, get_engine

V_X, V_Y, CORR, S2N = "v_x", "v_y", "corr", "s2n"          # pyorc/const.py:6-9
ENCODING_PARAMS = {"zlib": True, "dtype": "int16", "scale_factor": 0.01, "_FillValue": -9999}   # pyorc/const.py:80
ENCODE_VARS = [V_X, V_Y, CORR, S2N]                                                               # pyorc/const.py:82
OP_MINMAX, OP_ANGLE, OP_THRESHOLD = 0, 1, 2

__all__ = ["Masks", "minmax", "angle", "count", "corr", "s2n", "outliers", "variance", "rolling", "window_nan", "window_mean",
           "window_replace", "apply_masks", "time_mean", "encode_int16", "decode_int16", "rotate_u_v", "set_encoding"]


# ---- plumbing --------------------------------------------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _dev(a, eng: Engine, dtype=None):
    """-> (contiguous CUDA tensor, came_from_numpy)."""
    import torch

    dtype = dtype or torch.float32
    if _is_torch(a):
        if not a.is_cuda:
            raise TypeError("torch fields must be CUDA tensors (pass numpy for host data)")
        return a.to(dtype).contiguous(), False
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype={torch.float32: np.float32, torch.int16: np.int16,
                                                                         torch.uint8: np.uint8}[dtype])).to(f"cuda:{eng.device}"), True


def _stream(t):
    import torch

    return torch.cuda.current_stream(t.device).cuda_stream


def _mask_out(t, shape):
    import torch

    # torch.bool is one byte holding 0 / 1 - exactly what the kernels write, so no conversion pass afterwards
    return torch.empty(shape, dtype=torch.bool, device=t.device)


def _back(m, was_np, as_bool=True):
    return m.cpu().numpy() if was_np else m


def _tyx(t):
    """[time, y, x] view of a [y, x] or [time, y, x] tensor."""
    if t.dim() == 2:
        return t.unsqueeze(0), False
    if t.dim() != 3:
        raise ValueError("fields must be [time, y, x] or [y, x]")
    return t, True


def _strides(wdw=1, wdw_x_min=None, wdw_x_max=None, wdw_y_min=None, wdw_y_max=None):
    """helpers.stack_window's defaults (pyorc/helpers.py:667-670) -> the four ints of the C ABI."""
    vals = [-wdw if wdw_x_min is None else wdw_x_min, wdw if wdw_x_max is None else wdw_x_max,
            -wdw if wdw_y_min is None else wdw_y_min, wdw if wdw_y_max is None else wdw_y_max]
    return (ctypes.c_int * 4)(*[int(v) for v in vals])


def _ptrs(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


# ---- functions on arrays ---------------------------------------------------------------------------------------------------
def _elementwise(op, a, b, p0, p1, device):
    eng = get_engine(device)
    ta, was_np = _dev(a, eng)
    tb = _dev(b, eng)[0] if b is not None else ta
    if tb.shape != ta.shape:
        raise ValueError("fields must have the same shape")
    m = _mask_out(ta, ta.shape)
    eng._check(eng._lib.b2piv_mask_elementwise(eng._h, op, ta.data_ptr(), tb.data_ptr(), ta.numel(), float(p0), float(p1), m.data_ptr(),
                                               _stream(ta)), "b2piv_mask_elementwise")
    return _back(m, was_np)


def minmax(v_x, v_y, s_min=0.1, s_max=5.0, device: int = 0):
    """Keep where ``s_min < sqrt(v_x^2 + v_y^2) < s_max`` (mask.py:147-161)."""
    return _elementwise(OP_MINMAX, v_x, v_y, np.float32(s_min), np.float32(s_max), device)


def angle(v_x, v_y, angle_expected=0.5 * np.pi, angle_tolerance=0.25 * np.pi, device: int = 0):
    """Keep where ``|arctan2(v_x, v_y) - angle_expected| < angle_tolerance`` (mask.py:163-186)."""
    return _elementwise(OP_ANGLE, v_x, v_y, np.float32(angle_expected), np.float32(angle_tolerance), device)


def corr(c, tolerance=0.1, device: int = 0):
    """Keep where the correlation exceeds ``tolerance`` (mask.py:203-213)."""
    return _elementwise(OP_THRESHOLD, c, None, np.float32(tolerance), 0.0, device)


def s2n(s, tolerance=10, device: int = 0):
    """Keep where the signal-to-noise ratio exceeds ``tolerance`` (mask.py:215-225)."""
    return _elementwise(OP_THRESHOLD, s, None, np.float32(tolerance), 0.0, device)


def _time_fields(eng, *fields):
    ts, was_np = [], False
    for f in fields:
        t, w = _dev(f, eng)
        was_np = was_np or w
        if t.dim() != 3:
            raise ValueError('this mask requires dimension "time": fields must be [time, y, x]')
        ts.append(t)
    if any(t.shape != ts[0].shape for t in ts):
        raise ValueError("fields must have the same shape")
    return ts, was_np


def time_stats(field, device: int = 0):
    """(count, mean, std) over time with NaNs skipped, ddof = 0 - ``da.count / mean / std(dim="time")``."""
    import torch

    eng = get_engine(device)
    (t,), was_np = _time_fields(eng, field)
    T, ny, nx = t.shape
    cnt = torch.empty((ny, nx), dtype=torch.int32, device=t.device)
    mean = torch.empty((ny, nx), dtype=torch.float32, device=t.device)
    std = torch.empty_like(mean)
    eng._check(eng._lib.b2piv_time_stats(eng._h, t.data_ptr(), T, ny * nx, cnt.data_ptr(), mean.data_ptr(), std.data_ptr(), _stream(t)),
               "b2piv_time_stats")
    return tuple(_back(x, was_np, as_bool=False) for x in (cnt, mean, std))


def time_mean(field, device: int = 0):
    """``da.mean(dim="time")`` (the wrapper's ``reduce_time``, mask.py:49-50)."""
    return time_stats(field, device)[1]


def count(v_x, tolerance=0.33, device: int = 0):
    """Keep locations with more than ``tolerance * len(time)`` valid velocities -> [y, x] (mask.py:188-201)."""
    eng = get_engine(device)
    (t,), was_np = _time_fields(eng, v_x)
    T, ny, nx = t.shape
    m = _mask_out(t, (ny, nx))
    eng._check(eng._lib.b2piv_mask_count(eng._h, t.data_ptr(), T, ny * nx, float(tolerance), m.data_ptr(), _stream(t)), "b2piv_mask_count")
    return _back(m, was_np)


def _mode(mode):
    if mode not in ("or", "and"):
        raise ValueError('mode must be "or" or "and"')
    return 1 if mode == "and" else 0


def outliers(v_x, v_y, tolerance=1.0, mode="or", device: int = 0):
    """Keep where ``|(v - mean_t) / std_t| < tolerance`` for one ("or") / both ("and") components (mask.py:227-252)."""
    eng = get_engine(device)
    (tx, ty), was_np = _time_fields(eng, v_x, v_y)
    T, ny, nx = tx.shape
    m = _mask_out(tx, tx.shape)
    eng._check(eng._lib.b2piv_mask_outliers(eng._h, tx.data_ptr(), ty.data_ptr(), T, ny * nx, float(tolerance), _mode(mode), m.data_ptr(),
                                            _stream(tx)), "b2piv_mask_outliers")
    return _back(m, was_np)


def variance(v_x, v_y, tolerance=5, mode="and", device: int = 0):
    """Keep locations where ``|std_t / max(mean_t, 1e30)| < tolerance`` -> [y, x] (mask.py:254-285, its clamp included)."""
    eng = get_engine(device)
    (tx, ty), was_np = _time_fields(eng, v_x, v_y)
    T, ny, nx = tx.shape
    m = _mask_out(tx, (ny, nx))
    eng._check(eng._lib.b2piv_mask_variance(eng._h, tx.data_ptr(), ty.data_ptr(), T, ny * nx, float(tolerance), _mode(mode), m.data_ptr(),
                                            _stream(tx)), "b2piv_mask_variance")
    return _back(m, was_np)


def rolling(v_x, v_y, wdw=5, tolerance=0.5, device: int = 0):
    """Keep where the speed exceeds ``tolerance`` times the maximum speed of the centred window of ``wdw`` time steps
    (NaN counts as 0; the first ``wdw // 2`` and last ``wdw - 1 - wdw // 2`` steps have no full window -> False;
    mask.py:287-303)."""
    eng = get_engine(device)
    (tx, ty), was_np = _time_fields(eng, v_x, v_y)
    T, ny, nx = tx.shape
    m = _mask_out(tx, tx.shape)
    eng._check(eng._lib.b2piv_mask_rolling(eng._h, tx.data_ptr(), ty.data_ptr(), T, ny * nx, int(wdw), float(tolerance), m.data_ptr(),
                                           _stream(tx)), "b2piv_mask_rolling")
    return _back(m, was_np)


def window_nan(v_x, tolerance=0.7, wdw=1, device: int = 0, **kwargs):
    """Keep where at least ``tolerance`` of the neighbourhood (helpers.stack_window) is valid (mask.py:305-337)."""
    eng = get_engine(device)
    t, was_np = _dev(v_x, eng)
    t3, had_time = _tyx(t)
    T, ny, nx = t3.shape
    m = _mask_out(t, t3.shape)
    eng._check(eng._lib.b2piv_mask_window_nan(eng._h, t3.data_ptr(), T, ny, nx, _strides(wdw, **kwargs), float(tolerance), m.data_ptr(),
                                              _stream(t)), "b2piv_mask_window_nan")
    return _back(m if had_time else m[0], was_np)


def window_mean(v_x, v_y, tolerance=0.7, wdw=1, mode="or", device: int = 0, **kwargs):
    """Keep where ``|v - mean| / mean < tolerance`` against the neighbourhood mean (mask.py:339-377)."""
    eng = get_engine(device)
    tx, was_np = _dev(v_x, eng)
    ty = _dev(v_y, eng)[0]
    if tx.shape != ty.shape:
        raise ValueError("fields must have the same shape")
    t3, had_time = _tyx(tx)
    T, ny, nx = t3.shape
    m = _mask_out(tx, t3.shape)
    eng._check(eng._lib.b2piv_mask_window_mean(eng._h, tx.data_ptr(), ty.data_ptr(), T, ny, nx, _strides(wdw, **kwargs), float(tolerance),
                                               _mode(mode), m.data_ptr(), _stream(tx)), "b2piv_mask_window_mean")
    return _back(m if had_time else m[0], was_np)


def window_replace(fields, wdw=1, iter=1, device: int = 0, **kwargs):  # noqa: A002  (reference argument name)
    """NaNs of every field replaced by the mean of their neighbourhood, ``iter`` times (mask.py:379-403).  Returns new
    arrays (the reference deep-copies the Dataset)."""
    eng = get_engine(device)
    pairs = [_dev(f, eng) for f in fields]
    ts = [t.clone() if not w else t for t, w in pairs]
    was_np = any(w for _, w in pairs)
    if not 1 <= len(ts) <= 4 or any(t.shape != ts[0].shape for t in ts):
        raise ValueError("1..4 fields of the same shape")
    t3, _ = _tyx(ts[0])
    T, ny, nx = t3.shape
    eng._check(eng._lib.b2piv_window_replace(eng._h, _ptrs(ts), len(ts), T, ny, nx, _strides(wdw, **kwargs), int(iter), _stream(ts[0])),
               "b2piv_window_replace")
    return [_back(t, was_np, as_bool=False) for t in ts]


def apply_masks(fields, masks, device: int = 0):
    """``field.where(mask)`` for every field and every mask, on copies (mask.py:131-144).  Masks are [time, y, x] or
    [y, x] (broadcast over time)."""
    import torch

    eng = get_engine(device)
    pairs = [_dev(f, eng) for f in fields]
    ts = [t.clone() if not w else t for t, w in pairs]
    was_np = any(w for _, w in pairs)
    if not 1 <= len(ts) <= 4 or any(t.shape != ts[0].shape for t in ts):
        raise ValueError("1..4 fields of the same shape")
    t3, _ = _tyx(ts[0])
    T, ny, nx = t3.shape
    if not isinstance(masks, (list, tuple)):
        masks = [masks]
    for m in masks:
        if _is_torch(m):
            tm = m.to(ts[0].device).contiguous()
            tm = tm.view(torch.uint8) if tm.dtype == torch.bool else (tm != 0).view(torch.uint8)
        else:
            tm = torch.from_numpy(np.ascontiguousarray(np.asarray(m), dtype=np.uint8)).to(ts[0].device)
        if tuple(tm.shape) == (T, ny, nx) and ts[0].dim() == 3:
            has_time = 1
        elif tuple(tm.shape) == (ny, nx):
            has_time = 0
        else:
            raise ValueError(f"mask shape {tuple(tm.shape)} does not match fields {tuple(ts[0].shape)}")
        eng._check(eng._lib.b2piv_mask_apply(eng._h, _ptrs(ts), len(ts), T, ny * nx, tm.data_ptr(), has_time, _stream(ts[0])),
                   "b2piv_mask_apply")
    return [_back(t, was_np, as_bool=False) for t in ts]


# ---- f-4: result packing ---------------------------------------------------------------------------------------------------
def encode_int16(field, scale_factor=ENCODING_PARAMS["scale_factor"], fill_value=ENCODING_PARAMS["_FillValue"], device: int = 0):
    """int16 packing xarray writes for pyorc's encoding (const.py:80): ``round(field / scale_factor)``, NaN -> fill."""
    import torch

    eng = get_engine(device)
    t, was_np = _dev(field, eng)
    q = torch.empty(t.shape, dtype=torch.int16, device=t.device)
    eng._check(eng._lib.b2piv_encode_int16(eng._h, t.data_ptr(), t.numel(), float(np.float32(scale_factor)), int(fill_value), q.data_ptr(),
                                           _stream(t)), "b2piv_encode_int16")
    return _back(q, was_np, as_bool=False)


def decode_int16(packed, scale_factor=ENCODING_PARAMS["scale_factor"], fill_value=ENCODING_PARAMS["_FillValue"], device: int = 0):
    """Inverse of :func:`encode_int16`: fill -> NaN, ``packed * scale_factor`` in float32."""
    import torch

    eng = get_engine(device)
    q, was_np = _dev(packed, eng, dtype=torch.int16)
    out = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    eng._check(eng._lib.b2piv_decode_int16(eng._h, q.data_ptr(), q.numel(), float(np.float32(scale_factor)), int(fill_value),
                                           out.data_ptr(), _stream(q)), "b2piv_decode_int16")
    return _back(out, was_np, as_bool=False)


def rotate_u_v(u, v, theta, deg=False, device: int = 0):
    """Rotate vectors counter-clockwise by ``theta`` (pyorc/helpers.py:602-630); float64 results like numpy's promotion."""
    import torch

    eng = get_engine(device)
    theta = float(np.radians(theta)) if deg else float(theta)
    tu, was_np = _dev(u, eng)
    tv = _dev(v, eng)[0]
    if tu.shape != tv.shape:
        raise ValueError("fields must have the same shape")
    u2 = torch.empty(tu.shape, dtype=torch.float64, device=tu.device)
    v2 = torch.empty_like(u2)
    eng._check(eng._lib.b2piv_rotate_uv(eng._h, tu.data_ptr(), tv.data_ptr(), tu.numel(), theta, u2.data_ptr(), v2.data_ptr(),
                                        _stream(tu)), "b2piv_rotate_uv")
    return _back(u2, was_np, as_bool=False), _back(v2, was_np, as_bool=False)


def set_encoding(ds, enc_pars=None):
    """``Velocimetry.set_encoding`` (pyorc/api/velocimetry.py:239-253): attach the packing parameters to the variables."""
    enc_pars = ENCODING_PARAMS if enc_pars is None else enc_pars
    for k in ENCODE_VARS:
        ds[k].encoding = enc_pars
    return ds


def pack_dataset(ds, enc_pars=None, device: int = 0):
    """The four variables as the int16 arrays ``to_netcdf`` would store with pyorc's encoding: {name: int16 ndarray}."""
    enc_pars = ENCODING_PARAMS if enc_pars is None else enc_pars
    return {k: encode_int16(np.asarray(ds[k].values, np.float32), enc_pars["scale_factor"], enc_pars["_FillValue"], device)
            for k in ENCODE_VARS if k in ds}


# ---- the accessor ----------------------------------------------------------------------------------------------------------
def _dims(da):
    return tuple(da.dims)


def _new_da(template, data, dims):
    """DataArray of the same family as `template` (xarray or the stand-in) on the template's coordinates."""
    coords = {}
    for d in dims:
        try:
            coords[d] = np.asarray(template[d].values if hasattr(template[d], "values") else template[d])
        except Exception:
            pass
    return type(template)(data, dims=dims, coords=coords)


def _set_values(ds, var, data):
    da = ds[var]
    if type(da).__module__.startswith("xarray"):
        ds[var] = (da.dims, data, da.attrs)
    else:
        da.values = data


def _base_mask(time_allowed=False, time_required=False, multi_timestep_required=False):
    """The reference's wrapper (mask.py:23-90): ``inplace`` / ``reduce_time`` handling and the time-dimension rules."""

    def decorator_func(mask_func):
        def wrapper_func(ref, inplace=False, reduce_time=False, *args, **kwargs):
            ds = ref._fields(reduce_time)
            has_time = ds["dims"][0] == "time"
            single = False
            if time_required:
                if not has_time:
                    raise AssertionError(
                        'This mask requires dimension "time". The dataset does not contain dimension "time" or you '
                        "have set `reduce_time=True`. Apply this mask without applying any reducers in time."
                    )
                if multi_timestep_required and ds[V_X].shape[0] < 2:
                    warnings.warn(
                        "This mask requires multiple timesteps in the dataset in order have an effect. This "
                        "warning typically occurs when applying `Frames.get_piv(ensemble_corr=True)` as this only "
                        "yields one single time step.",
                        stacklevel=2,
                    )
                    single = True
            if single:
                data, dims = np.ones(ds[V_X].shape[-2:], bool), ("y", "x")      # "just pass Trues everywhere" (mask.py:78-80)
            else:
                data = mask_func(ref, ds, *args, **kwargs)
                dims = ds["dims"] if np.ndim(data) == len(ds["dims"]) else ds["dims"][-2:]
            mask = _new_da(ref._obj[V_X], data, dims)
            if inplace:
                ref._where(data)
            return mask

        wrapper_func.__name__ = mask_func.__name__
        wrapper_func.__doc__ = mask_func.__doc__
        return wrapper_func

    return decorator_func


class Masks:
    """``ds.velocimetry.mask`` on the GPU: the methods of pyorc's ``_Velocimetry_MaskMethods`` (mask.py:93-403).

    ``Masks(ds).minmax(s_min=0.2)`` returns the mask (a bool DataArray); with ``inplace=True`` the Dataset's variables are
    masked as well; ``Masks(ds)([m1, m2])`` returns a masked copy (``inplace=True``: masks ``ds`` itself).
    """

    def __init__(self, ds, device: int = 0):
        for k in ENCODE_VARS:
            if k not in ds:
                raise AssertionError("Dataset is not a valid velocimetry dataset")
        self._obj = ds
        self.device = device

    # -- plumbing ----------------------------------------------------------------------------------------------
    def _fields(self, reduce_time):
        """The four variables as float32 numpy arrays (+ dims); ``reduce_time``: their mean over time (mask.py:49-50)."""
        dims = _dims(self._obj[V_X])
        out = {k: np.ascontiguousarray(self._obj[k].values, dtype=np.float32) for k in ENCODE_VARS}
        if reduce_time and "time" in dims:
            if dims[0] != "time":
                raise ValueError('"time" must be the leading dimension')
            out = {k: time_mean(v, self.device) for k, v in out.items()}
            dims = dims[1:]
        out["dims"] = dims
        return out

    def _where(self, mask, ds=None):
        ds = self._obj if ds is None else ds
        vals = [np.ascontiguousarray(ds[k].values, dtype=np.float32) for k in ENCODE_VARS]
        m = np.asarray(mask.values if hasattr(mask, "values") else mask)
        for k, v in zip(ENCODE_VARS, apply_masks(vals, [m], self.device)):
            _set_values(ds, k, v)

    def __call__(self, mask, inplace=False, *args, **kwargs):
        """Apply one mask or a list of masks (mask.py:110-144)."""
        if not isinstance(mask, list):
            mask = [mask]
        ds = self._obj if inplace else copy.deepcopy(self._obj)
        for m in mask:
            self._where(m, ds)
        if not inplace:
            return ds

    # -- the masks (defaults and argument names of the reference) ----------------------------------------------------------
    @_base_mask(time_allowed=True)
    def minmax(self, ds, s_min=0.1, s_max=5.0):
        """Masks values if the velocity scalar lies outside a user-defined valid range (mask.py:147-161)."""
        return minmax(ds[V_X], ds[V_Y], s_min, s_max, self.device)

    @_base_mask(time_allowed=True)
    def angle(self, ds, angle_expected=0.5 * np.pi, angle_tolerance=0.25 * np.pi):
        """Mask values that are outside expected direction with angle tolerance (mask.py:163-186)."""
        return angle(ds[V_X], ds[V_Y], angle_expected, angle_tolerance, self.device)

    @_base_mask(time_required=True, multi_timestep_required=True)
    def count(self, ds, tolerance=0.33):
        """Mask locations with a too low amount of valid velocities in time (mask.py:188-201)."""
        return count(ds[V_X], tolerance, self.device)

    @_base_mask(time_allowed=True)
    def corr(self, ds, tolerance=0.1):
        """Mask values with too low correlation (mask.py:203-213)."""
        return corr(ds[CORR], tolerance, self.device)

    @_base_mask(time_allowed=True)
    def s2n(self, ds, tolerance=10):
        """Mask values with too low signal to noise (mask.py:215-225)."""
        return s2n(ds[S2N], tolerance, self.device)

    @_base_mask(time_required=True, multi_timestep_required=True)
    def outliers(self, ds, tolerance=1.0, mode="or"):
        """Mask outliers measured by amount of standard deviations from the mean (mask.py:227-252)."""
        return outliers(ds[V_X], ds[V_Y], tolerance, mode, self.device)

    @_base_mask(time_required=True, multi_timestep_required=True)
    def variance(self, ds, tolerance=5, mode="and"):
        """Mask locations if their variance (std/mean in time) is above a tolerance level (mask.py:254-285)."""
        return variance(ds[V_X], ds[V_Y], tolerance, mode, self.device)

    @_base_mask(time_required=True, multi_timestep_required=True)
    def rolling(self, ds, wdw=5, tolerance=0.5):
        """Mask values for strongly deviating values from neighbours over rolling length (mask.py:287-303)."""
        return rolling(ds[V_X], ds[V_Y], wdw, tolerance, self.device)

    @_base_mask()
    def window_nan(self, ds, tolerance=0.7, wdw=1, **kwargs):
        """Masks values if their surrounding neighbours (inc. value itself) contain too many NaNs (mask.py:305-337)."""
        return window_nan(ds[V_X], tolerance, wdw, self.device, **kwargs)

    @_base_mask()
    def window_mean(self, ds, tolerance=0.7, wdw=1, mode="or", **kwargs):
        """Mask values when their value deviates significantly from the mean of their neighbours (mask.py:339-377)."""
        return window_mean(ds[V_X], ds[V_Y], tolerance, wdw, mode, self.device, **kwargs)

    def window_replace(self, inplace=False, reduce_time=False, wdw=1, iter=1, **kwargs):  # noqa: A002
        """Replace NaNs with the mean of their neighbours; returns a Dataset instead of a mask (mask.py:379-403)."""
        f = self._fields(reduce_time)
        new = window_replace([f[k] for k in ENCODE_VARS], wdw, iter, self.device, **kwargs)
        ds = copy.deepcopy(self._obj)
        if reduce_time and "time" in _dims(self._obj[V_X]):
            ds = ds.mean(dim="time", keep_attrs=True)
        for k, v in zip(ENCODE_VARS, new):
            _set_values(ds, k, v)
        return ds

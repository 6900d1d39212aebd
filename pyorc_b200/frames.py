"""``get_piv`` - B200 counterpart of ``pyorc.api.frames.Frames.get_piv`` (pyorc/api/frames.py:114-197).

Everything above the engine call (window rounding, default overlap, result coordinates, engine whitelist and its
``ValueError``) follows the reference; the engine call goes to :func:`pyorc_b200.velocimetry.get_b2piv`.
``install()`` registers ``engine="b200"`` inside a real pyorc installation.
"""

from __future__ import annotations

import contextvars
from typing import Optional

import numpy as np

from . import window
from .velocimetry import get_b2piv

__all__ = ["get_piv", "get_piv_coords", "install", "uninstall"]

ENGINES = ["b200"]
ENCODING_PARAMS = {"zlib": True, "dtype": "int16", "scale_factor": 0.01, "_FillValue": -9999}   # pyorc/const.py:80


def get_piv_coords(frames, window_size, search_area_size, overlap):
    """Row/column centres -> local ``y``/``x`` axes of the result (frames.py:47-112, helpers.get_axes :142-168).
    Only the two 1-D axes the engine binding needs are derived here; the 2-D geographic grids stay pyorc's."""
    dim_size = frames[0].shape
    cols_vector, rows_vector = window.get_rect_coordinates(
        dim_size=dim_size, window_size=window_size, search_area_size=search_area_size, overlap=overlap
    )
    coords_in = getattr(frames, "coords", {})
    fx = np.asarray(coords_in["x"]) if "x" in coords_in else np.arange(dim_size[1], dtype=np.float64)
    fy = np.asarray(coords_in["y"]) if "y" in coords_in else np.arange(dim_size[0], dtype=np.float64)
    return {"y": fy[rows_vector], "x": fx[cols_vector]}, {"cols": cols_vector, "rows": rows_vector}


def get_piv(
    frames,
    window_size=None,
    overlap=None,
    engine: str = "b200",
    ensemble_corr: bool = False,
    resolution: Optional[float] = None,
    camera_config=None,
    **kwargs,
):
    """Perform PIV on projected frames ``[time, y, x]`` with the B200 engine.

    ``window_size`` / ``resolution`` default to the camera configuration's (``camera_config.window_size``,
    ``camera_config.resolution``) like the reference (frames.py:156-185); ``overlap`` defaults to
    ``int(round(window_size) / 2)`` of the un-rounded size (frames.py:170-171).
    """
    if camera_config is None:
        camera_config = getattr(frames, "camera_config", None)
    cfg_ws = window_size if window_size is not None else getattr(camera_config, "window_size", None)
    if cfg_ws is None:
        raise ValueError("window_size must be given (directly or through camera_config)")
    if resolution is None:
        resolution = getattr(camera_config, "resolution", None)
    if resolution is None:
        raise ValueError("resolution must be given (directly or through camera_config)")
    ws = 2 * (cfg_ws,) if isinstance(cfg_ws, (int, np.integer)) else tuple(cfg_ws)
    ws = window.round_to_even(ws)
    search_area_size = ws
    if overlap is None:
        overlap = 2 * (int(round(cfg_ws) / 2),)  # raises for tuple sizes exactly like the reference
    if engine not in ENGINES:
        raise ValueError(f"Selected PIV engine {engine} does not exist.")
    coords, _ = get_piv_coords(frames, ws, search_area_size, overlap)
    n = len(frames)
    coords_in = getattr(frames, "coords", {})
    if "time" in coords_in:
        t = np.asarray(coords_in["time"], dtype=np.float64)
        dt = np.diff(t)
    else:
        dt = kwargs.pop("dt", None)
        if dt is None:
            raise ValueError("frames carry no time coordinate: pass dt=")
        dt = np.broadcast_to(np.asarray(dt, dtype=np.float64), (n - 1,)).copy()
    kwargs = {
        **kwargs,
        "search_area_size": search_area_size,
        "window_size": ws,
        "overlap": tuple(overlap),
        "res_x": resolution,
        "res_y": resolution,
    }
    ds = get_b2piv(frames, coords["y"], coords["x"], dt, engine=engine, ensemble_corr=ensemble_corr, **kwargs)
    # metadata like the reference body (frames.py:190-196): the frames' attrs, the camera configuration actually used when it
    # can serialise itself, and pyorc's int16 CF encoding on the four variables (const.py:80-83).  The 2-D mesh coordinates
    # (xs, ys, lon, lat, xp, yp) need pyorc's CameraConfig and are added by the reference's own body - `install()`.
    ds.attrs = dict(getattr(frames, "attrs", {}))
    if camera_config is not None and hasattr(camera_config, "to_json"):
        try:
            import copy

            cc = copy.deepcopy(camera_config)
            if window_size is not None:
                cc.window_size = window_size
            ds.attrs["camera_config"] = cc.to_json()
        except Exception:  # a foreign object that cannot be copied / serialised keeps the frames' own attribute
            pass
    for k in ("v_x", "v_y", "corr", "s2n"):
        try:
            ds[k].encoding = dict(ENCODING_PARAMS)
        except Exception:
            pass
    return ds


# per-call options of a ``get_piv(engine="b200", ...)`` in flight in THIS thread / task (None: not a b200 call)
_B200_CALL = contextvars.ContextVar("pyorc_b200_call", default=None)
_B200_KWARGS = ("device", "devices", "coarse_pass", "multipass")


def install():
    """Register ``engine="b200"`` in an importable pyorc (see INTEGRATION.md for the two-line upstream patch).

    Two wrappers are installed ONCE and stay: ``Frames.get_piv`` accepts ``engine="b200"`` (plus ``device=``, ``devices=``,
    ``coarse_pass=``, ``multipass=``), runs the reference body unchanged (window rounding, coords, attrs, encoding - frames.py:156-197) with
    ``engine="numba"`` passing its whitelist (frames.py:176-177) and marks the call in a context variable;
    ``pyorc.velocimetry.ffpiv.get_ffpiv`` - looked up at call time at frames.py:186 - sends a marked call to
    :func:`pyorc_b200.velocimetry.get_b2piv` and every other call to the original.  Context variables are per thread (and per
    asyncio task), so concurrent ``get_piv`` calls with different engines - pyorc under dask's threaded scheduler - do
    not see each other, and no module global is swapped at call time.
    """
    import pyorc.api.frames as ref_frames  # noqa: PLC0415  (only present in a pyorc installation)
    import pyorc.velocimetry.ffpiv as ref_ffpiv  # noqa: PLC0415

    orig_get_piv = ref_frames.Frames.get_piv
    if getattr(orig_get_piv, "_b200", False):
        return True
    orig_get_ffpiv = ref_ffpiv.get_ffpiv

    def get_ffpiv(frames, y, x, dt, *args, **kwargs):
        call = _B200_CALL.get()
        if call is None:
            return orig_get_ffpiv(frames, y, x, dt, *args, **kwargs)
        kwargs = {**kwargs, **call, "engine": "b200"}
        return get_b2piv(frames, y, x, dt, *args, **kwargs)

    def get_piv_b200(self, window_size=None, overlap=None, engine="numba", ensemble_corr=False, **kwargs):
        if engine != "b200":
            return orig_get_piv(self, window_size=window_size, overlap=overlap, engine=engine, ensemble_corr=ensemble_corr, **kwargs)
        call = {k: kwargs.pop(k) for k in _B200_KWARGS if k in kwargs}
        token = _B200_CALL.set(call)
        try:
            return orig_get_piv(self, window_size=window_size, overlap=overlap, engine="numba", ensemble_corr=ensemble_corr, **kwargs)
        finally:
            _B200_CALL.reset(token)

    get_ffpiv.__doc__ = orig_get_ffpiv.__doc__
    get_ffpiv._b200_original = orig_get_ffpiv
    get_piv_b200._b200 = True
    get_piv_b200._b200_original = orig_get_piv
    get_piv_b200.__doc__ = orig_get_piv.__doc__
    ref_ffpiv.get_ffpiv = get_ffpiv
    ref_frames.Frames.get_piv = get_piv_b200
    return True


def uninstall():
    """Undo :func:`install` (tests)."""
    import pyorc.api.frames as ref_frames  # noqa: PLC0415
    import pyorc.velocimetry.ffpiv as ref_ffpiv  # noqa: PLC0415

    if getattr(ref_frames.Frames.get_piv, "_b200", False):
        ref_frames.Frames.get_piv = ref_frames.Frames.get_piv._b200_original
    if hasattr(ref_ffpiv.get_ffpiv, "_b200_original"):
        ref_ffpiv.get_ffpiv = ref_ffpiv.get_ffpiv._b200_original

"""``get_piv`` - B200 counterpart of ``pyorc.api.frames.Frames.get_piv`` (pyorc/api/frames.py:114-197).

Everything above the engine call (window rounding, default overlap, result coordinates, engine whitelist and its
``ValueError``) follows the reference; the engine call goes to :func:`pyorc_b200.velocimetry.get_b2piv`.
``install()`` registers ``engine="b200"`` inside a real pyorc installation.
"""

from __future__ import annotations

from typing import Optional

import numpy as np

from . import window
from .velocimetry import get_b2piv

__all__ = ["get_piv", "get_piv_coords", "install"]

ENGINES = ["b200"]


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
    ds.attrs = dict(getattr(frames, "attrs", {}))
    return ds


def install():
    """Register ``engine="b200"`` in an importable pyorc (see INTEGRATION.md for the two-line upstream patch).

    ``Frames.get_piv`` is wrapped: for ``engine="b200"`` the reference body runs unchanged (window rounding, coords,
    attrs, encoding - frames.py:156-197) with ``engine="numba"`` passing its whitelist (frames.py:176-177), while
    ``pyorc.velocimetry.ffpiv.get_ffpiv`` - looked up at call time at frames.py:186 - is redirected to
    :func:`pyorc_b200.velocimetry.get_b2piv` for the duration of the call.
    """
    import pyorc.api.frames as ref_frames  # noqa: PLC0415  (only present in a pyorc installation)
    import pyorc.velocimetry.ffpiv as ref_ffpiv  # noqa: PLC0415

    orig_get_piv = ref_frames.Frames.get_piv
    if getattr(orig_get_piv, "_b200", False):
        return True

    def get_piv_b200(self, window_size=None, overlap=None, engine="numba", ensemble_corr=False, **kwargs):
        if engine != "b200":
            return orig_get_piv(self, window_size=window_size, overlap=overlap, engine=engine, ensemble_corr=ensemble_corr, **kwargs)
        device = kwargs.pop("device", 0)
        saved = ref_ffpiv.get_ffpiv

        def forced(frames, y, x, dt, *a, engine="numba", **kw):
            return get_b2piv(frames, y, x, dt.values, *a, engine="b200", device=device, **kw)

        ref_ffpiv.get_ffpiv = forced
        try:
            return orig_get_piv(self, window_size=window_size, overlap=overlap, engine="numba", ensemble_corr=ensemble_corr, **kwargs)
        finally:
            ref_ffpiv.get_ffpiv = saved

    get_piv_b200._b200 = True
    get_piv_b200.__doc__ = orig_get_piv.__doc__
    ref_frames.Frames.get_piv = get_piv_b200
    return True

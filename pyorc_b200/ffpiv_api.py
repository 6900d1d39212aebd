"""The three names pyorc imports from ffpiv - ``cross_corr``, ``u_v_displacement``, ``window``
(pyorc/velocimetry/ffpiv.py:9, pyorc/api/frames.py:10) - served by the B200 engine with ffpiv's call signatures, so
``from pyorc_b200.ffpiv_api import cross_corr, u_v_displacement, window`` is a line-for-line substitute in pyorc's own
binding.  The fast path never materialises the planes (use :func:`pyorc_b200.velocimetry.get_b2piv`); this module is
the compatible, plane-returning form of the same kernels.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import window  # noqa: F401  (re-exported, mirrors ``ffpiv.window``)
from .engine import get_engine

__all__ = ["cross_corr", "u_v_displacement", "window"]


def cross_corr(
    imgs: np.ndarray,
    window_size: Tuple[int, int] = (64, 64),
    overlap: Tuple[int, int] = (32, 32),
    search_area_size: Optional[Tuple[int, int]] = None,
    engine: str = "b200",
    normalize: bool = False,
    verbose: bool = False,
    signal_threshold: Optional[float] = None,
    device: int = 0,
):
    """``x, y, corr`` as ``ffpiv.cross_corr`` returns them (call sites pyorc/velocimetry/ffpiv.py:222-231, :450-459):
    window-centre column / row vectors and float32 planes ``[n-1, n_rows*n_cols, wy, wx]`` (fftshifted, /N, clipped to
    [0, 1]; NaN planes for windows below ``signal_threshold``)."""
    if search_area_size is not None and tuple(search_area_size) != tuple(window_size):
        raise NotImplementedError("search_area_size must equal window_size (pyorc/api/frames.py:168)")
    if normalize:
        raise NotImplementedError("pyorc always passes normalize=False (pyorc/velocimetry/ffpiv.py:227,455)")
    if engine not in ("b200", "numba", "numpy"):
        raise ValueError(f"Selected PIV engine {engine} does not exist.")
    imgs = np.asarray(imgs)
    x, y = window.get_rect_coordinates(imgs.shape[-2:], tuple(window_size), tuple(overlap))
    corr = get_engine(device).corr_planes(imgs, tuple(window_size), tuple(overlap), signal_threshold=signal_threshold)
    return x, y, corr


def u_v_displacement(corr: np.ndarray, n_rows: int, n_cols: int, engine: str = "b200", device: int = 0):
    """``u, v`` in pixels, ``[..., n_rows, n_cols]``, from planes ``[..., n_rows*n_cols, wy, wx]`` (ffpiv.py:324,471)."""
    corr = np.asarray(corr)
    u, v = get_engine(device).peaks(corr)
    lead = corr.shape[:-3]
    return u.reshape(lead + (n_rows, n_cols)), v.reshape(lead + (n_rows, n_cols))

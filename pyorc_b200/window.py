"""Window geometry and memory model - host-side mirror of ``ffpiv.window`` as pyorc uses it.

Reference call sites: ``window.round_to_even`` (pyorc/api/frames.py:167), ``window.get_rect_coordinates``
(pyorc/api/frames.py:85-90), ``window.required_memory`` / ``window.available_memory``
(pyorc/velocimetry/ffpiv.py:120-129).  Same names, argument meaning and return conventions.
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "round_to_even",
    "get_axis_shape",
    "get_array_shape",
    "get_axis_coords",
    "get_rect_coordinates",
    "required_memory",
    "available_memory",
]


def round_to_even(input_tuple):
    """Round each entry to an even integer; odd values go up to the next even one (frames.py:167).  ffpiv's own rule for odd
    sizes is not observable in pyorc's source or tests; rounding up keeps the requested window covered."""
    out = []
    for x in input_tuple:
        r = int(round(x))
        out.append(r if r % 2 == 0 else r + 1)
    return tuple(out)


def get_axis_shape(dim_size: int, window_size: int, overlap: int) -> int:
    """Number of interrogation windows along one axis."""
    if window_size <= overlap:
        raise ValueError("overlap must be smaller than window_size")
    return int((dim_size - window_size) // (window_size - overlap) + 1)


def get_array_shape(dim_size, window_size, overlap):
    """``(n_rows, n_cols)`` of the velocity field."""
    return (
        get_axis_shape(dim_size[0], window_size[0], overlap[0]),
        get_axis_shape(dim_size[1], window_size[1], overlap[1]),
    )


def get_axis_coords(dim_size: int, window_size: int, overlap: int) -> np.ndarray:
    """Integer centre coordinate of every window along one axis: ``i*(w-o) + w//2`` (int64, pyorc indexes with it)."""
    n = get_axis_shape(dim_size, window_size, overlap)
    return np.int64(np.arange(n) * (window_size - overlap) + window_size / 2.0)


def get_rect_coordinates(dim_size, window_size, overlap, search_area_size=None):
    """``(cols_vector, rows_vector)`` of window centres (frames.py:85-90)."""
    if search_area_size is not None and tuple(search_area_size) != tuple(window_size):
        raise NotImplementedError("search_area_size must equal window_size (pyorc/api/frames.py:168)")
    y = get_axis_coords(dim_size[0], window_size[0], overlap[0])
    x = get_axis_coords(dim_size[1], window_size[1], overlap[1])
    return x, y


def required_memory(n_frames, dim_size, window_size, overlap, search_area_size=None, safety=1.0) -> float:
    """DEVICE bytes the B200 engine needs for ``n_frames`` frames (ffpiv.py:120-126 sizes the CPU's window stack +
    correlation planes; the fused engine holds neither): resident frames + four float32 result fields."""
    n_rows, n_cols = get_array_shape(dim_size, window_size, overlap)
    itemsize = 4  # worst case float32 frames
    return safety * (n_frames * dim_size[0] * dim_size[1] * itemsize + max(n_frames - 1, 0) * n_rows * n_cols * 16)


_FREE_CACHE = {}          # device -> (monotonic time of the query, free bytes)


def available_memory(device=None, need: float = 0.0, max_age: float = 10.0) -> float:
    """Free HBM in bytes of CUDA device ``device`` (default: the current one); ffpiv.py:129 asks for free host RAM.

    ``cudaMemGetInfo`` is no cheap call (measured 0.9 ms between engine calls and up to 7 ms on an idle B200 - a 100-pair 1080p
    ``get_b2piv`` takes 4.6 ms), so a caller that only needs to know whether ``need`` bytes fit passes it: a value younger than
    ``max_age`` seconds is reused as long as it is at least TWICE ``need``, i.e. the decision would survive the free memory
    halving since the query.  ``need = 0`` (default) always asks the driver."""
    import time

    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("pyorc_b200 needs a CUDA device (no CPU fallback)")
    key = torch.cuda.current_device() if device is None else int(device)
    now = time.monotonic()
    hit = _FREE_CACHE.get(key)
    if need > 0 and hit is not None and now - hit[0] < max_age and hit[1] >= 2.0 * need:
        return hit[1]
    free, _total = torch.cuda.mem_get_info(key)
    _FREE_CACHE[key] = (now, float(free))
    return float(free)

"""CPU oracle for the pre-processing filters - TEST INFRASTRUCTURE, never imported by pyorc_b200.

Restates, with numpy + OpenCV (the reference's own dependency for these filters), the arithmetic of pyorc's ``Frames``
methods on plain arrays (the reference wraps the same expressions in xarray/dask):

* ``normalize``   pyorc/api/frames.py:297-306
* ``minmax``      pyorc/api/frames.py:361
* ``time_diff``   pyorc/api/frames.py:424-430
* ``smooth``      pyorc/api/frames.py:449-466 -> pyorc/cv.py:158   cv2.GaussianBlur(img.astype("float32"), (k, k), 0)
* ``edge_detect`` pyorc/api/frames.py:326-341 -> pyorc/cv.py:180-183
"""

from __future__ import annotations

import numpy as np


def normalize(frames: np.ndarray, samples: int = 15) -> np.ndarray:
    time_interval = round(len(frames) / samples)
    assert time_interval != 0, f"Amount of frames is too small to provide {samples} samples"
    mean = frames[::time_interval].mean(axis=0).astype("float32")
    frames_reduce = frames.astype("float32") - mean
    frames_min = frames_reduce.min(axis=-1).min(axis=-1)[:, None, None]
    frames_max = frames_reduce.max(axis=-1).max(axis=-1)[:, None, None]
    with np.errstate(all="ignore"):
        return np.nan_to_num((frames_reduce - frames_min) / (frames_max - frames_min) * 255, nan=0.0).astype("uint8")


def minmax(frames: np.ndarray, min=-np.inf, max=np.inf) -> np.ndarray:  # noqa: A002
    return np.maximum(np.minimum(frames, max), min)   # pyorc/api/frames.py:361 verbatim: float64 whenever a bound is a Python float


def time_diff(frames: np.ndarray, thres: float = 0.0, abs: bool = False) -> np.ndarray:  # noqa: A002
    d = np.diff(frames.astype(np.float32), axis=0)
    d = np.where(d > thres, d, 0.0).astype(np.float32)
    return np.abs(d) if abs else d


def smooth(frames: np.ndarray, wdw: int = 1) -> np.ndarray:
    import cv2

    k = 2 * wdw + 1
    return np.stack([cv2.GaussianBlur(f.astype("float32"), (k, k), 0) for f in frames])


def edge_detect(frames: np.ndarray, wdw_1: int = 1, wdw_2: int = 2) -> np.ndarray:
    import cv2

    k1, k2 = 2 * wdw_1 + 1, 2 * wdw_2 + 1
    out = []
    for f in frames:
        b1 = cv2.GaussianBlur(f.astype("float32"), (k1, k1), 0)
        b2 = cv2.GaussianBlur(f.astype("float32"), (k2, k2), 0)
        out.append(b2 - b1)
    return np.stack(out)

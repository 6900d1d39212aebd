"""Frame pre-processing on the GPU - counterparts of pyorc's ``Frames`` filters that run before ``get_piv``
(pyorc/api/frames.py:279-466), so that frames stay in HBM between pre-processing and the PIV engine.

Same names, arguments and result dtypes as the reference methods:

* :func:`normalize`   <- ``Frames.normalize(samples=15)``        frames.py:279-306   -> uint8
* :func:`time_diff`   <- ``Frames.time_diff(thres=0.0, abs=False)`` frames.py:403-430 -> float32, one frame less
* :func:`minmax`      <- ``Frames.minmax(min, max)``               frames.py:343-361   -> same dtype
* :func:`smooth`      <- ``Frames.smooth(wdw=1)``                  frames.py:432-466   -> float32
* :func:`edge_detect` <- ``Frames.edge_detect(wdw_1=1, wdw_2=2)``  frames.py:308-341   -> float32

numpy in -> numpy out (copied through the device); torch CUDA tensor in -> torch CUDA tensor out (stream-ordered, no
copy), ready for :meth:`pyorc_b200.engine.Engine.pairs`.  No CPU fallback.
"""

from __future__ import annotations

import numpy as np

from .engine import B2PIV_F32, B2PIV_U8, Engine, get_engine

__all__ = ["normalize", "time_diff", "minmax", "smooth", "edge_detect"]


def _to_device(frames, eng: Engine):
    import torch

    if type(frames).__module__.startswith("torch"):
        if not frames.is_cuda:
            raise TypeError("torch frames must be CUDA tensors (pass numpy for host data)")
        t, was_np = frames, False
    else:
        a = np.asarray(frames)
        if a.dtype not in (np.uint8, np.float32):
            a = a.astype(np.float32)
        t, was_np = torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{eng.device}"), True
    if t.dim() != 3:
        raise ValueError("frames must be [time, y, x]")
    if t.dtype not in (torch.uint8, torch.float32):
        t = t.float()
    return t.contiguous(), was_np


def _code(t):
    import torch

    return B2PIV_U8 if t.dtype == torch.uint8 else B2PIV_F32


def _stream(t):
    import torch

    return torch.cuda.current_stream(t.device).cuda_stream


def _back(t, was_np):
    return t.cpu().numpy() if was_np else t


def normalize(frames, samples: int = 15, device: int = 0):
    """Remove the temporal mean of ``samples`` evenly spaced frames and stretch every frame to 0..255 (uint8)."""
    import torch

    eng = get_engine(device)
    t, was_np = _to_device(frames, eng)
    n, H, W = t.shape
    time_interval = round(n / samples)        # frames.py:297 (Python round: half to even)
    assert time_interval != 0, f"Amount of frames is too small to provide {samples} samples"
    out = torch.empty((n, H, W), dtype=torch.uint8, device=t.device)
    eng._check(eng._lib.b2piv_pre_normalize_device(eng._h, t.data_ptr(), _code(t), n, H, W, int(time_interval), out.data_ptr(), _stream(t)),
               "b2piv_pre_normalize_device")
    return _back(out, was_np)


def time_diff(frames, thres: float = 0.0, abs: bool = False, device: int = 0):  # noqa: A002  (reference argument name)
    """``frame[k+1] - frame[k]`` in float32; differences not larger than ``thres`` become 0; optional absolute value."""
    import torch

    eng = get_engine(device)
    t, was_np = _to_device(frames, eng)
    n, H, W = t.shape
    if n < 2:
        raise ValueError("need at least 2 frames")
    out = torch.empty((n - 1, H, W), dtype=torch.float32, device=t.device)
    eng._check(eng._lib.b2piv_pre_time_diff_device(eng._h, t.data_ptr(), _code(t), n, H, W, float(thres), int(bool(abs)), out.data_ptr(),
                                                   _stream(t)), "b2piv_pre_time_diff_device")
    return _back(out, was_np)


def minmax(frames, min=-np.inf, max=np.inf, device: int = 0):  # noqa: A002
    """Bound intensities to ``[min, max]``: ``np.maximum(np.minimum(frames, max), min)`` (pyorc/api/frames.py:343-361).
    VALUES are the reference's.  dtype: numpy promotes uint8 frames to float64 as soon as a bound is a Python float - the
    +-inf defaults included - although the clamped values stay whole numbers unless a bound has a fraction; here uint8 frames
    stay uint8 (a quarter of the bytes, the integer PIV path) when the finite bounds are whole numbers, and are clamped as
    float32 when a bound has a fraction (``min=10.5`` clamps a pixel to 10.5, not 10)."""
    import torch

    eng = get_engine(device)
    t, was_np = _to_device(frames, eng)
    if t.dtype == torch.uint8 and any(np.isfinite(b) and float(b) != np.floor(b) for b in (min, max)):
        t = t.float()
    out = torch.empty_like(t)
    lo = float(np.clip(min, -3.0e38, 3.0e38))
    hi = float(np.clip(max, -3.0e38, 3.0e38))
    eng._check(eng._lib.b2piv_pre_minmax_device(eng._h, t.data_ptr(), _code(t), t.numel(), lo, hi, out.data_ptr(), _stream(t)),
               "b2piv_pre_minmax_device")
    return _back(out, was_np)


def _gauss(frames, k1: int, k2: int, device: int):
    import torch

    eng = get_engine(device)
    t, was_np = _to_device(frames, eng)
    n, H, W = t.shape
    out = torch.empty((n, H, W), dtype=torch.float32, device=t.device)
    eng._check(eng._lib.b2piv_pre_gauss_device(eng._h, t.data_ptr(), _code(t), n, H, W, int(k1), int(k2), out.data_ptr(), _stream(t)),
               "b2piv_pre_gauss_device")
    return _back(out, was_np)


def smooth(frames, wdw: int = 1, device: int = 0):
    """Gaussian smoothing with a ``(2*wdw+1)``-square kernel, ``cv2.GaussianBlur(img, (k, k), 0)`` semantics; float32."""
    return _gauss(frames, 0, 2 * int(wdw) + 1, device)


def edge_detect(frames, wdw_1: int = 1, wdw_2: int = 2, device: int = 0):
    """Band filter: ``GaussianBlur(2*wdw_2+1) - GaussianBlur(2*wdw_1+1)``; float32."""
    return _gauss(frames, 2 * int(wdw_1) + 1, 2 * int(wdw_2) + 1, device)
